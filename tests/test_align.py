"""Banded alignment operator (extension, north_star kernel 4): oracle DP against a plain full edit distance on the CPU,
CUDA kernel against the oracle on the GPU."""
import numpy as np
import pytest


def _edit(a: bytes, b: bytes) -> int:
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            sub = 0 if (a[i - 1] == b[j - 1] and a[i - 1] != ord("N")) else 1
            cur[j] = min(prev[j - 1] + sub, prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[len(b)]


def _pairs(rng, n, maxlen=140):
    out = []
    for t in range(n):
        m = int(rng.randint(0, maxlen))
        ref = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), m))
        q = bytearray(ref)
        for _ in range(int(rng.randint(0, 7))):                    # substitutions, insertions, deletions, Ns
            kind = int(rng.randint(0, 4))
            pos = int(rng.randint(0, len(q) + 1))
            if kind == 0 and pos < len(q): q[pos] = int(rng.choice(np.frombuffer(b"ACGT", np.uint8)))
            elif kind == 1: q.insert(pos, int(rng.choice(np.frombuffer(b"ACGT", np.uint8))))
            elif kind == 2 and pos < len(q): del q[pos]
            elif kind == 3 and pos < len(q): q[pos] = ord("N")
        if t % 11 == 0:
            q = bytearray(rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.randint(0, maxlen))))    # unrelated
        out.append((bytes(q), ref))
    out += [(b"", b""), (b"A", b""), (b"", b"ACG"), (b"NNNN", b"NNNN"), (b"ACGT" * 30, b"ACGT" * 30), (b"A" * 100, b"A" * 112)]
    return out


def test_oracle_banded_edit_equals_full_edit_distance():
    import oracle_py as O
    rng = np.random.RandomState(2)
    for q, r in _pairs(rng, 300, maxlen=60):
        full = _edit(q, r)
        assert O.banded_edit(q, r, 200) == full                      # band wider than both strings: plain edit distance
        for w in (0, 3, 8, 15):
            got = O.banded_edit(q, r, w)
            if abs(len(q) - len(r)) > w:
                assert got == -1
            else:
                assert got >= full                                   # restricting the paths can only cost more
                if full <= w // 2:
                    assert got == full                               # a path with <= w/2 indels never leaves the band


@pytest.mark.gpu
def test_banded_align_kernel_equals_oracle():
    import oracle_py as O
    from breakid_b200 import api
    c = api.Context([1000], ["chr1"], device=0)
    rng = np.random.RandomState(7)
    pairs = _pairs(rng, 3000) + [(b"ACGT" * 128, b"ACGT" * 128), (b"A" * 513, b"A" * 513), (b"C" * 512, b"C" * 505 + b"GGGGGGG")]
    for w in (0, 1, 4, 8, 15):
        got = c.op_banded_align([p[0] for p in pairs], [p[1] for p in pairs], w)
        for (q, r), g in zip(pairs, got):
            exp = -2 if max(len(q), len(r)) > 512 else O.banded_edit(q, r, w)
            assert int(g) == exp, (w, len(q), len(r), int(g), exp)
    assert len(c.op_banded_align([], [], 8)) == 0
    with pytest.raises(api.BkidError):
        c.op_banded_align([b"A"], [b"A"], 16)
    c.close()
