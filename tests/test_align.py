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


def _validity_by_oracle(hb, genome, reads):
    """which SA records the validator must keep: python restatement of k7_validate_rows' mapping + the oracle DP"""
    import oracle_py as O
    from breakid_b200 import synth
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    s = hb.side
    names = {n: t for t, n in enumerate(hb.target_names)}
    keep = np.ones(hb.n_sa, bool)
    for k in range(hb.n_sa):
        i = int(s["sa_rec"][k])
        ops = [(int(o) >> 4, int(o) & 0xf) for o in s["cig_ops"][int(s["cig_off"][k]):int(s["cig_off"][k + 1])]]
        read = reads[k]
        if ops[0][1] == synth.OP_S: q = read[:ops[0][0]]
        elif ops[-1][1] == synth.OP_S: q = read[len(read) - ops[-1][0]:]
        else: continue
        f = bytes(s["sa_txt"][int(s["sa_off"][k]):int(s["sa_off"][k + 1])]).split(b";")[0].split(b",")
        t, sa_pos, sa_minus = names[f[0].decode()], int(f[1]), f[2] == b"-"
        m, num = 0, b""
        for ch in f[3]:
            if 48 <= ch <= 57: num += bytes([ch])
            else:
                if ch in b"M": m += int(num)
                num = b""
        ref = bytes(genome[t][sa_pos - 1:sa_pos - 1 + m])
        if sa_minus != bool(int(hb.cols["flag"][i]) & 0x10):
            q = q.translate(comp)[::-1]
        dist = O.banded_edit(q, ref, 12)
        keep[k] = dist >= 0 and dist * 10 <= len(q) + 20
    return keep


@pytest.mark.gpu
def test_validate_align_drops_split_reads_whose_clip_does_not_align(small_data):
    """default-off evidence validator (bkid_params.validate_align): with true read bases nothing changes; split reads
    whose clipped bases are random stop being evidence -- identical to the oracle on the same batch with the SA tags
    of exactly those records removed"""
    import oracle_py as O
    from breakid_b200 import api, synth
    d, hb0, nibs = small_data
    genome = [synth.nib_ascii(p, l) for p, l in nibs]
    rng = np.random.RandomState(3)
    for frac in (0.0, 0.35):
        hb = api.HostBatch(hb0.cols, hb0.name_hash, hb0.side, hb0.target_len, hb0.target_names)
        corrupt = set(np.nonzero(rng.rand(hb.n_sa) < frac)[0].tolist())
        seq, reads = synth.split_read_sequences(hb, genome, corrupt=corrupt, seed=9)
        hb.set_seq(seq)
        keep = _validity_by_oracle(hb, genome, reads)
        assert (~keep).sum() == len(corrupt)                     # random 60+ base clips never align by chance
        # expected: SA text of rejected records removed -> they are ordinary records (still counted for coverage)
        side = dict(hb.side)
        so = side["sa_off"].astype(np.int64)
        parts = [side["sa_txt"][so[k]:so[k + 1]] if keep[k] else side["sa_txt"][:0] for k in range(hb.n_sa)]
        side["sa_txt"] = np.concatenate(parts) if parts else side["sa_txt"]
        side["sa_off"] = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.uint32)
        hb_exp = api.HostBatch(hb.cols, hb.name_hash, side, hb.target_len, hb.target_names)
        om, osd, od, exp = O.run(hb_exp, nibs, mode=0)
        c = api.Context(hb.target_len, hb.target_names, device=0, validate_align=1)
        c.push(hb)
        for t, (p, l) in enumerate(nibs):
            c.set_nib(t, p, l)
        mean, sd, dist, ncall = c.run()
        got = c.fetch_clusters()
        assert (mean, sd, dist) == (om, osd, od)
        assert got.tobytes() == exp.tobytes(), frac
        if frac == 0.0:
            base = O.run(hb0, nibs, mode=0)[3]
            assert got.tobytes() == base.tobytes()               # validator on, all evidence true: nothing changes
        else:
            assert int(got["n_split_read"].sum()) < int(O.run(hb0, nibs, mode=0)[3]["n_split_read"].sum())
        c.close()
    # the flag without read bases is an error, not a silent no-op
    c = api.Context(hb0.target_len, hb0.target_names, device=0, validate_align=1)
    c.push(hb0)
    with pytest.raises(api.BkidError, match="read bases"):
        c.run()
    c.close()


@pytest.mark.gpu
def test_driver_validate_flag(tmp_path):
    """BreakID -validate (device decode extracts the read bases): with true bases the call file is the reference binary's;
    with a third of the split reads carrying random clips the supporting split-read counts drop"""
    import os
    import subprocess
    import oracle_py as O
    from breakid_b200 import api, bamio, synth
    cfg = synth.SynthConfig(chrom_lens=[300000, 200000, 150000], n_tra=3, n_inv=2, n_dup=2, n_del=2, seed=21, sv_jitter=1)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    genome = [synth.nib_ascii(synth.random_nib_bytes(l, cfg.seed * 1000 + t), l) for t, l in enumerate(cfg.chrom_lens)]
    drv = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "breakid_b200", "host", "BreakID")
    outs = {}
    for tag, corrupt in (("true", set()), ("bad", set(range(0, hb.n_sa, 3)))):
        seq, _ = synth.split_read_sequences(hb, genome, corrupt=corrupt, seed=2)
        wd = tmp_path / tag
        wd.mkdir()
        paths = bamio.write_dataset(str(wd), d, genes_per_mb=25.0)
        bamio.write_bam(paths["bam"], d, sa_seq=seq)
        open(paths["bam"] + ".bai", "wb").close()                     # the driver only checks that an index file exists
        for flags in ([], ["-validate"]):
            g = subprocess.run([drv, "-i", paths["bam"], "-o", str(wd / ("out" + "".join(flags))), "-n", paths["nib"], "-r", paths["refgene"], "-all"] + flags,
                               timeout=600, capture_output=True, text=True)
            assert g.returncode == 0, g.stderr[-500:]
            outs[(tag, bool(flags))] = open(str(wd / ("out" + "".join(flags))) + "_fusion_all.txt").read()
    nsr = lambda txt: sum(int(l.split("\t")[8]) for l in txt.splitlines()[1:])
    assert outs[("true", True)] == outs[("true", False)] == outs[("bad", False)]      # the flag is neutral on true bases; off = bases ignored
    assert nsr(outs[("bad", True)]) < nsr(outs[("true", True)])
    if O.have_ref():
        O.ref_index(str(tmp_path / "true" / "reads.bam"))
        O.ref_install_refgene(str(tmp_path / "true" / "ref_files" / "refGene.txt"))
        r = O.ref_run_binary(str(tmp_path / "true" / "reads.bam"), str(tmp_path / "ref"), str(tmp_path / "true" / "nib"))
        assert r.returncode == 0 and open(str(tmp_path / "ref") + "_fusion_all.txt").read() == outs[("true", True)]
