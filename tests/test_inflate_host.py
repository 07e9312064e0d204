"""The DEFLATE decoder of the device BGZF path (breakid_b200/csrc/bkid_inflate.cuh) compiled as plain C++ and
checked against zlib on the CPU: stored / fixed / dynamic blocks, long codes, overlapping matches of every short
distance, every alignment of the input and output windows, tiny outputs, corrupt input."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("inflate") / "libinflate_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, os.path.join(ROOT, "tests", "host", "inflate_host.cc")])
    L = C.CDLL(out)
    L.bki_host_inflate.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.bki_host_inflate.restype = C.c_int
    L.bki_host_inflate_var.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    L.bki_host_inflate_var.restype = C.c_int
    L.bki_host_crc32_sliced.argtypes = [C.c_char_p, C.c_uint32, C.c_int]
    L.bki_host_crc32_sliced.restype = C.c_uint32
    return L


def _raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem=8):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, mem, strategy)
    return c.compress(data) + c.flush()


def _inflate(lib, comp, n):
    out = np.zeros(max(n, 1), np.uint8)
    rc = lib.bki_host_inflate(comp, len(comp), out.ctypes.data, n)
    return rc, out[:n].tobytes()


def _samples():
    rng = np.random.RandomState(3)
    yield b""
    yield b"a"
    yield b"abc" * 1000                                    # overlapping matches, distance 3
    yield b"\x00" * 65280                                  # distance-1 runs of length 258
    yield rng.randint(0, 256, 65280, dtype=np.uint8).tobytes()          # incompressible
    yield rng.randint(0, 4, 65280, dtype=np.uint8).tobytes()            # 2-bit alphabet: short codes
    yield bytes(rng.choice(np.arange(256, dtype=np.uint8), 60000, p=np.r_[np.full(16, 0.05), np.full(240, 0.2 / 240)]))   # skewed: long codes
    # BAM-like records: names, packed sequence, qualities, tags
    recs = []
    for i in range(400):
        recs.append(b"r%010d\0" % i + rng.randint(0, 256, 75, dtype=np.uint8).tobytes() + bytes(rng.randint(2, 41, 150, dtype=np.uint8)) + b"NMC\x01MDZ150\0SAZchr2,12345,+,60M90S,60,0;\0")
    yield b"".join(recs)[:65280]
    yield rng.randint(0, 256, 300000, dtype=np.uint8).tobytes() + b"xyz" * 50000      # several deflate blocks, beyond 64 KiB


def test_matches_zlib(lib):
    n = 0
    for data in _samples():
        for level, strategy in ((0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)):
            for mem in (1, 8):
                comp = _raw(data, level, strategy, mem)
                rc, out = _inflate(lib, comp, len(data))
                assert rc == 0, (len(data), level, strategy, mem, rc)
                assert out == data, (len(data), level, strategy, mem)
                out2 = np.zeros(max(len(data), 1), np.uint8)            # one literal per step; misaligned input / output windows
                for litmax, im, om in ((1, 0, 0), (4, 1, 3), (4, 3, 7), (1, 2, 5)):
                    assert lib.bki_host_inflate_var(comp, len(comp), out2.ctypes.data, len(data), litmax, im, om) == 0
                    assert out2[:len(data)].tobytes() == data, (len(data), level, strategy, mem, litmax, im, om)
                n += 1
    assert n > 100


def test_rejects_corrupt_streams(lib):
    rng = np.random.RandomState(5)
    data = (b"the quick brown fox " * 500) + rng.randint(0, 256, 5000, dtype=np.uint8).tobytes()
    comp = _raw(data)
    # wrong declared size
    assert _inflate(lib, comp, len(data) - 1)[0] != 0
    assert _inflate(lib, comp, len(data) + 1)[0] != 0
    # truncated payload
    assert _inflate(lib, comp[: len(comp) // 2], len(data))[0] != 0
    # reserved block type
    assert _inflate(lib, b"\x07" + comp[1:], len(data))[0] != 0
    # random bit flips never crash; when the damaged stream still decodes to the declared size, zlib reads the same bytes
    bad = 0
    for t in range(200):
        c = bytearray(comp)
        c[rng.randint(0, len(c))] ^= 1 << rng.randint(0, 8)
        rc, out = _inflate(lib, bytes(c), len(data))
        if rc != 0:
            bad += 1
        else:
            try:
                assert zlib.decompress(bytes(c), -15) == out      # zlib agrees on what the damaged stream says
            except zlib.error:
                pass                                               # zlib is stricter (e.g. incomplete code sets)
    assert bad >= 1         # most flips only change a literal: catching those is the CRC's job, not the decoder's


def test_sliced_crc32_matches_zlib(lib):
    rng = np.random.RandomState(11)
    for n in (0, 1, 2, 31, 32, 33, 1000, 65279, 65280, 65536):
        data = rng.randint(0, 256, n, dtype=np.uint8).tobytes()
        for k in (1, 2, 8, 32):
            assert lib.bki_host_crc32_sliced(data, n, k) == (zlib.crc32(data) & 0xffffffff), (n, k)
    assert lib.bki_host_crc32_sliced(b"\x00" * 5000, 5000, 32) == (zlib.crc32(b"\x00" * 5000) & 0xffffffff)


def test_every_short_distance_length_and_alignment(lib):
    """the 8-byte pending word: matches of distance 1..40 and lengths 3..30 at all eight output alignments, and
    outputs of 0..20 bytes (first word == last word) -- bytes outside the output window must stay untouched"""
    rng = np.random.RandomState(17)
    for dist in list(range(1, 20)) + [23, 24, 25, 31, 32, 33, 40]:
        seedb = rng.randint(0, 256, dist, dtype=np.uint8).tobytes()
        data = b""
        for ln in range(3, 31):
            data += rng.randint(0, 256, 3, dtype=np.uint8).tobytes() + seedb + (seedb * 40)[:ln]
        comp = _raw(data, 9)
        out = np.zeros(len(data), np.uint8)
        for om in range(8):
            assert lib.bki_host_inflate_var(comp, len(comp), out.ctypes.data, len(data), 4, om & 3, om) == 0
            assert out.tobytes() == data, (dist, om)
    for n in range(0, 21):
        data = rng.randint(0, 4, n, dtype=np.uint8).tobytes()
        for level in (0, 6):
            comp = _raw(data, level)
            out = np.zeros(max(n, 1), np.uint8)
            for om in range(8):
                assert lib.bki_host_inflate_var(comp, len(comp), out.ctypes.data, n, 4, 0, om) == 0, (n, level, om)
                assert out[:n].tobytes() == data
