"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/breakid_b200.h
declares, refuses to run without a GPU (no CPU fallback), the host BAM decoder reproduces the
generated record batch, and the BreakID driver keeps the reference's argument errors."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from breakid_b200 import api
    hdr = open(os.path.join(ROOT, "include", "breakid_b200.h")).read()
    declared = set(re.findall(r"\b(bkid_[a-z_0-9]+)\s*\(", hdr)) - {"bkid_name_hash"}
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    lib = ctypes.CDLL(api.LIB_CUDA)
    for s in declared:
        assert hasattr(lib, s), s
    assert api.cuda_lib().bkid_abi_version() == 3


def test_no_cpu_fallback():
    import torch
    from breakid_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.BkidError) as e:
        api.Context([1000], ["chr1"])
    assert "no usable CUDA device" in str(e.value)


def test_pod_layouts_match_header():
    from breakid_b200 import api
    assert api.PAIR_DTYPE.itemsize == 64 and api.CLUSTER_DTYPE.itemsize == 192
    import ctypes
    assert ctypes.sizeof(api.Batch) == 28 * 8 and ctypes.sizeof(api.Params) == 32


def test_host_bam_decoder_roundtrip(tmp_path):
    from breakid_b200 import api, bamio, synth
    cfg = synth.SynthConfig(chrom_lens=[80000, 50000], n_tra=1, n_inv=1, n_dup=1, n_del=1, seed=3, min_sv_sep=4000)
    d = synth.generate(cfg)
    a = api.HostBatch.from_synth(d)
    bam = str(tmp_path / "x.bam")
    bamio.write_bam(bam, d, random_qual=True)
    for threads in (1, 4):
        b = api.HostBatch.from_bam(bam, threads)
        assert a.n == b.n and a.n_sa == b.n_sa
        for k in a.cols:
            assert np.array_equal(a.cols[k], b.cols[k]), k
        assert np.array_equal(a.name_hash, b.name_hash)
        assert a.n_x == b.n_x and 0 < a.n_x < a.n // 10
        for k in a.x:
            assert np.array_equal(a.x[k], b.x[k]), k
        for k in a.side:
            assert np.array_equal(a.side[k], b.side[k]), k
        assert b.target_names == ["chr1", "chr2"] and list(b.target_len) == [80000, 50000]


def test_name_hash_vectorised_equals_c():
    import torch
    from breakid_b200 import synth
    ids = torch.tensor([0, 1, 17, 123456789, 9999999999], dtype=torch.int64)
    got = synth.name_hash_ids(ids).numpy().view(np.uint64)
    for i, v in enumerate(ids.tolist()):
        lo, hi = synth.name_hash_py(synth.name_of(v))
        assert (int(got[i, 0]), int(got[i, 1])) == (lo, hi)


def test_driver_argument_errors():
    drv = os.path.join(ROOT, "breakid_b200", "host", "BreakID")
    if not os.path.exists(drv):
        pytest.skip("driver not built")
    r = subprocess.run([drv], timeout=600, capture_output=True, text=True)
    assert r.returncode == 1 and "Error: input- and output file is required." in r.stderr
    r = subprocess.run([drv, "-i", "a.bam", "-o", "x"], timeout=600, capture_output=True, text=True)
    assert r.returncode == 1 and "Error: nib file's root dir is required." in r.stderr
    r = subprocess.run([drv, "-i", "/nonexistent.bam", "-o", "x", "-n", "."], timeout=600, capture_output=True, text=True)
    assert r.returncode == 1 and "Error: can not open bam-file" in r.stderr
    r = subprocess.run([drv, "-h"], timeout=600, capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([drv, "-bogus"], timeout=600, capture_output=True, text=True)       # the reference segfaults here
    assert r.returncode == 1


def test_indexed_refgene_lookup_equals_linear_scan(tmp_path):
    """the interval index over refGene gives exactly what the reference-order linear scan gives
    (last overlapping transcript with a CDS wins, NR_ rows skipped, exon numbering), on nested / overlapping genes"""
    import ctypes as C
    import numpy as np
    from breakid_b200 import api, synth
    p = str(tmp_path / "refGene.txt")
    synth.write_refgene(p, [400000, 300000, 200000], genes_per_mb=120.0, seed=3)     # dense: many overlapping transcripts
    rows = open(p).read().splitlines()
    # add nested transcripts, a CDS-less one covering others, duplicates of a locus in different file positions
    extra = []
    for i, r in enumerate(rows[:40]):
        f = r.split("\t")
        f[1] = "NM_9%05d" % i; f[12] = "NEST%d" % i
        extra.append("\t".join(f))
        g = list(f); g[1] = "NM_8%05d" % i; g[6] = g[7] = g[4]; g[12] = "NOCDS%d" % i       # cdsStart == cdsEnd: no CDS
        extra.append("\t".join(g))
    open(p, "w").write("\n".join(rows[:60] + extra + rows[60:]) + "\n")
    L = api.host_lib()
    L.bkid_host_annotate_both.argtypes = [C.c_char_p, C.c_char_p, C.c_long, C.c_char_p, C.c_char_p, C.c_int]
    a, b = C.create_string_buffer(512), C.create_string_buffer(512)
    rng = np.random.RandomState(5)
    kinds = set()
    starts = [int(r.split("\t")[4]) for r in rows] + [int(r.split("\t")[5]) for r in rows]
    for q in range(6000):
        chrom = ["chr1", "chr2", "chr3", "chr9"][int(rng.randint(0, 4))]
        pos = int(rng.choice(starts)) + int(rng.randint(-2, 3)) if q % 3 == 0 else int(rng.randint(0, 420000))
        assert L.bkid_host_annotate_both(p.encode(), chrom.encode(), pos, a, b, 512) == 0
        assert a.value == b.value, (chrom, pos, a.value, b.value)
        kinds.add(a.value.split(b"\t")[0][:4])
    assert b"inte" in kinds and b"GENE" in kinds and b"NEST" in kinds


def test_host_decoder_narrow_batch_matches_python_narrowing(tmp_path):
    """bkid_host_bam_batch_narrow (C++) picks the same narrow encodings, with the same values, as HostBatch.narrow()"""
    import ctypes as C
    from breakid_b200 import api, bamio, synth
    cfg = synth.SynthConfig(chrom_lens=[80000, 50000], n_tra=1, n_inv=1, n_dup=0, n_del=1, seed=3)
    d = synth.generate(cfg)
    p = str(tmp_path / "t.bam")
    bamio.write_bam(p, d)
    hb = api.HostBatch.from_bam(p, threads=2)
    lib = api.host_lib()
    err = C.create_string_buffer(256)
    h = lib.bkid_host_read_bam(p.encode(), 2, err, 256)
    assert h
    try:
        b = C.cast(lib.bkid_host_bam_batch_narrow(h), C.POINTER(api.Batch)).contents
        nr = hb.narrow()
        assert set(nr) == {"span16", "isize16", "tid_run_start", "tid_run_tid"}
        assert not b.tid and not b.isize and not b.endpos and b.flag and b.pos          # wide columns replaced, others kept
        get = lambda ptr, n, dt: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (n * np.dtype(dt).itemsize,)).view(dt).copy()
        assert np.array_equal(get(b.span16, hb.n, np.uint16), nr["span16"])
        assert np.array_equal(get(b.isize16, hb.n, np.int16), nr["isize16"])
        assert int(b.n_tid_runs) == nr["tid_run_start"].shape[0]
        assert np.array_equal(get(b.tid_run_start, int(b.n_tid_runs), np.uint32), nr["tid_run_start"])
        assert np.array_equal(get(b.tid_run_tid, int(b.n_tid_runs), np.int32), nr["tid_run_tid"])
        assert b.seq_off and np.array_equal(get(b.seq_len, hb.n_sa, np.int32), hb.seq["seq_len"])
    finally:
        lib.bkid_host_bam_free(h)


def test_bucket_owner_table_is_balanced_and_deterministic():
    """bkid_lpt_owner_table (host part of the multi-GPU pair exchange): every non-empty bucket gets an owner < world, the
    estimated loads differ by at most the largest bucket, the same histogram gives the same table, world 1 owns everything"""
    import ctypes as C
    import numpy as np
    from breakid_b200 import api
    L = api.cuda_lib()
    rng = np.random.RandomState(4)
    for world in (1, 2, 3, 8, 32):
        for trial in range(20):
            nb = int(rng.choice([1, 7, 300, 625]))
            hist = (rng.pareto(1.2, nb) * 1000).astype(np.uint64) * (rng.rand(nb) < 0.8)
            hist = np.ascontiguousarray(hist, np.uint64)
            own = np.full(nb, 255, np.uint8); own2 = own.copy()
            assert L.bkid_lpt_owner_table(hist.ctypes.data, nb, world, own.ctypes.data) == 0
            assert L.bkid_lpt_owner_table(hist.ctypes.data, nb, world, own2.ctypes.data) == 0
            assert np.array_equal(own, own2) and int(own.max()) < world
            m = hist.astype(np.float64)
            cost = m * (1.0 + np.log2(m + 1.0) / 16.0)
            load = np.array([cost[own == r].sum() for r in range(world)])
            assert load.max() - load.min() <= cost.max() + 1e-6
    assert L.bkid_lpt_owner_table(None, 3, 2, None) != 0 and L.bkid_lpt_owner_table(hist.ctypes.data, nb, 33, own.ctypes.data) != 0
