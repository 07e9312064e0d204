"""CPU tests that pin the oracle against the REAL reference compiled from /root/reference
(oracle/_ref).  Skipped where the reference artefacts are absent (the GPU box has the prebuilt
files, so they run there too)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import lattice_points

import oracle_py as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    from breakid_b200 import api, bamio, synth
    tmp = str(tmp_path_factory.mktemp("ds"))
    cfg = synth.SynthConfig(chrom_lens=[200000, 140000], n_tra=2, n_inv=2, n_dup=1, n_del=2, seed=8, sv_jitter=2, min_sv_sep=5000)
    d = synth.generate(cfg)
    paths = bamio.write_dataset(tmp, d, random_qual=False, genes_per_mb=30.0)
    O.ref_index(paths["bam"])
    O.ref_install_refgene(paths["refgene"])
    hb = api.HostBatch.from_bam(paths["bam"], 4)
    return d, hb, paths


def test_sort_model_vs_reference_std_sort():
    R = O.rlib()
    rng = np.random.RandomState(0)
    for trial in range(400):
        n = int(rng.choice([0, 1, 2, 3, 15, 16, 17, 33, 100, 257, 1000, 3000])) if trial % 3 else int(rng.randint(0, 400))
        kind = trial % 5
        key = [rng.randint(0, 5, n), rng.randint(0, max(1, n // 3 + 1), n), np.sort(rng.randint(0, n + 1, n)),
               np.sort(rng.randint(0, n + 1, n))[::-1].copy(), rng.randint(0, 2 ** 32, n)][kind].astype(np.uint32)
        ref = np.zeros(n, np.uint32)
        R.ref_std_sort_perm(n, key, trial % 2 if kind == 4 else trial % 3, ref)     # cmp_enspan_id compares int: keep its keys < 2^31
        assert np.array_equal(ref, O.sort_perm(key)) and np.array_equal(ref, O.sort_perm(key, model=True)), (trial, n, kind)


def test_mask_fast_ahc_vs_reference():
    R = O.rlib()
    rng = np.random.RandomState(1)
    for trial in range(300):
        n = int(rng.randint(0, 90)); x, y = lattice_points(rng, n, trial % 4)
        w = float(rng.choice([120.7, 260.2, 99.0, 1243.15]))
        ref = np.zeros(n + 2, np.uint32)
        with O.quiet():
            k = R.ref_remove_isolated(n, x, y, w, ref)
        got = O.remove_isolated(x, y, w)
        assert np.array_equal(ref[:k], got), trial
        if k < 2:
            continue
        xs = np.ascontiguousarray(x[got]); ys = np.ascontiguousarray(y[got])
        for f, mode, model in ((R.ref_cluster_fast, 1, False), (R.ref_cluster_ahc, 0, False), (R.ref_cluster_ahc, 0, True)):
            ri = np.zeros(k + 2, np.uint32); rc = np.zeros(k + 2, np.int32); rr = C.c_int()
            with O.quiet():
                m = f(k, xs, ys, w, ri, rc, C.byref(rr), 0)
            gi, gc, gr = O.cluster(mode, xs, ys, w, model=model)
            assert np.array_equal(ri[:m], gi) and np.array_equal(rc[:m], gc) and rr.value == gr, (trial, mode, model)


def test_ahc_trees_vs_reference():
    R = O.rlib()
    rng = np.random.RandomState(7)
    for trial in range(600):
        n = int(rng.randint(2, 60)); kind = trial % 4
        if kind == 0:
            x = rng.randint(0, 5, n) * 50.; y = rng.randint(0, 5, n) * 50.
        elif kind == 1:
            k = max(1, n // 6); a = rng.randint(0, k, n); x = a * 1000. + rng.randint(0, 4, n) * 50; y = (a % 2) * 800. + rng.randint(0, 4, n) * 50
        elif kind == 2:
            k = max(1, n // 5); a = rng.randint(0, k, n); x = a * 500. + rng.randint(0, 3, n) * 50; y = rng.randint(0, 3, n) * 50. + (a % 3) * 400
        else:
            x = rng.randint(0, 300, n) * 1.; y = rng.randint(0, 300, n) * 1.
        thr = int(rng.choice([60, 120, 200, 260]))
        outs = []
        for f in (R.ref_ahc_tree, O.olib().orc_ahc_tree, O.olib().orc_model_ahc_tree):
            r = np.zeros(2 * n + 1, np.int32); a_ = np.zeros(2 * n + 1, np.int32); b_ = np.zeros(2 * n + 1, np.int32)
            with O.quiet():
                nn = f(n, np.ascontiguousarray(x), np.ascontiguousarray(y), thr, r, a_, b_)
            outs.append((nn, r[:nn].tobytes(), a_[:nn].tobytes(), b_[:nn].tobytes()))
        assert outs[0] == outs[1] == outs[2], (trial, kind, n, thr)


def test_stats_scan_on_bam(dataset):
    d, hb, paths = dataset
    m, s = C.c_double(), C.c_double()
    with O.quiet():
        O.rlib().ref_insert_stats(paths["bam"].encode(), C.byref(m), C.byref(s))
    om, osd, _, _, _ = O.insert_stats(hb)
    assert (om, osd) == (m.value, s.value)
    for qual in (20, 0, 45):
        w = O.dist(om, osd)
        assert O.scan(hb, qual, w).tobytes() == O.ref_scan(paths["bam"], qual, w, paths["nib"]).tobytes()


def _sorted_rows(a):
    return np.sort(a, order=["name_lo", "name_hi", "secondary", "primary_start"])


def test_split_read_evidence_vote_depth(dataset):
    d, hb, paths = dataset
    R = O.rlib()
    names = hb.target_names
    checked = 0
    for j in range(len(d.truth["type"])):
        A, a, B, b = (int(d.truth[k][j]) for k in ("A", "a", "B", "b"))
        for (t, p) in ((A, a), (B, b)):
            for half in (1200, 40, 3):
                s, e = max(0, p - half), p + half
                ref = O.ref_find_sa_reads(paths["bam"], names[t], s, e)
                got = O.find_sa_reads(hb, t, s, e)
                assert len(ref) == len(got)
                r, g = _sorted_rows(ref), _sorted_rows(got)
                for k in ref.dtype.names:
                    if k not in ("_pad",):
                        assert np.array_equal(r[k], g[k]), (k, j, half)
                checked += len(ref)
            depth = R.ref_single_base_depth(paths["bam"].encode(), names[t].encode(), p)
            assert depth == O.single_base_depth(hb, t, p)
        p1, p2 = C.c_int32(), C.c_int32()
        with O.quiet():
            v = R.ref_find_bp(paths["bam"].encode(), names[A].encode(), a - 1200, a + 1200, names[B].encode(), b - 1200, b + 1200, C.byref(p1), C.byref(p2))
        assert (v, p1.value, p2.value) == O.find_bp(hb, A, a - 1200, a + 1200, B, b - 1200, b + 1200)
    assert checked > 50


def test_nib_neighbour_sequences(dataset):
    from breakid_b200 import synth
    d, hb, paths = dataset
    R = O.rlib()
    for t, l in enumerate(d.cfg.chrom_lens):
        pay = synth.random_nib_bytes(l, d.cfg.seed * 1000 + t).numpy()
        for bp in (21, 22, 1000, 77777, l - 21, l - 20):
            buf = C.create_string_buffer(42)
            R.ref_neighbor_41(paths["nib"].encode(), hb.target_names[t].encode(), bp, buf)
            assert buf.value == O.neighbor_41(pay, l, bp), (t, bp)
    for s in (b"AAAAAAAAAAAAC", b"ACGT", b"TTTTTTTTTTTGGGGGGGGGGGGG"):
        assert R.ref_longest_repeat(s) == O.olib().orc_longest_repeat(s)


@pytest.mark.parametrize("mode", [0, 1])
def test_whole_path_vs_reference_binary(dataset, mode, tmp_path):
    from breakid_b200 import synth
    from test_golden import _format_calls
    d, hb, paths = dataset
    r = O.ref_run_binary(paths["bam"], str(tmp_path / "ref"), paths["nib"], fast=bool(mode))
    assert r.returncode == 0, r.stderr[-300:]
    nibs = [(synth.random_nib_bytes(l, d.cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(d.cfg.chrom_lens)]
    _, _, dist, cl = O.run(hb, nibs, mode=mode)
    exp = set()
    for ln in open(str(tmp_path / "ref") + "_fusion_all.txt").read().splitlines()[1:]:
        f = ln.split("\t")
        exp.add((f[0], f[1], f[2], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14]))
    assert _format_calls(cl, hb.target_names) == exp and len(exp) >= 5
    w_line = [l for l in open(str(tmp_path / "ref") + "_params.txt").read().splitlines() if l.startswith("w\t")][0]
    assert w_line == "w\t%g" % dist


def _binary_vs_oracle(cfg, tmp_path, mode=0, qual=None, genes_per_mb=4.0, min_calls=1):
    """the unmodified reference binary and the oracle on one synthetic dataset: same call records"""
    from breakid_b200 import api, bamio, synth
    from test_golden import _format_calls
    d = synth.generate(cfg)
    paths = bamio.write_dataset(str(tmp_path), d, random_qual=False, genes_per_mb=genes_per_mb)
    O.ref_index(paths["bam"])
    O.ref_install_refgene(paths["refgene"])
    r = O.ref_run_binary(paths["bam"], str(tmp_path / "ref"), paths["nib"], fast=bool(mode), qual=qual)
    assert r.returncode == 0, r.stderr[-300:]
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(cfg.chrom_lens)]
    _, _, dist, cl = O.run(hb, nibs, mode=mode, **({"qual": qual} if qual is not None else {}))
    exp = set()
    for ln in open(str(tmp_path / "ref") + "_fusion_all.txt").read().splitlines()[1:]:
        f = ln.split("\t")
        exp.add((f[0], f[1], f[2], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14]))
    assert _format_calls(cl, hb.target_names) == exp, (len(exp), len(cl))
    assert len(exp) >= min_calls
    return cl


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_hotspots_config5_shape_vs_reference_binary(tmp_path):
    """BASELINE.json configs[4] shape: hundreds of split reads per breakpoint, clips U(20,130), +-3 bp jitter -- pins the
    oracle's pairing / vote restatement (find_bp_pair, src/BreakID.cc:577-857) where it is quadratic in the reference"""
    from breakid_b200 import synth
    cfg = synth.config5(scale=0.004, split_per_sv=300)
    cl = _binary_vs_oracle(cfg, tmp_path, min_calls=4)
    assert int(cl["n_split_read"].max()) > 150


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("mode,qual", [(0, None), (1, 35)])
def test_tumour_config4_shape_vs_reference_binary(tmp_path, mode, qual):
    """BASELINE.json configs[3] shape: 100x coverage, translocations only, dense refGene; default and -fast -q 35"""
    from breakid_b200 import synth
    cfg = synth.SynthConfig(chrom_lens=[260000, 220000, 180000, 140000], coverage=100.0, n_tra=10, n_inv=0, n_dup=0, n_del=0,
                            span_per_sv=40, split_per_sv=20, seed=44, sv_jitter=1)
    _binary_vs_oracle(cfg, tmp_path, mode=mode, qual=qual, genes_per_mb=30.0, min_calls=8)


@pytest.mark.skipif(not O.have_ref() or not os.path.exists(O.REF_BIN + "_s"), reason="oracle/_ref/BreakID_ref_s not built")
@pytest.mark.parametrize("sd_mult,mode,qual", [(15, 0, 20), (15, 1, None), (1, 0, None), (3, 0, None)])
def test_sd_multiplier_oracle_vs_patched_reference(dataset, tmp_path, sd_mult, mode, qual):
    """the oracle's sd multiplier (the -s extension) against the reference binary whose literal 3 (src/BreakID.cc:103) is read
    from the environment -- with 3 that binary must reproduce the unmodified one"""
    from breakid_b200 import synth
    from test_golden import _format_calls
    d, hb, paths = dataset
    r = O.ref_run_binary(paths["bam"], str(tmp_path / "ref"), paths["nib"], fast=bool(mode), qual=qual, sd_mult=sd_mult)
    assert r.returncode == 0, r.stderr[-300:]
    nibs = [(synth.random_nib_bytes(l, d.cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(d.cfg.chrom_lens)]
    _, _, dist, cl = O.run(hb, nibs, mode=mode, sd_mult=sd_mult, **({"qual": qual} if qual is not None else {}))
    exp = set()
    for ln in open(str(tmp_path / "ref") + "_fusion_all.txt").read().splitlines()[1:]:
        f = ln.split("\t")
        exp.add((f[0], f[1], f[2], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14]))
    assert _format_calls(cl, hb.target_names) == exp
    w_line = [l for l in open(str(tmp_path / "ref") + "_params.txt").read().splitlines() if l.startswith("w\t")][0]
    assert w_line == "w\t%g" % dist
    if sd_mult == 3:
        r0 = O.ref_run_binary(paths["bam"], str(tmp_path / "ref0"), paths["nib"], fast=bool(mode), qual=qual)
        assert r0.returncode == 0
        for suffix in ("_fusion.txt", "_fusion_all.txt"):
            assert open(str(tmp_path / "ref") + suffix).read() == open(str(tmp_path / "ref0") + suffix).read()


def test_scan_wraps_genome_coordinates_like_the_reference(tmp_path):
    """a genome longer than 2^32 bases: combine_genome_chr_pos (src/util_bam.cc:57-68) adds target lengths in uint32, so the
    genome-wide coordinate of a pair far enough into the third target wraps; the oracle must wrap exactly like the reference
    (scan output compared byte for byte on a BAM whose targets are 2.1 + 2.1 + 0.4 Gb)"""
    from breakid_b200 import api, bamio, synth
    cfg = synth.SynthConfig(chrom_lens=[2_100_000_000, 2_100_000_000, 400_000_000], n_pairs=40000, n_tra=6, n_inv=3, n_dup=3, n_del=3, seed=7, sv_jitter=1)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    bam = str(tmp_path / "w.bam")
    bamio.write_bam(bam, d)
    synth.write_ref_names(str(tmp_path / "ref_names.txt"), 3)
    om, osd, _, _, _ = O.insert_stats(hb)
    w = O.dist(om, osd)
    pairs = O.scan(hb, 20, w)
    wrapped = pairs[(pairs["p2_tid"] == 2) & (pairs["p2_pos"] > 95_000_000)]
    assert len(wrapped) >= 3 and np.all(wrapped["p2_chr_pos"] == ((4_200_000_000 + wrapped["p2_pos"].astype(np.int64) - 1) % 2 ** 32))
    assert np.all(wrapped["p2_chr_pos"] < 400_000_000)                       # i.e. it really wrapped
    assert pairs.tobytes() == O.ref_scan(bam, 20, w, str(tmp_path)).tobytes()
