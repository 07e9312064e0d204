"""CPU tests (no GPU, no /root/reference): the oracle against the committed golden fixtures that
tests/golden/make_golden.py produced with the REAL reference (binary + function wrappers)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def mini():
    from breakid_b200 import api
    z = np.load(os.path.join(G, "mini.npz"))
    hb = api.HostBatch({k[4:]: z[k] for k in z.files if k.startswith("col_")}, z["name_hash"],
                       {k[5:]: z[k] for k in z.files if k.startswith("side_")}, z["target_len"], ["chr1", "chr2"])
    nibs = [(z["nib0"], int(z["target_len"][0])), (z["nib1"], int(z["target_len"][1]))]
    return hb, nibs


@pytest.fixture(scope="module")
def ops():
    return np.load(os.path.join(G, "ops.npz"))


def test_insert_stats_and_scan_match_reference(mini):
    import oracle_py as O
    hb, _ = mini
    st = json.load(open(os.path.join(G, "mini_ref_stats.json")))
    m, s, _, _, _ = O.insert_stats(hb)
    assert (m, s) == (st["mean"], st["sd"])
    assert O.dist(m, s) == st["dist"]
    exp = np.load(os.path.join(G, "mini_ref_scan.npy"))
    got = O.scan(hb, 20, st["dist"])
    assert got.tobytes() == exp.tobytes()


def _format_calls(cl, names):
    from breakid_b200.api import FUSION_TYPES
    rows = set()
    for c in cl:
        rows.add((FUSION_TYPES[c["fusion_type"]], "%s:%d" % (names[c["p1_tid"]], c["p1_exact_pos"]), "%s:%d" % (names[c["p2_tid"]], c["p2_exact_pos"]),
                  str(int(c["n_discordant_pair"])), str(int(c["n_split_read"])), "%g" % c["p1_bp_depth"], "%g" % c["p2_bp_depth"],
                  "%g" % c["p1_alle_freq"], "%g" % c["p2_alle_freq"], c["p1_rpt"].decode(), c["p2_rpt"].decode()))
    return rows


@pytest.mark.parametrize("mode", [0, 1])
def test_whole_path_matches_reference_call_file(mini, mode):
    """every numeric column of the reference's _fusion_all.txt (types, breakpoints, N_DRP, N_SR, depths,
    AFs, 41-mers) is reproduced by the oracle on the same records"""
    import oracle_py as O
    hb, nibs = mini
    _, _, _, cl = O.run(hb, nibs, mode=mode)
    got = _format_calls(cl, hb.target_names)
    lines = open(os.path.join(G, "mini_ref_%s_fusion_all.txt" % ("fast" if mode else "ahc"))).read().splitlines()[1:]
    exp = set()
    for ln in lines:
        f = ln.split("\t")
        exp.add((f[0], f[1], f[2], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14]))
    assert got == exp and len(exp) >= 5


def test_std_sort_replay(ops):
    import oracle_py as O
    for t in range(40):
        key = ops["sort_key_%d" % t]
        assert np.array_equal(O.sort_perm(key), ops["sort_perm_%d" % t])
        assert np.array_equal(O.sort_perm(key, model=True), ops["sort_perm_%d" % t])


def test_mask_and_clustering(ops):
    import oracle_py as O
    n_ahc = 0
    for t in range(60):
        x, y, w = ops["mask_x_%d" % t], ops["mask_y_%d" % t], float(ops["mask_w_%d" % t][0])
        keep = O.remove_isolated(x, y, w)
        assert np.array_equal(keep, ops["mask_out_%d" % t]), t
        if "ahc_idx_%d" % t in ops.files:
            xs = np.ascontiguousarray(x[keep]); ys = np.ascontiguousarray(y[keep])
            for name, mode, model in (("ahc", 0, False), ("ahc", 0, True), ("fast", 1, False)):
                i, c, r = O.cluster(mode, xs, ys, w, model=model)
                assert np.array_equal(i, ops["%s_idx_%d" % (name, t)]), (name, t)
                assert np.array_equal(c, ops["%s_cl_%d" % (name, t)]), (name, t)
                assert r == int(ops["%s_roots_%d" % (name, t)][0])
            n_ahc += 1
    assert n_ahc > 20


def test_ahc_merge_trees(ops):
    """whole merge tree (children of every node in creation order) on tie-heavy lattices, literal
    restatement and the device-formulation model"""
    import ctypes as C
    import oracle_py as O
    for t in range(60):
        x, y, thr = ops["tree_x_%d" % t], ops["tree_y_%d" % t], int(ops["tree_thr_%d" % t][0])
        n = len(x)
        for f in (O.olib().orc_ahc_tree, O.olib().orc_model_ahc_tree):
            r = np.zeros(2 * n + 1, np.int32); a = np.zeros(2 * n + 1, np.int32); b = np.zeros(2 * n + 1, np.int32)
            nn = f(n, np.ascontiguousarray(x), np.ascontiguousarray(y), thr, r, a, b)
            assert nn == len(ops["tree_root_%d" % t])
            assert np.array_equal(r[:nn], ops["tree_root_%d" % t]) and np.array_equal(a[:nn], ops["tree_a_%d" % t]) and np.array_equal(b[:nn], ops["tree_b_%d" % t]), t


def test_complementary_cigars():
    import oracle_py as O
    cig = json.load(open(os.path.join(G, "cigars.json")))
    bad = [c for c in cig if O.olib().orc_is_complementary(c[0].encode(), c[1].encode(), 10) != c[2]]
    assert not bad, bad[:5]
    assert sum(c[2] for c in cig) > 100


def test_depth_exhausted_keys_model_equals_std_sort():
    """the level-synchronous model of std::sort (what the GPU kernels implement) on keys that exhaust the introsort depth
    budget on ~4 000 element segments (real bucket, tests/golden/make_depth_exhausted_keys.py)"""
    import oracle_py as O
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth_exhausted_sort_keys.npz"))
    for name in ("sort2", "sort3"):
        key = g[name]
        assert np.array_equal(O.sort_perm(key), O.sort_perm(key, model=True)), name
