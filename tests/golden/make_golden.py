#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the REAL reference (needs /root/reference and
`make -C oracle ref`).  Run from the repo root:  python tests/golden/make_golden.py

Fixtures (all small):
  mini.npz                         record batch (SoA) + nib payloads of a 2-chromosome 30x dataset
  mini_ref_{ahc,fast}_fusion_all.txt, mini_ref_params.txt     reference binary output on that dataset's BAM
  mini_ref_scan.npy                ref_scan() pair table
  ops.npz                          std::sort permutations, isolated-pair masks, AHC / -fast clusterings and
                                   merge trees produced by the reference functions on seeded tie-heavy inputs
  cigars.json                      is_complementary_cigar known answers
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C  # noqa: E402

import oracle_py as O  # noqa: E402
from breakid_b200 import api, bamio, synth  # noqa: E402
from conftest import lattice_points  # noqa: E402


def mini_cfg():
    return synth.SynthConfig(chrom_lens=[120000, 90000], n_tra=2, n_inv=1, n_dup=1, n_del=1, seed=33, min_sv_sep=5000, sv_jitter=1, chimeric_frac=0.02)


def main():
    assert O.have_ref(), "build oracle/_ref first"
    cfg = mini_cfg()
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    nibs = [synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy() for t, l in enumerate(cfg.chrom_lens)]
    np.savez_compressed(os.path.join(HERE, "mini.npz"), **{"col_" + k: v for k, v in hb.cols.items()}, name_hash=hb.name_hash,
                        **{"side_" + k: v for k, v in hb.side.items()}, target_len=hb.target_len, nib0=nibs[0], nib1=nibs[1])
    tmp = tempfile.mkdtemp(prefix="golden_")
    paths = bamio.write_dataset(tmp, d, random_qual=False, genes_per_mb=40.0)
    O.ref_index(paths["bam"])
    O.ref_install_refgene(paths["refgene"])
    import shutil
    shutil.copy(paths["refgene"], os.path.join(HERE, "mini_refGene.txt"))
    for mode in ("ahc", "fast"):
        r = O.ref_run_binary(paths["bam"], os.path.join(tmp, mode), paths["nib"], fast=(mode == "fast"))
        assert r.returncode == 0, r.stderr
        shutil.copy(os.path.join(tmp, mode + "_fusion_all.txt"), os.path.join(HERE, "mini_ref_%s_fusion_all.txt" % mode))
        shutil.copy(os.path.join(tmp, mode + "_fusion.txt"), os.path.join(HERE, "mini_ref_%s_fusion.txt" % mode))
    txt = open(os.path.join(tmp, "ahc_params.txt")).read().replace(tmp, "TMP")
    open(os.path.join(HERE, "mini_ref_params.txt"), "w").write(txt)
    m, s = C.c_double(), C.c_double()
    with O.quiet():
        O.rlib().ref_insert_stats(paths["bam"].encode(), C.byref(m), C.byref(s))
    w = O.dist(m.value, s.value)
    np.save(os.path.join(HERE, "mini_ref_scan.npy"), O.ref_scan(paths["bam"], 20, w, paths["nib"]))
    json.dump({"mean": m.value, "sd": s.value, "dist": w}, open(os.path.join(HERE, "mini_ref_stats.json"), "w"))

    # ---- function-level vectors ----
    R = O.rlib()
    rng = np.random.RandomState(77)
    ops = {}
    for t in range(40):
        n = int(rng.choice([3, 17, 40, 100, 300, 1000]))
        key = (rng.randint(0, max(2, n // 3), n) if t % 2 else rng.randint(0, 2 ** 32, n)).astype(np.uint32)
        perm = np.zeros(n, np.uint32)
        R.ref_std_sort_perm(n, key, t % 2, perm)      # 0 = cmp_p1, 1 = cmp_p2 (unsigned keys)
        ops["sort_key_%d" % t] = key; ops["sort_perm_%d" % t] = perm
    for t in range(60):
        n = int(rng.randint(3, 120))
        x, y = lattice_points(rng, n, t % 4)
        w_ = float(rng.choice([120.7, 260.2, 99.0, 1243.15]))
        out = np.zeros(n + 2, np.uint32)
        with O.quiet():
            k = R.ref_remove_isolated(n, x, y, w_, out)
        ops["mask_x_%d" % t] = x; ops["mask_y_%d" % t] = y; ops["mask_w_%d" % t] = np.array([w_]); ops["mask_out_%d" % t] = out[:k].copy()
        xs = np.ascontiguousarray(x[out[:k]]); ys = np.ascontiguousarray(y[out[:k]])
        if k >= 2:
            for name, f in (("ahc", R.ref_cluster_ahc), ("fast", R.ref_cluster_fast)):
                oi = np.zeros(k + 2, np.uint32); oc = np.zeros(k + 2, np.int32); r = C.c_int()
                with O.quiet():
                    m_ = f(k, xs, ys, w_, oi, oc, C.byref(r), 0)
                ops["%s_idx_%d" % (name, t)] = oi[:m_].copy(); ops["%s_cl_%d" % (name, t)] = oc[:m_].copy(); ops["%s_roots_%d" % (name, t)] = np.array([r.value])
    for t in range(60):
        n = int(rng.randint(2, 60)); kind = t % 3
        if kind == 0:
            x = rng.randint(0, 5, n) * 50.; y = rng.randint(0, 5, n) * 50.
        elif kind == 1:
            k = max(1, n // 6); a = rng.randint(0, k, n); x = a * 1000. + rng.randint(0, 4, n) * 50; y = (a % 2) * 800. + rng.randint(0, 4, n) * 50
        else:
            k = max(1, n // 5); a = rng.randint(0, k, n); x = a * 500. + rng.randint(0, 3, n) * 50; y = rng.randint(0, 3, n) * 50. + (a % 3) * 400
        thr = int(rng.choice([60, 120, 200, 260]))
        r_ = np.zeros(2 * n + 1, np.int32); a_ = np.zeros(2 * n + 1, np.int32); b_ = np.zeros(2 * n + 1, np.int32)
        with O.quiet():
            nn = R.ref_ahc_tree(n, np.ascontiguousarray(x), np.ascontiguousarray(y), thr, r_, a_, b_)
        ops["tree_x_%d" % t] = x; ops["tree_y_%d" % t] = y; ops["tree_thr_%d" % t] = np.array([thr])
        ops["tree_root_%d" % t] = r_[:nn].copy(); ops["tree_a_%d" % t] = a_[:nn].copy(); ops["tree_b_%d" % t] = b_[:nn].copy()
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)

    cig = []
    ops_ = "MSIDHN=X"
    for t in range(3000):
        def rc():
            k = int(rng.choice([1, 2, 2, 2, 3]))
            return "".join("%d%s" % (rng.randint(0, 160), ops_[rng.randint(0, 2) if rng.rand() < 0.85 else rng.randint(0, len(ops_))]) for _ in range(k))
        a, b = rc(), rc()
        if rng.rand() < 0.4:       # make complementary-looking pairs common
            k = int(rng.randint(1, 149)); j = int(rng.randint(-12, 13))
            a = "%dM%dS" % (k, 150 - k) if rng.rand() < 0.5 else "%dS%dM" % (150 - k, k)
            b = "%dS%dM" % (k + j, 150 - k - j) if rng.rand() < 0.5 else "%dM%dS" % (150 - k - j, k + j)
        cig.append([a, b, int(R.ref_is_complementary(a.encode(), b.encode(), 10))])
    json.dump(cig, open(os.path.join(HERE, "cigars.json"), "w"))
    print("golden fixtures written; complementary positives:", sum(c[2] for c in cig))


if __name__ == "__main__":
    main()
