#!/usr/bin/env python
"""Writes tests/golden/depth_exhausted_sort_keys.npz: the keys of the second and third std::sort of remove_isolated_pairs
(src/BreakID.cc:1278,1282) on one same-chromosome bucket of the 8-GPU bench workload (bench.rank_slice world 8, rank 3,
bucket chr3_chr3, 144 777 pairs, dumped on a B200 box by `BKID_PROBE_DUMP=1 python tools/dist_probe_one_gpu.py 8 1.0` ->
gpurun_out/slow_bucket_rank3.npz).  Near-sorted input drives libstdc++'s median-of-3 introsort to its depth limit here:
segments of 3 901 and 4 459 elements end in the heapsort fallback -- the case is_heap (bkid_core.cu) exists for.
   python tests/golden/make_depth_exhausted_keys.py gpurun_out/slow_bucket_rank3.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py as O  # noqa: E402


def mask_pass(idx, p1, p2, dist):
    """mask_pairs_chr_pos (src/BreakID.cc:1813-1877) on the current order, vectorised"""
    if len(idx) <= 2:
        return idx[:0]
    a1 = p1[idx].astype(np.int64); a2 = p2[idx].astype(np.int64)

    def gap(a, b):
        d = (a - b) & 0xffffffff
        return np.abs(np.where(d >= 2 ** 31, d - 2 ** 32, d))
    head = []
    if not (gap(a1[1:2], a1[2:3])[0] > dist or gap(a2[1:2], a2[2:3])[0] > dist):
        head.append(idx[1])
    lx = np.minimum(gap(a1[:-2], a1[1:-1]), gap(a1[2:], a1[1:-1]))
    ly = np.minimum(gap(a2[:-2], a2[1:-1]), gap(a2[2:], a2[1:-1]))
    return np.concatenate([np.array(head, dtype=idx.dtype), idx[1:-1][~((lx > dist) | (ly > dist))]])


def main():
    d = np.load(sys.argv[1])
    x, y, w = d["x"], d["y"], int(float(d["w"]))
    v = mask_pass(O.sort_perm(x).astype(np.int64), x, y, w)
    k2 = y[v].astype(np.uint32)
    v2 = mask_pass(v[O.sort_perm(k2)], x, y, w)
    k3 = x[v2].astype(np.uint32)
    assert len(v2) == len(O.remove_isolated(x, y, float(d["w"])))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "depth_exhausted_sort_keys.npz"), sort2=k2, sort3=k3)
    print(len(k2), len(k3))


if __name__ == "__main__":
    main()
