#!/usr/bin/env python
"""Where does a rank of the N-GPU bench spend its cluster stage?  Runs the in-library sharded path with the N ranks as host
threads on ONE GPU (device-copy communicator) on exactly the slices bench.py --gpus N gives the ranks, and prints every
rank's stage timers and counts (CUDA events on the rank's own stream; the ranks share the SMs, so absolute times are
inflated -- the imbalance between ranks is what this shows).
   python tools/dist_probe_one_gpu.py [world] [scale]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from breakid_b200 import api, synth  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    dev = torch.device("cuda", 0)
    cfg = bench.workload_cfg(scale)
    names = [synth.chrom_name(t) for t in range(len(cfg.chrom_lens))]
    ctxs, keeps = [], []
    for r in range(world):
        d = bench.rank_slice(cfg, r, world, dev)
        b_dev, keep = bench.device_batch(d)
        del d
        torch.cuda.empty_cache()
        c = api.Context(cfg.chrom_lens, names, device=0)
        c.push_device(b_dev)
        ctxs.append(c); keeps.append((b_dev, keep))
    for it in range(2):
        if it:
            for c, (b, _) in zip(ctxs, keeps):
                c.reset(); c.push_device(b)
        res = api.dist_run_local(ctxs, 0)
    for r, c in enumerate(ctxs):
        tm = c.timings()
        print(json.dumps({"rank": r, "n_called": res[r][3], "stage_ms": {k: round(tm[k], 3) for k in ("classify", "mask", "cluster", "summarize", "evidence", "refine")},
                          "counts": {k: tm[k] for k in api.TIMING_FIELDS_I}}), flush=True)
    # replay every rank's mask + cluster stage ALONE on the GPU (the pairs the rank owned after the all-to-all), per-kernel times
    import ctypes as C
    import numpy as np
    dist_thr = res[0][2]
    owned = []
    for c in ctxs:
        p = c.fetch_pairs(0).copy()
        nb = C.c_int64()
        c._chk(c.lib.bkid_fetch_bucket_ranks(c.ctx, None, 0, C.byref(nb)))
        ranks = np.zeros(max(1, nb.value), np.int32)
        c._chk(c.lib.bkid_fetch_bucket_ranks(c.ctx, ranks.ctypes.data, ranks.shape[0], C.byref(nb)))
        p["bucket"] = ranks[p["bucket"]]
        owned.append(p)
        c.close()
    del keeps
    torch.cuda.empty_cache()
    for r, p in enumerate(owned):
        t = torch.from_numpy(p.view(np.uint8).reshape(-1)).to(dev)
        c2 = api.Context(cfg.chrom_lens, names, device=0)
        for it in range(3):
            if it == 2:
                api.profile_kernels(True); api.profile_report()
            c2._chk(c2.lib.bkid_shard_set_pairs(c2.ctx, C.c_void_p(t.data_ptr()), int(p.shape[0])))
            c2.cluster(dist_thr, 0)
        rep = api.profile_report()
        api.profile_kernels(False)
        tm = c2.timings()
        sizes = np.bincount(p["bucket"])
        tot = sum(v[1] for v in rep.values())
        print(json.dumps({"replay_rank": r, "pairs": int(p.shape[0]), "buckets": int((sizes > 0).sum()), "largest_buckets": sorted(sizes.tolist())[-3:],
                          "mask_ms": round(tm["mask"], 3), "cluster_ms": round(tm["cluster"], 3),
                          "kernels": {nm: [v[0], round(v[1], 3)] for nm, v in sorted(rep.items(), key=lambda kv: -kv[1][1]) if v[1] > 0.04 * tot}}), flush=True)
        if tm["mask"] > 6.0 and os.environ.get("BKID_PROBE_DUMP"):
            # which bucket is it?  time the mask of every large bucket on its own and keep the slowest one's coordinates
            import time
            worst = (0.0, None)
            for bk_ in np.nonzero(sizes > 20000)[0]:
                q = p[p["bucket"] == bk_]
                q = q[np.argsort(q["orig"], kind="stable")]
                x = q["p1_chr_pos"].copy(); y = q["p2_chr_pos"].copy()
                c2.op_remove_isolated(x, y, dist_thr)
                api.profile_kernels(True); api.profile_report()
                out = c2.op_remove_isolated(x, y, dist_thr)
                rep2 = api.profile_report(); api.profile_kernels(False)
                dt = sum(v[1] for v in rep2.values())
                lv = rep2.get("is_level", (0, 0.0))
                print(json.dumps({"rank": r, "bucket": int(bk_), "pairs": int(len(q)), "kept": int(len(out)), "kernel_ms": round(dt, 3), "is_level": [lv[0], round(lv[1], 3)],
                                  "tids": [int(q["p1_tid"][0]), int(q["p2_tid"][0])]}), flush=True)
                if dt > worst[0]:
                    worst = (dt, (x, y, int(bk_)))
            if worst[1] is not None:
                np.savez_compressed(os.path.join(ROOT, "gpurun_out", "slow_bucket_rank%d.npz" % r), x=worst[1][0], y=worst[1][1], bucket=worst[1][2], w=dist_thr)
        c2.close()


if __name__ == "__main__":
    main()
