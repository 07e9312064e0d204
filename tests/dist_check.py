#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a multi-GPU box):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
Every rank takes a contiguous slice of one synthetic dataset; the sharded result must be byte-identical to the
CPU oracle on the unsplit input (rank 0 checks and prints)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from breakid_b200 import api, synth
    from breakid_b200.dist import GpuEngine, LibraryDist, run_sharded
    from test_multi_gloo import _slice
    import oracle_py as O
    ok = True
    for mode in (0, 1):
        cfg = synth.SynthConfig(chrom_lens=[400000, 300000, 250000, 200000], n_tra=6, n_inv=3, n_dup=3, n_del=3, seed=41, sv_jitter=1)
        d = synth.generate(cfg)
        hb = api.HostBatch.from_synth(d)
        nibs = [(synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(cfg.chrom_lens)]
        cuts = [hb.n * i // world for i in range(world + 1)]
        part = _slice(hb, cuts[rank], cuts[rank + 1])
        ctx = api.Context(hb.target_len, hb.target_names, device=local, fast=mode)
        ctx.push(part)
        for t, (p, l) in enumerate(nibs):
            ctx.set_nib(t, p, l)
        mean, sd, dd, out = run_sharded(GpuEngine(ctx, dev), part.n, mode=mode)
        # the same through the library's own exchanges (NCCL calls from C++, bkid_dist_run): every rank checks its own copy
        ctx2 = api.Context(hb.target_len, hb.target_names, device=local, fast=mode)
        ctx2.push(part)
        for t, (p, l) in enumerate(nibs):
            ctx2.set_nib(t, p, l)
        ld = LibraryDist(ctx2, dev)
        lm, ls, ldd, ln, _ = ld.run(mode)
        lout = ctx2.fetch_clusters()
        m, s, d0, exp = O.run(hb, nibs, mode=mode)
        lib_same = torch.tensor([1 if ((lm, ls, ldd) == (m, s, d0) and lout.tobytes() == exp.tobytes()) else 0], device=dev)
        dist.all_reduce(lib_same, op=dist.ReduceOp.MIN)
        ld.close(); ctx2.close()
        if rank == 0:
            same = (mean, sd, dd) == (m, s, d0) and out.tobytes() == exp.tobytes()
            print("mode %d world %d: %d calls, identical to oracle: %s; library exchanges (NCCL in C++) identical on every rank: %s" % (mode, world, len(out), same, bool(int(lib_same[0]))), flush=True)
            ok = ok and same and bool(int(lib_same[0])) and len(exp) >= 10
        ctx.close()
    # exclude intervals on every rank's context (BASELINE.json configs[2] shape: exclude-BED + genomic-bin sharding)
    from test_gpu_parity import _exclude_intervals, _prefilter
    iv = _exclude_intervals(d, np.random.RandomState(8))
    ctx = api.Context(hb.target_len, hb.target_names, device=local)
    ctx.push(part)
    ctx.set_exclude(iv[:, 0], iv[:, 1], iv[:, 2])
    mean, sd, dd, out = run_sharded(GpuEngine(ctx, dev), part.n, mode=0)
    if rank == 0:
        tid, pos = hb.cols["tid"].astype(np.int64), hb.cols["pos"].astype(np.int64)
        keep = np.ones(hb.n, bool)
        for t, b, e in iv:
            keep &= ~((tid == t) & (pos >= max(b, 0)) & (pos < e))
        m, s, d0, exp = O.run(_prefilter(hb, keep), None, mode=0)
        same = (mean, sd, dd) == (m, s, d0) and out.tobytes() == exp.tobytes()
        print("exclude world %d: %d calls, identical to oracle on the pre-filtered input: %s" % (world, len(out), same), flush=True)
        ok = ok and same
    ctx.close()
    # the full configs[2] shape through the library's exchanges: exclude intervals + -q 20 -s 15
    ctx = api.Context(hb.target_len, hb.target_names, device=local, qual=20, sd_mult=15)
    ctx.push(part)
    ctx.set_exclude(iv[:, 0], iv[:, 1], iv[:, 2])
    ld = LibraryDist(ctx, dev)
    lm, ls, ldd, ln, _ = ld.run(0)
    lout = ctx.fetch_clusters()
    ld.close(); ctx.close()
    if rank == 0:
        m, s, d0, exp = O.run(_prefilter(hb, keep), None, qual=20, mode=0, sd_mult=15)
        same = (lm, ls, ldd) == (m, s, d0) and lout.tobytes() == exp.tobytes()
        print("exclude + -q 20 -s 15 world %d (library exchanges): %d calls, identical to oracle on the pre-filtered input: %s" % (world, len(lout), same), flush=True)
        ok = ok and same
    # ingest sharded too: every rank inflates and decodes its own BGZF block range of one BAM file on its GPU
    import tempfile
    from breakid_b200 import bamio
    tmpd = [tempfile.mkdtemp(prefix="bkid_dist_") if rank == 0 else None]
    dist.broadcast_object_list(tmpd, src=0)
    bam = os.path.join(tmpd[0], "reads.bam")
    if rank == 0:
        bamio.write_bam(bam, d)
    dist.barrier()
    f = api.BgzfFile(bam)
    cuts = [f.n_blocks * i // world for i in range(world + 1)]
    ctx = api.Context(f.target_len, f.target_names, device=local)
    n_loc, a, b = ctx.push_bgzf_range(f, cuts[rank], cuts[rank + 1])
    marks = torch.tensor([a, b, n_loc], dtype=torch.int64, device=dev)
    allm = [torch.zeros_like(marks) for _ in range(world)]
    dist.all_gather(allm, marks)
    stitched = all(int(allm[i][1]) == int(allm[i + 1][0]) for i in range(world - 1)) and sum(int(m[2]) for m in allm) == hb.n
    for t, (p, l) in enumerate(nibs):
        ctx.set_nib(t, p, l)
    mean, sd, dd, out = run_sharded(GpuEngine(ctx, dev), n_loc, mode=0)
    if rank == 0:
        m, s, d0, exp = O.run(hb, nibs, mode=0)
        same = stitched and (mean, sd, dd) == (m, s, d0) and out.tobytes() == exp.tobytes()
        print("sharded decode world %d: ranges stitch: %s, %d calls, identical to oracle: %s" % (world, stitched, len(out), same), flush=True)
        ok = ok and same
    ctx.close()
    f.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
