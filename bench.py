#!/usr/bin/env python
"""bench.py -- throughput of the BreakID hot path on B200 (BASELINE.json metric: read pairs/s
classified+clustered(+refined), decode excluded = `value`, host buffers + H2D/D2H included = `e2e`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--scale S] [--impl reference]

A "step" is one pass of the whole hot path (insert statistics, classify, mate join, bucket sort,
isolated-pair mask, AHC clustering, split-read refinement) over one synthetic record batch of
BASELINE.json configs[1] (hg19-shaped 3.1 Gb genome, 30x 2x150 bp, 2000 planted SVs, 1 % chimeric
noise; --scale shrinks it).  Inputs are larger than L2 (>= 20 GB at scale 1), so no L2 flush is
needed between iterations.  `--impl reference` times the unmodified reference CPU binary
(oracle/_ref/BreakID_ref) on a bounded sample of the same workload on the box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from breakid_b200 import api, synth  # noqa: E402

METRIC = "read_pairs_per_sec_classified_clustered_refined"
UNIT = "read pairs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def device_batch(d: synth.SynthData):
    """SynthData (cuda tensors) -> (api.Batch with device pointers, keep-alive dict).  The insert-size and span columns are
    handed over in their narrow forms (isize16 / span16, include/breakid_b200.h) whenever the whole batch fits them -- what
    a decoder does -- and stay narrow in HBM; tid stays wide (resident)."""
    keep = {}
    c = d.cols
    keep["flag"] = c["flag"].contiguous()
    keep["mapq"] = c["mapq"].contiguous()
    for k in ("tid", "pos"):
        keep[k] = c[k].contiguous()
    n = d.n
    flag = c["flag"].to(torch.int32)
    span = c["endpos"].to(torch.int64) - c["pos"].to(torch.int64)
    if n and int(span.min()) >= 0 and int(span.max()) <= 65535:
        keep["span16"] = span.to(torch.int16).contiguous()          # two's-complement wrap: same 16 bits as uint16
    else:
        keep["endpos"] = c["endpos"].contiguous()
    read = ((flag & 1) != 0) & ((flag & 2) != 0) & ((flag & (0x4 | 0x100 | 0x200 | 0x400)) == 0)
    iz = c["isize"][read]
    if n and (iz.numel() == 0 or (int(iz.min()) >= -32768 and int(iz.max()) <= 32767)):
        keep["isize16"] = c["isize"].clamp(-32768, 32767).to(torch.int16).contiguous()
    else:
        keep["isize"] = c["isize"].contiguous()
    del span, read, iz
    # sparse mate/name table: records that are not proper pairs or carry an SA tag (what a host decoder lists)
    in_x = (flag & 2) == 0
    in_x[d.sa_rec] = True
    del flag
    xr = torch.nonzero(in_x).flatten()
    keep["x_rec"] = xr.to(torch.int32).contiguous()
    keep["x_mtid"] = c["mtid"][xr].contiguous()
    keep["x_mpos"] = c["mpos"][xr].contiguous()
    keep["x_name_hash"] = synth.name_hash_ids(c["name_id"][xr]).contiguous()
    keep["sa_rec"] = d.sa_rec.to(torch.int32).contiguous()
    keep["cig_off"] = d.cig_off.to(torch.int32).contiguous()
    keep["cig_ops"] = d.cig_ops.to(torch.int32).contiguous()
    keep["sa_off"] = d.sa_off.to(torch.int32).contiguous()
    keep["sa_txt"] = d.sa_txt.contiguous()
    keep["oc_off"] = torch.zeros(d.sa_rec.numel() + 1, dtype=torch.int32, device=d.sa_rec.device)
    keep["oc_txt"] = torch.zeros(16, dtype=torch.uint8, device=d.sa_rec.device)
    b = api.Batch()
    b.n = d.n
    b.n_x = int(xr.numel())
    b.n_sa = int(d.sa_rec.numel())
    for k, v in keep.items():
        setattr(b, k, v.data_ptr())
    return b, keep


def host_batch_pinned(keep, n, n_sa, n_x):
    """pinned host copies of the device columns -> (api.Batch with host pointers, keep-alive, bytes).  Same column forms as
    the resident batch, plus tid as one run per target (records are coordinate sorted): 11 B/record instead of 19."""
    hk = {}
    nbytes = 0
    src = dict(keep)
    tid = keep["tid"]
    if n:
        start = torch.nonzero(torch.cat([torch.ones(1, dtype=torch.bool, device=tid.device), tid[1:] != tid[:-1]])).flatten()
        if start.numel() <= 65536:
            src["tid_run_start"] = start.to(torch.int32)
            src["tid_run_tid"] = tid[start].contiguous()
            del src["tid"]
    for k, v in src.items():
        h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        h.copy_(v)
        hk[k] = h
        nbytes += h.numel() * h.element_size()
    b = api.Batch()
    b.n = n
    b.n_sa = n_sa
    b.n_x = n_x
    for k, v in hk.items():
        setattr(b, k, v.data_ptr())
    if "tid_run_start" in hk:
        b.n_tid_runs = int(hk["tid_run_start"].numel())
    return b, hk, nbytes


def workload_cfg(scale):
    return synth.config2(scale=scale)


def measured_traffic(kernel, n_records):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/*.json written by
    tools/summarise_profiles.py), scaled per record to this run's launch; None when no capture is committed."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_top_kernels.json"))):
        try:
            k = json.load(open(p))["kernels"].get(kernel)
            if k and k.get("dram_bytes_per_record"):
                best = (k["dram_bytes_per_record"] * n_records, os.path.basename(p))
        except Exception:
            pass
    return best


def split_read_stress(local, dev, steps, scale):
    """BASELINE.json configs[4] (split-read refinement stress) as a second, short leg: 1000 x scale hotspots with
    5000 split reads each (1e7 x scale SA-tagged records), clip lengths U(20,130), +-3 bp breakpoint jitter.
    Reports split reads refined per second = SA-tagged records / (evidence + refine stage time)."""
    cfg = synth.config5(scale=scale)
    d = synth.generate(cfg, device=str(dev))
    names = [synth.chrom_name(t) for t in range(len(cfg.chrom_lens))]
    b_dev, keep = device_batch(d)
    n, n_sa = d.n, int(d.sa_rec.numel())
    del d
    torch.cuda.empty_cache()
    ctx = api.Context(cfg.chrom_lens, names, device=local)
    ev_ms = rf_ms = tot = 0.0
    for i in range(3 + steps):
        ctx.reset()
        ctx.push_device(b_dev)
        res = ctx.run()
        tm = ctx.timings()
        if i >= 3:
            ev_ms += tm["evidence"]; rf_ms += tm["refine"]; tot += tm["total"]
    ctx.close()
    del keep
    torch.cuda.empty_cache()
    return {"workload": "BASELINE.json configs[4] x scale %g: %d records, %d SA-tagged, %d hotspots called" % (scale, n, n_sa, int(res[3])),
            "value": n_sa / ((ev_ms + rf_ms) / steps * 1e-3), "unit": "split reads refined/s",
            "evidence_ms": ev_ms / steps, "refine_ms": rf_ms / steps, "whole_step_ms": tot / steps,
            "n_evidence": int(tm["n_evidence"]), "steps": steps}


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = workload_cfg(args.scale)
    cfg.seed = cfg.seed + 1000 * rank            # weak scaling: every rank holds one 30x sample-sized slice of the record stream
    t0 = time.time()
    d = synth.generate(cfg, device=str(dev))
    d.cols["name_id"] += rank * 1_000_000_000    # read names are unique over the whole job
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    names = [synth.chrom_name(t) for t in range(len(cfg.chrom_lens))]
    b_dev, keep = device_batch(d)
    n, n_sa = d.n, int(d.sa_rec.numel())
    del d
    torch.cuda.empty_cache()
    ctx = api.Context(cfg.chrom_lens, names, device=local)
    nibs = None
    if args.nib:
        for t, l in enumerate(cfg.chrom_lens):
            ctx.set_nib(t, synth.random_nib_bytes(l, cfg.seed * 1000 + t, device=str(dev)).cpu().numpy(), l)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    engine = None
    shard_timing = {}
    timing_on = [False]
    if world > 1:
        from breakid_b200.dist import GpuEngine, run_sharded
        engine = GpuEngine(ctx, dev)

    def run_path():
        if world == 1:
            return ctx.run()
        mean, sd, dd, out = run_sharded(engine, n, mode=0, timing=shard_timing if (os.environ.get("BKID_DEBUG_TIMING") and timing_on[0]) else None)
        return mean, sd, dd, len(out)

    def step_resident():
        ctx.reset()
        ctx.push_device(b_dev)
        return run_path()

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    stage = {k: 0.0 for k in api.TIMING_FIELDS_F}
    launches = 0
    with ClockSampler(local) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        dev_ms = 0.0
        timing_on[0] = True
        for _ in range(args.steps):
            res = step_resident()
            tm = ctx.timings()
            dev_ms += tm["total"]
            for k in stage:
                stage[k] += tm[k]
            launches += tm["kernel_launches"]
        timing_on[0] = False
        barrier()
        ev1.record()
        ev1.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dev_ms = ev0.elapsed_time(ev1)           # every library call is host-synchronous, so device span == wall span
    counts = {k: tm[k] for k in api.TIMING_FIELDS_I}
    ms_step = max(dev_ms, wall * 1e3) / args.steps       # device events and the wall clock must agree; report the slower
    ncall = res[3]
    # ---- e2e: host buffers, H2D + D2H inside the timed region ----
    b_host, hkeep, h2d_bytes = host_batch_pinned(keep, n, n_sa, int(b_dev.n_x))

    def step_e2e():
        ctx.reset()
        ctx.lib.bkid_push_batch(ctx.ctx, C.byref(b_host))
        r = run_path()
        out = ctx.fetch_clusters()
        return r, out
    del keep
    ctx.reset()
    torch.cuda.empty_cache()
    ctx.reserve(n, int(b_dev.n_x), n_sa, 2 * n_sa + 16, int(hkeep["sa_txt"].numel()) + 16, 16)
    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        r, out = step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    d2h_bytes = int(out.nbytes)

    # max over ranks
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_step, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
    pairs = n / 2.0 * world
    peak, peak_src = peaks()
    k1_ms = stage["classify"] / args.steps
    k1_bytes = 8.0 * n                                    # flag 2 + mapq 1 + isize 4 read, class 1 written
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None
    traffic = measured_traffic("k1_classify", n)
    sr_ms = (stage["evidence"] + stage["refine"]) / args.steps
    line = {
        "metric": METRIC, "value": pairs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/i64 integer + f64 (AHC distances)",
        "data": "synthetic (device-generated hg19-shaped genome, 30x 2x150bp, planted TRA/INV/DUP/DEL, 1% chimeric noise)",
        "config": {"workload": "BASELINE.json configs[1] x scale %g per GPU: %d records (%d read pairs), %d SA-tagged, AHC mode" % (args.scale, n, n // 2, n_sa),
                   "records_per_gpu": n, "input_bytes_per_gpu": h2d_bytes, "l2": "inputs larger than L2, no flush", "parallelism": ("single GPU" if world == 1 else "record-stream slices x%d; candidates all-to-all by name hash, pairs all-to-all by bucket owner, coverage/depth all-reduce (NCCL)" % world),
                   "calls": int(ncall)},
        "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                "note": "decode excluded: pinned host SoA batch (narrow isize16 / span16 / tid-run encodings where the batch fits) -> bkid_push_batch -> bkid_run -> bkid_fetch_clusters"},
        "gpu_launches": int(launches),
        "split_reads": {"value": (n_sa * world / (sr_ms * 1e-3)) if sr_ms > 0 else None, "unit": "split reads refined/s",
                        "note": "SA-tagged records of this workload / (evidence + refine stage time); the stress workload is in split_read_stress"},
        "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
        "counts": counts,
        "roofline": {"bound": "hbm", "kernel": "k1_classify", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "traffic": (traffic[0] if traffic else None), "traffic_source": (traffic[1] if traffic else None),
                     "peak_source": peak_src, "algorithmic_bytes_per_record": 8, "algorithmic_bytes_per_launch": k1_bytes},
        "clocks": clk.summary(),
        "gen_seconds": gen_s,
    }
    if rank == 0 and shard_timing:
        from breakid_b200 import dist as _d
        for k, v in _d._A2A_T.items():
            print("[a2a-timing] %-10s %8.3f ms total over all calls" % (k, v), file=sys.stderr)
        tot = sum(shard_timing.values())
        for k, v in shard_timing.items():
            print("[shard-timing] %-28s %8.3f ms/step" % (k, v / args.steps), file=sys.stderr)
    ctx.close()
    del hkeep, b_host
    torch.cuda.empty_cache()
    if world == 1 and not args.no_stress:
        line["split_read_stress"] = split_read_stress(local, dev, max(1, min(args.steps, 3)), args.stress_scale)
    if world == 1 and not args.no_decode:
        line["e2e_decode_included"] = decode_included(local, max(1, min(args.steps, 3)), min(args.scale, 1.0 / 256))
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_port(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def decode_included(local, steps, scale):
    """Decode INCLUDED, on the bounded sample the reference arm runs on (BASELINE.json configs[1] x `scale`, written as
    a real BGZF BAM with random qualities): compressed file bytes in pinned host memory -> bkid_push_bgzf (device
    inflate + record decode) -> bkid_run -> bkid_fetch_clusters, all inside the timed region."""
    from breakid_b200 import bamio
    cfg = workload_cfg(scale)
    d = synth.generate(cfg)
    tmp = tempfile.mkdtemp(prefix="bkid_dec_")
    bam = os.path.join(tmp, "reads.bam")
    bamio.write_bam(bam, d, random_qual=True)
    f = api.BgzfFile(bam)
    raw = torch.from_numpy(np.fromfile(bam, dtype=np.uint8)).pin_memory()
    ctx = api.Context(f.target_len, f.target_names, device=local)
    ts = []
    for i in range(2 + steps):
        ctx.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = ctx.push_bgzf(f, data_ptr=raw.data_ptr())
        ctx.run()
        out = ctx.fetch_clusters()
        dt = time.perf_counter() - t0
        if i >= 2:
            ts.append(dt)
    st = ctx.decode_stats()
    ctx.close()
    f.close()
    os.remove(bam)
    ms = float(np.mean(ts)) * 1e3
    return {"value": n / 2.0 / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(st["compressed_bytes"]), "d2h_bytes_per_step": int(out.nbytes),
            "sample": "configs[1] x scale %g: %d records, %d-byte BAM (%d BGZF blocks, %d bytes uncompressed)" % (scale, n, raw.numel(), st["n_blocks"], st["uncompressed_bytes"]),
            "decode_ms": st["total_ms"], "inflate_ms": st["inflate_ms"], "boundaries_ms": st["boundaries_ms"], "extract_ms": st["extract_ms"],
            "inflate_GBps_uncompressed": st["uncompressed_bytes"] / (st["inflate_ms"] * 1e-3) / 1e9 if st["inflate_ms"] > 0 else None,
            "note": "compressed BAM bytes in pinned host memory -> bkid_push_bgzf -> bkid_run -> bkid_fetch_clusters"}


def cpu_baseline_port(args):
    """oracle port (single thread) on a bounded sample of the same workload"""
    import oracle_py as O
    scale = min(args.scale, 1.0 / 64)
    cfg = workload_cfg(scale)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    t0 = time.perf_counter()
    O.run(hb, None, mode=0)
    dt = time.perf_counter() - t0
    return {"value": hb.n / 2.0 / dt, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": "configs[1] x scale %g: %d records, oracle/liboracle.so orc_run (AHC mode, decode excluded), host has %d cores" % (scale, hb.n, os.cpu_count())}


def run_reference(args):
    """the unmodified reference CPU binary on a bounded sample (rank 0 only)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle_py as O
    from breakid_b200 import bamio
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    scale = min(args.scale, 1.0 / 256)
    cfg = workload_cfg(scale)
    d = synth.generate(cfg)
    tmp = tempfile.mkdtemp(prefix="bkid_ref_")
    paths = bamio.write_dataset(tmp, d, random_qual=True)          # same bytes as the decode-included leg of our arm
    O.ref_index(paths["bam"])
    O.ref_install_refgene(paths["refgene"])
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = O.ref_run_binary(paths["bam"], os.path.join(tmp, "ref"), paths["nib"])
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            print(json.dumps({"impl": "reference", "unavailable": "BreakID_ref exited %d: %s" % (r.returncode, r.stderr[-200:])}))
            return
        if i >= args.warmup:
            times.append(dt)
    ms = float(np.mean(times)) * 1e3
    val = d.n / 2.0 / (ms * 1e-3)
    cores = 1
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/i64 integer + f64 (AHC distances)",
            "data": "synthetic", "config": {"workload": "BASELINE.json configs[1] x scale %g (bounded sample): %d records, BAM decode included (the reference has no other entry)" % (scale, d.n)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "oracle/_ref/BreakID_ref (single-threaded program; host has %d cores), default AHC mode, whole process wall clock" % os.cpu_count()},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the hg19-shaped 30x workload per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nib", action="store_true", help="also upload a random 4-bit genome so 41-mers are produced")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the configs[4] split-read stress leg")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode-included leg (BAM file -> device inflate -> hot path)")
    ap.add_argument("--stress-scale", type=float, default=1.0, help="fraction of the 1e7-SA-record stress workload")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
