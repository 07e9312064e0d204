#!/usr/bin/env python
"""Decode-included throughput on a bounded sample: BAM file bytes (pinned host memory) -> bkid_push_bgzf (device
inflate + record decode) -> bkid_run -> bkid_fetch_clusters, next to the host decoder and the two driver binaries.
   python tests/decode_bench.py [--scale 0.00390625] [--reps 5] [--replicate 1]
--replicate N repeats the record blocks N times inside one file (decode-throughput measurement at multi-GB scale;
such a file is not coordinate sorted, so only the decode is timed on it)."""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from breakid_b200 import api, bamio, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0 / 256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--replicate", type=int, default=1)
    ap.add_argument("--drivers", action="store_true", help="also time the BreakID driver binaries (ours and, if built, the reference)")
    a = ap.parse_args()
    cfg = synth.config2(scale=a.scale)
    d = synth.generate(cfg)
    tmp = tempfile.mkdtemp(prefix="bkid_dec_")
    t0 = time.time()
    paths = bamio.write_dataset(tmp, d, random_qual=True)
    t_write = time.time() - t0
    bam = paths["bam"]
    out = {"records": d.n, "bam_bytes": os.path.getsize(bam), "write_s": t_write}
    f = api.BgzfFile(bam)
    raw = np.fromfile(bam, dtype=np.uint8)
    bt = f.block_table()
    nblk = f.n_blocks
    if a.replicate > 1:
        # one valid (unsorted) BAM whose record section is repeated: inflate everything, repeat the records, recompress
        import zlib
        from concurrent.futures import ThreadPoolExecutor
        rawb = raw.tobytes()
        payload = b"".join(zlib.decompress(rawb[int(b["payload_off"]):int(b["payload_off"]) + int(b["payload_len"])], -15) for b in bt if b["usize"])
        payload = payload[:f.first_record] + payload[f.first_record:] * a.replicate
        chunks = [payload[o:o + 0xff00] for o in range(0, len(payload), 0xff00)]
        with ThreadPoolExecutor(os.cpu_count()) as pool:
            comp = list(pool.map(lambda c: bamio._bgzf_block(c, 1), chunks))
        p2 = os.path.join(tmp, "rep.bam")
        with open(p2, "wb") as fo:
            for c in comp:
                fo.write(c)
            fo.write(bamio._BGZF_EOF)
        del payload, chunks, comp, rawb
        raw = np.fromfile(p2, dtype=np.uint8)
        f.close()
        f = api.BgzfFile(p2)
        bt = f.block_table(); nblk = f.n_blocks
        out["replicated_bam_bytes"] = int(raw.size)
    pinned = torch.from_numpy(raw).pin_memory()
    ctx = api.Context(f.target_len, f.target_names, device=0)
    times = []
    for i in range(a.reps + 2):
        ctx.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = ctx.push_bgzf(f, data_ptr=pinned.data_ptr())
        t1 = time.perf_counter()
        if a.replicate == 1:
            r = ctx.run(); cl = ctx.fetch_clusters()
        t2 = time.perf_counter()
        if i >= 2:
            times.append((t1 - t0, t2 - t0))
    st = ctx.decode_stats()
    dec = float(np.median([t[0] for t in times])); tot = float(np.median([t[1] for t in times]))
    out.update({"decode_s": dec, "decode_plus_path_s": tot, "n_decoded": n, "stats": st,
                "uncompressed_GBps": st["uncompressed_bytes"] / dec / 1e9, "compressed_GBps": st["compressed_bytes"] / dec / 1e9,
                "inflate_kernel_GBps": st["uncompressed_bytes"] / (st["inflate_ms"] * 1e-3) / 1e9,
                "read_pairs_per_s_decode_included": n / 2 / tot})
    ctx.close()
    if a.replicate == 1:
        t0 = time.perf_counter(); hb = api.HostBatch.from_bam(bam, threads=os.cpu_count()); out["host_decoder_s"] = time.perf_counter() - t0
        out["host_decoder_threads"] = os.cpu_count()
    if a.drivers and a.replicate == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as O
        O.ref_index(bam)
        env = dict(os.environ, BREAKID_INSTALLDIR=tmp)
        drv = os.path.join(ROOT, "breakid_b200", "host", "BreakID")
        for name, extra in (("driver_device_decode_s", {}), ("driver_host_decode_s", {"BKID_HOST_DECODE": "1"})):
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                subprocess.run([drv, "-i", bam, "-o", os.path.join(tmp, name), "-n", paths["nib"]], env=dict(env, **extra), check=True, capture_output=True)
                ts.append(time.perf_counter() - t0)
            out[name] = min(ts)
        O.ref_install_refgene(paths["refgene"])
        t0 = time.perf_counter(); r = O.ref_run_binary(bam, os.path.join(tmp, "ref"), paths["nib"]); out["reference_binary_s"] = time.perf_counter() - t0
        a_ = open(os.path.join(tmp, "driver_device_decode_s_fusion.txt")).read(); b_ = open(os.path.join(tmp, "ref_fusion.txt")).read()
        out["call_files_identical"] = (a_ == b_)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
