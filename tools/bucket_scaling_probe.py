#!/usr/bin/env python
"""How does the resident step behave when chr-pair buckets grow (what a rank sees at N GPUs, where the merged stream has
N x the pairs per bucket and a rank owns 1/N of the buckets)?  Same record count and pair count as configs[1], fewer and
longer chromosomes: 24 (hg19), 6, 3.  Prints stage times and the kernels above 2 % of the step (CUDA events per launch).
   python tools/bucket_scaling_probe.py [n_chrom ...]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from breakid_b200 import api, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    base = synth.config2()
    G = sum(base.chrom_lens)
    for k in [int(a) for a in sys.argv[1:]] or [24, 6, 3]:
        cfg = base if k == 24 else synth.SynthConfig(**{**base.__dict__, "chrom_lens": [G // k] * k})
        d = synth.generate(cfg, device=str(dev))
        names = [synth.chrom_name(t) for t in range(len(cfg.chrom_lens))]
        b_dev, keep = bench.device_batch(d)
        n = d.n
        del d
        torch.cuda.empty_cache()
        ctx = api.Context(cfg.chrom_lens, names, device=0)
        for i in range(3):
            ctx.reset(); ctx.push_device(b_dev); ctx.run()
        tm = ctx.timings()
        api.profile_kernels(True); api.profile_report()
        ctx.reset(); ctx.push_device(b_dev); ctx.run()
        rep = api.profile_report()
        api.profile_kernels(False)
        tot = sum(v[1] for v in rep.values())
        print(json.dumps({"n_chrom": k, "records": n, "stage_ms": {f: round(tm[f], 3) for f in api.TIMING_FIELDS_F}, "counts": {f: tm[f] for f in api.TIMING_FIELDS_I},
                          "kernels": {nm: [v[0], round(v[1], 3)] for nm, v in sorted(rep.items(), key=lambda kv: -kv[1][1]) if v[1] > 0.02 * tot}}), flush=True)
        ctx.close()
        del keep, b_dev
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
