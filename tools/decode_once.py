#!/usr/bin/env python
"""One device decode of a small synthetic BAM (for ncu captures of the decode kernels).  usage: decode_once.py [scale] [reps]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from breakid_b200 import api, bamio, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0 / 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = synth.generate(synth.config2(scale=scale))
tmp = tempfile.mkdtemp(prefix="bkid_once_")
bam = os.path.join(tmp, "r.bam")
bamio.write_bam(bam, d, random_qual=True)
f = api.BgzfFile(bam)
raw = torch.from_numpy(np.fromfile(bam, dtype=np.uint8)).pin_memory()
ctx = api.Context(f.target_len, f.target_names, device=0)
for i in range(reps):
    ctx.reset()
    n = ctx.push_bgzf(f, data_ptr=raw.data_ptr())
st = ctx.decode_stats()
print(n, st)
