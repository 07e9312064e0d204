import sys, torch
sys.path.insert(0,'/root/repo')
import bench
from breakid_b200 import api, synth
cfg = synth.config2(1.0)
d = synth.generate(cfg, device='cuda:0')
b, keep = bench.device_batch(d)
names=[synth.chrom_name(t) for t in range(24)]
ctx = api.Context(cfg.chrom_lens, names)
for i in range(3):
    ctx.reset(); ctx.push_device(b); print(ctx.insert_stats())
