import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_data():
    """config-1-shaped workload, reduced: 2 chromosomes, 30x, 7 planted SVs, noise."""
    from breakid_b200 import api, synth
    cfg = synth.SynthConfig(chrom_lens=[260000, 180000], n_tra=2, n_inv=2, n_dup=2, n_del=1, seed=5, sv_jitter=1)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(cfg.chrom_lens)]
    return d, hb, nibs


@pytest.fixture(scope="session")
def config1_data():
    """BASELINE.json configs[0]: 1 Mb two-chromosome genome, 30x, 10 planted SVs."""
    from breakid_b200 import api, synth
    d = synth.generate(synth.config1())
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, d.cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(d.cfg.chrom_lens)]
    return d, hb, nibs


def lattice_points(rng, n, kind):
    """adversarial inputs for sort / mask / clustering: heavy ties, several far-apart groups"""
    if kind == 0:
        x = rng.randint(0, 6, n) * 50 + 1000; y = rng.randint(0, 6, n) * 50 + 5000
    elif kind == 1:
        k = max(1, n // 12); cx = rng.randint(0, 200000, k); cy = rng.randint(0, 200000, k)
        a = rng.randint(0, k, n); x = cx[a] + rng.randint(0, 300, n); y = cy[a] + rng.randint(0, 300, n)
    elif kind == 2:
        x = rng.randint(0, 5, n) * 40 + 3000000000; y = rng.randint(0, 5, n) * 40 + 4000000000
    else:
        k = max(1, n // 8); a = rng.randint(0, k, n)
        x = a * 100000 + rng.randint(0, 4, n) * 50; y = (a % 3) * 70000 + rng.randint(0, 4, n) * 50 + 100
    return x.astype(np.uint32), y.astype(np.uint32)
