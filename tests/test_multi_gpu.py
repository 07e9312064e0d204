"""Multi-GPU parity on real devices: runs tests/dist_check.py under torchrun with every visible GPU (>= 2), NCCL
backend.  Skipped on single-GPU boxes; the host-side logic of the same path is covered on CPU by test_multi_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_path_matches_oracle_on_all_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(here, "dist_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    assert r.stdout.count("identical to oracle: True") == 3 and r.stdout.count("pre-filtered input: True") == 2, r.stdout[-2000:]
    assert r.stdout.count("identical on every rank: True") == 2, r.stdout[-2000:]
    assert "ranges stitch: True" in r.stdout
