"""CPU test of the N>1 path: breakid_b200.dist.run_sharded in two processes over gloo, per-rank compute on
the CPU oracle's sharded decomposition (oracle/oracle_engine.py).  The result has to be byte-identical to
orc_run on the unsplit input -- i.e. routing by name hash / bucket owner, the chained sd accumulator,
partial coverage / depth sums and the global evidence table reproduce the reference exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slice(hb, a, b):
    from breakid_b200 import api
    sa = hb.side["sa_rec"]; lo = np.searchsorted(sa, a); hi = np.searchsorted(sa, b)
    co, so, oo = hb.side["cig_off"], hb.side["sa_off"], hb.side["oc_off"]
    return api.HostBatch({k: v[a:b] for k, v in hb.cols.items()}, hb.name_hash[2 * a:2 * b],
                         {"sa_rec": sa[lo:hi] - a, "cig_off": co[lo:hi + 1] - co[lo], "cig_ops": hb.side["cig_ops"][co[lo]:co[hi]],
                          "sa_off": so[lo:hi + 1] - so[lo], "sa_txt": hb.side["sa_txt"][so[lo]:so[hi]],
                          "oc_off": oo[lo:hi + 1] - oo[lo], "oc_txt": hb.side["oc_txt"][oo[lo]:oo[hi]]}, hb.target_len, hb.target_names)


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from breakid_b200 import api, synth
    from breakid_b200.dist import run_sharded
    from oracle_engine import OracleEngine
    cfg = synth.SynthConfig(chrom_lens=[160000, 110000, 90000], n_tra=3, n_inv=1, n_dup=1, n_del=2, seed=17, sv_jitter=1, min_sv_sep=5000)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(cfg.chrom_lens)]
    cuts = [0] + [hb.n * (i + 1) // world + (7 if i + 1 < world else 0) for i in range(world)]      # uneven on purpose
    part = _slice(hb, cuts[rank], cuts[rank + 1])
    mean, sd, dd, out = run_sharded(OracleEngine(part, nibs), part.n, mode=mode)
    q.put((rank, mean, sd, dd, out.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_two_rank_sharded_path_equals_single(mode):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as O
    from breakid_b200 import api, synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + mode
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = synth.SynthConfig(chrom_lens=[160000, 110000, 90000], n_tra=3, n_inv=1, n_dup=1, n_del=2, seed=17, sv_jitter=1, min_sv_sep=5000)
    d = synth.generate(cfg)
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(cfg.chrom_lens)]
    m, s, dd, exp = O.run(hb, nibs, mode=mode)
    assert len(exp) >= 5
    for (rank, mean, sd, d2, blob) in res:
        assert (mean, sd, d2) == (m, s, dd), rank
        assert blob == exp.tobytes(), rank
