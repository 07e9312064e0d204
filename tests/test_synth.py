"""The synthetic generator (breakid_b200/synth.py).  The host and the device run of `generate` draw from
different torch generators, so their outputs are not equal element by element; both must satisfy the same invariants of a
coordinate-sorted, mate-consistent BAM with the planted SVs, and both must lead the oracle to the planted calls."""
import numpy as np
import pytest

from breakid_b200 import api, synth

CFG = dict(chrom_lens=[260000, 180000, 90000], n_tra=3, n_inv=2, n_dup=2, n_del=2, seed=17, sv_jitter=1)


def check_invariants(d):
    c = {k: v.cpu().numpy() for k, v in d.cols.items()}
    n = d.n
    tid, pos = c["tid"].astype(np.int64), c["pos"].astype(np.int64)
    key = np.where(tid >= 0, tid * (1 << 32) + pos, np.iinfo(np.int64).max)
    assert np.all(np.diff(key) >= 0), "not coordinate sorted"
    lens = np.asarray(d.cfg.chrom_lens)
    placed = tid >= 0
    assert np.all(pos[placed] >= 0) and np.all(c["endpos"][placed] <= lens[tid[placed]])
    flag = c["flag"].astype(np.int64) & 0xffff
    prim = (flag & 0x900) == 0                                   # neither secondary nor supplementary
    # every read name has exactly one first and one second primary record, and they point at each other
    order = np.lexsort((flag & 0xc0, c["name_id"]))
    nid = c["name_id"][order][prim[order]]
    assert len(nid) % 2 == 0 and np.all(nid[0::2] == nid[1::2]) and (len(nid) < 4 or np.all(nid[2::2] != nid[1:-1:2]))
    a, b = order[prim[order]][0::2], order[prim[order]][1::2]
    assert np.all(flag[a] & 0x40) and np.all(flag[b] & 0x80)
    assert np.array_equal(c["mtid"][a], c["tid"][b]) and np.array_equal(c["mpos"][a], c["pos"][b])
    assert np.array_equal(c["mtid"][b], c["tid"][a]) and np.array_equal(c["mpos"][b], c["pos"][a])
    assert np.array_equal(c["isize"][a], -c["isize"][b])
    # the SA side table is ascending in record index and points at SA-bearing records
    sa = d.sa_rec.cpu().numpy()
    assert np.all(np.diff(sa) > 0) and (len(sa) == 0 or sa[-1] < n)
    assert len(d.sa_off) == len(sa) + 1 and len(d.cig_off) == len(sa) + 1
    return n


def planted_calls(d):
    import oracle_py as O
    hb = api.HostBatch.from_synth(d)
    nibs = [(synth.random_nib_bytes(l, d.cfg.seed * 1000 + t).numpy(), l) for t, l in enumerate(d.cfg.chrom_lens)]
    _, _, _, calls = O.run(hb, nibs, mode=0)
    return calls


def test_host_generator_invariants_and_planted_calls():
    d = synth.generate(synth.SynthConfig(**CFG))
    check_invariants(d)
    calls = planted_calls(d)
    assert len(calls) >= 7 and int((calls["n_split_read"] > 0).sum()) >= 7        # 9 planted SVs


@pytest.mark.gpu
def test_device_generator_satisfies_the_same_invariants():
    cfg = synth.SynthConfig(**CFG)
    dh, dd = synth.generate(cfg), synth.generate(cfg, device="cuda")
    nh, nd = check_invariants(dh), check_invariants(dd)
    assert abs(nh - nd) <= 0.02 * nh                               # same coverage model, different random stream
    ch, cd = planted_calls(dh), planted_calls(dd)
    assert len(cd) >= 7 and abs(len(ch) - len(cd)) <= 2
