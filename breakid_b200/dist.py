"""Multi-GPU orchestration of the hot path: one process per GPU, torch.distributed for
the plumbing (NCCL over NVLink on GPUs; gloo on CPU for the host-logic tests).

Sharding: every rank holds a contiguous slice of the coordinate-sorted record stream (genomic bins).

  stage                         exchange
  ---------------------------   ----------------------------------------------------------------
  classify + insert statistics  all-reduce (sum |isize|, count, sum isize^2; max |isize|)   [C2]
  truncating sd accumulator     one pass on all ranks at once + all-reduce (sum floor, #correctable): exact and order
                                independent whenever no record can need a rounding correction below the accumulator's
                                final binade (the normal case); otherwise the exact replay chained rank to rank
  candidate records             all-to-all by name-hash owner                      [C1]  (mates meet)
  discordant pairs              all-to-all by bucket owner (chr-pair bucket rank)  [C1']
  mask + clustering + summary   none (buckets are independent, src/BreakID.cc:119-167)
  cluster summaries             all-gather (small)
  split-read evidence rows      all-gather (small, stays in coordinate order)      [C3]
  region coverage / bp depth    partial counts per shard, all-reduce (sum)
  vote / AF / 41-mers           replicated (tiny)

The per-rank compute is an *engine*: ``GpuEngine`` (the CUDA library through the bkid_shard_* C ABI) or,
in tests only, an engine built on the CPU oracle.  Everything here is backend agnostic.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import api

CAND_B, PAIR_B, CLUSTER_B, SAROW_B = 48, 64, 192, 88


class GpuEngine:
    """per-rank compute on one B200 through the shard entry points of the C ABI"""

    def __init__(self, ctx: api.Context, device: torch.device):
        self.ctx, self.lib, self.device = ctx, ctx.lib, device
        self._keep = {}

    def _chk(self, rc):
        self.ctx._chk(rc)

    def _ready(self, t):
        """tensors produced by torch / NCCL live on torch's streams; the library works on its own non-blocking
        stream, so make them visible before handing their pointers over"""
        t = t.contiguous()
        torch.cuda.synchronize(self.device)
        return t

    def _take(self, ptr, nbytes, row):
        t = torch.empty((nbytes // row, row) if row else (nbytes,), dtype=torch.uint8, device=self.device)
        if nbytes:
            self._chk(self.lib.bkid_device_copy(self.ctx.ctx, t.data_ptr(), ptr, nbytes))
        return t

    def insert_partial(self):
        s, n, q, m = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_uint64()
        self._chk(self.lib.bkid_shard_insert_partial(self.ctx.ctx, C.byref(s), C.byref(n), C.byref(q), C.byref(m)))
        return s.value, n.value, q.value, m.value

    def sd_fast(self, mean, kub):
        self._chk(self.lib.bkid_shard_sd_fast(self.ctx.ctx, mean, kub))

    def sd_fast_collect(self):
        f, e = C.c_uint64(), C.c_uint64()
        self._chk(self.lib.bkid_shard_sd_fast_collect(self.ctx.ctx, C.byref(f), C.byref(e)))
        return f.value, e.value

    def sd_partial(self, mean, t_in):
        t = C.c_int64()
        self._chk(self.lib.bkid_shard_sd_partial(self.ctx.ctx, mean, t_in, C.byref(t)))
        return t.value

    def sd_prepare(self, mean):
        self._chk(self.lib.bkid_shard_sd_prepare(self.ctx.ctx, mean))

    def set_stats(self, mean, sd):
        self._chk(self.lib.bkid_shard_set_stats(self.ctx.ctx, mean, sd))

    def candidates(self, index_offset):
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self.lib.bkid_shard_candidates(self.ctx.ctx, index_offset, C.byref(p), C.byref(n)))
        return self._take(p, n.value * CAND_B, CAND_B)

    def join(self, cands, w):
        cands = self._ready(cands)
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self.lib.bkid_shard_join(self.ctx.ctx, cands.data_ptr(), cands.shape[0], w, C.byref(p), C.byref(n)))
        return self._take(p, n.value * PAIR_B, PAIR_B)

    def set_pairs(self, pairs):
        pairs = self._ready(pairs)
        self._chk(self.lib.bkid_shard_set_pairs(self.ctx.ctx, pairs.data_ptr(), pairs.shape[0]))

    def cluster(self, d, mode):
        return self.ctx.cluster(d, mode)

    def clusters(self):
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self.lib.bkid_shard_clusters(self.ctx.ctx, C.byref(p), C.byref(n)))
        return self._take(p, n.value * CLUSTER_B, CLUSTER_B)

    def set_clusters(self, t):
        t = self._ready(t)
        self._chk(self.lib.bkid_shard_set_clusters(self.ctx.ctx, t.data_ptr(), t.shape[0]))

    def sa_rows(self):
        p, n = C.c_void_p(), C.c_int64()
        self._chk(self.lib.bkid_shard_sa_rows(self.ctx.ctx, C.byref(p), C.byref(n)))
        return self._take(p, n.value * SAROW_B, SAROW_B)

    def set_sa_rows(self, t):
        self._keep["rows"] = self._ready(t)
        self._chk(self.lib.bkid_shard_set_sa_rows(self.ctx.ctx, self._keep["rows"].data_ptr(), t.shape[0]))

    def maxspan(self):
        m = C.c_int32()
        self._chk(self.lib.bkid_shard_maxspan(self.ctx.ctx, C.byref(m)))
        return m.value

    def set_maxspan(self, m):
        self._chk(self.lib.bkid_shard_set_maxspan(self.ctx.ctx, m))

    def _counts(self, fn, *a):
        p, n = C.c_void_p(), C.c_int64()
        self._chk(fn(self.ctx.ctx, *a, C.byref(p), C.byref(n)))
        t = torch.empty(n.value, dtype=torch.int32, device=self.device)
        if n.value:
            self._chk(self.lib.bkid_device_copy(self.ctx.ctx, t.data_ptr(), p, n.value * 4))
        return t, p

    def coverage(self, d):
        t, self._cov_ptr = self._counts(self.lib.bkid_shard_coverage, d)
        return t

    def commit_coverage(self, t):
        t = self._ready(t)
        if t.numel():
            self._chk(self.lib.bkid_device_copy(self.ctx.ctx, self._cov_ptr, t.data_ptr(), t.numel() * 4))

    def vote(self):
        self._chk(self.lib.bkid_shard_vote(self.ctx.ctx))

    def depth(self):
        t, self._dep_ptr = self._counts(self.lib.bkid_shard_depth)
        return t

    def commit_depth(self, t):
        t = self._ready(t)
        if t.numel():
            self._chk(self.lib.bkid_device_copy(self.ctx.ctx, self._dep_ptr, t.data_ptr(), t.numel() * 4))

    def finish(self):
        n = C.c_int64()
        self._chk(self.lib.bkid_shard_finish(self.ctx.ctx, C.byref(n)))
        return self.ctx.fetch_clusters()

    def gather_rows(self, t, order):
        """t[order] for [n, row] byte rows (row % 16 == 0) with the library's 16-bytes-per-thread gather"""
        t = self._ready(t)
        out = torch.empty((order.shape[0], t.shape[1]), dtype=torch.uint8, device=self.device)
        order = self._ready(order.to(torch.int64))
        self._chk(self.lib.bkid_device_gather_rows(self.ctx.ctx, out.data_ptr(), t.data_ptr(), order.data_ptr(), order.shape[0], t.shape[1]))
        return out

    def bucket_ranks(self):
        nb = C.c_int64()
        self._chk(self.lib.bkid_fetch_bucket_ranks(self.ctx.ctx, None, 0, C.byref(nb)))
        out = np.zeros(max(1, nb.value), np.int32)
        self._chk(self.lib.bkid_fetch_bucket_ranks(self.ctx.ctx, out.ctypes.data, out.shape[0], C.byref(nb)))
        return torch.from_numpy(out[:nb.value].copy()).to(self.device)


class LibraryDist:
    """The sharded path with the exchanges INSIDE the library (``bkid_dist_run``, breakid_b200/csrc/bkid_dist.cuh): NCCL calls
    issued by the C++ code on the context's stream.  torch.distributed is used exactly once, to hand rank 0's NCCL unique id
    to the other ranks."""

    def __init__(self, ctx: api.Context, device: torch.device):
        self.ctx, self.lib = ctx, ctx.lib
        W, r = _world(), _rank()
        ident = torch.zeros(128, dtype=torch.uint8)
        if r == 0:
            buf = (C.c_uint8 * 128)()
            if self.lib.bkid_comm_nccl_unique_id(buf):
                raise api.BkidError("bkid_comm_nccl_unique_id: " + (self.lib.bkid_last_error(None) or b"").decode())
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if W > 1:
            t = ident.to(device)
            dist.broadcast(t, src=0)
            ident = t.cpu()
        raw = (C.c_uint8 * 128)(*ident.tolist())
        self.comm = self.lib.bkid_comm_nccl_init(raw, r, W, device.index if device.index is not None else 0)
        if not self.comm:
            raise api.BkidError("bkid_comm_nccl_init: " + (self.lib.bkid_last_error(None) or b"").decode())

    def run(self, mode: int = 0):
        """collective; returns (mean, sd, dist, n_called, stage_ms[9])"""
        m, s, d, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        tm = (C.c_float * 9)()
        self.ctx._chk(self.lib.bkid_dist_run(self.ctx.ctx, self.comm, mode, C.byref(m), C.byref(s), C.byref(d), C.byref(n), tm))
        return m.value, s.value, d.value, n.value, list(tm)

    def close(self):
        if self.comm:
            self.lib.bkid_comm_destroy(self.comm)
            self.comm = None


DIST_STAGES = ("insert statistics", "candidates", "a2a candidates", "join", "a2a pairs", "mask + cluster", "gathers", "refine", "total")


# ---------------------------------------------------------------------------------------------------
def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_initialized() else 0


def _all_gather_rows(t: torch.Tensor) -> torch.Tensor:
    """concatenate variable-length [n_r, row] byte tensors of all ranks in rank order"""
    W = _world()
    if W == 1:
        return t
    cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    cnts = [torch.zeros_like(cnt) for _ in range(W)]
    dist.all_gather(cnts, cnt)
    cnts = [int(c) for c in cnts]
    mx = max(cnts + [1])
    pad = torch.zeros((mx, t.shape[1]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(W)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, cnts)], 0)


_A2A_T = {}


def _all_to_all_rows(t: torch.Tensor, owner: torch.Tensor, engine=None) -> torch.Tensor:
    """route row i of t to rank owner[i]; rows arrive grouped by source rank (rank order) and keep their
    source order inside a group -- i.e. a stream that was globally ordered stays globally ordered"""
    import os, time
    W = _world()
    if W == 1:
        return t
    dbg = os.environ.get("BKID_DEBUG_TIMING") and t.is_cuda
    t0 = [time.perf_counter()]

    def lap(k):
        if dbg:
            torch.cuda.synchronize()
            now = time.perf_counter()
            _A2A_T[k] = _A2A_T.get(k, 0.0) + (now - t0[0]) * 1e3
            t0[0] = now
    order = torch.sort(owner.to(torch.uint8), stable=True).indices          # W <= 255: one 8-bit radix pass
    lap("sort")
    if engine is not None and hasattr(engine, "gather_rows") and t.shape[1] % 16 == 0:
        send = engine.gather_rows(t, order)
    elif t.shape[1] % 8 == 0:          # move rows as 8-byte words, not bytes
        send = torch.index_select(t.view(torch.int64), 0, order).view(torch.uint8)          # row gather (advanced indexing was 10x slower)
    else:
        send = torch.index_select(t, 0, order)
    lap("gather")
    scount = torch.bincount(owner.to(torch.int32), minlength=W).to(torch.int64)
    rcount = torch.empty_like(scount)
    dist.all_to_all_single(rcount, scount) if dist.get_backend() != "gloo" else _gloo_a2a_counts(rcount, scount)
    s_list, r_list = [int(x) for x in scount], [int(x) for x in rcount]
    lap("counts")
    recv = torch.empty((sum(r_list), t.shape[1]), dtype=t.dtype, device=t.device)
    if dist.get_backend() == "gloo":
        _gloo_a2a_rows(recv, send, s_list, r_list)
    else:
        dist.all_to_all_single(recv, send, output_split_sizes=r_list, input_split_sizes=s_list)
    lap("a2a")
    return recv


def _gloo_a2a_counts(rcount, scount):
    W = _world()
    allc = [torch.zeros_like(scount) for _ in range(W)]
    dist.all_gather(allc, scount)
    r = _rank()
    for src in range(W):
        rcount[src] = allc[src][r]


def _gloo_a2a_rows(recv, send, s_list, r_list):
    """all-to-all emulated with all-gather (gloo CPU tests only)"""
    W, r = _world(), _rank()
    allsend = _all_gather_rows(send)
    # counts matrix
    sc = torch.tensor(s_list, dtype=torch.int64)
    allc = [torch.zeros_like(sc) for _ in range(W)]
    dist.all_gather(allc, sc)
    off = 0
    o = 0
    for src in range(W):
        row0 = off + int(allc[src][:r].sum())
        k = int(allc[src][r])
        recv[o:o + k] = allsend[row0:row0 + k]
        o += k
        off += int(allc[src].sum())


def _reduce(t: torch.Tensor, op):
    if _world() > 1 and t.numel():
        dist.all_reduce(t, op=op)
    return t


def sd_upper_binade(S: int, N: int, SQ: int, XM: int) -> int:
    """upper bound of the binade of the truncating sd accumulator from the global sums -- the same integer arithmetic as
    bkid_sd_upper_binade (breakid_b200/csrc/bkid_api.cuh: sd_upper_binade)"""
    if N == 0:
        return 0
    if XM > 65535:
        return 51
    T = (SQ * N - S * S) // N + 1
    T += (T >> 30) + 2 * N + 1024
    if T >> 51:
        return 51
    return T.bit_length() - 1


def run_sharded(engine, n_local: int, times: int = 2, sd_mult: int = 3, mode: int = 0, timing: Optional[dict] = None):
    """the whole hot path over all ranks; returns (mean, sd, dist, called cluster records [numpy, bucket =
    dense id]) -- identical on every rank and identical to the single-GPU / reference result"""
    import time as _time
    W, r = _world(), _rank()
    dev = engine.device
    _t = [_time.perf_counter()]

    def lap(name):
        if timing is not None:
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            now = _time.perf_counter()
            timing[name] = timing.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now
    # insert statistics: exact integer sum/count, then the order-dependent sd accumulator chained through the ranks
    s, n, sq, xm = engine.insert_partial()
    sn = _reduce(torch.tensor([s, n, sq & 0xffffffff, sq >> 32], dtype=torch.int64, device=dev), dist.ReduceOp.SUM)
    S, N, SQ = int(sn[0]), int(sn[1]), int(sn[2]) + (int(sn[3]) << 32)
    XM = int(_reduce(torch.tensor([xm], dtype=torch.int64, device=dev), dist.ReduceOp.MAX)[0])
    mean = (float(S) / float(N)) if N else float("nan")      # no proper pair: NaN like the single-GPU path and the reference (0/0 in double)
    kub = sd_upper_binade(S, N, SQ, XM)
    engine.sd_fast(mean, kub)            # one streaming pass on all ranks at once, asynchronously: it overlaps the candidate extraction below
    lap('insert sum/count')
    # global index of the first local record
    ns = torch.zeros(W, dtype=torch.int64, device=dev)
    ns[r] = n_local
    _reduce(ns, dist.ReduceOp.SUM)
    offset = int(ns[:r].sum())
    # candidates meet their mates on the owner of their name hash (independent of the distance: before the sd chain)
    cands = engine.candidates(offset)
    lo = cands.view(torch.int64)[:, 0] if cands.shape[0] else torch.zeros(0, dtype=torch.int64, device=dev)
    owner = ((lo >> 8) & 0x7fffffff) % W
    lap('candidates')
    cands = _all_to_all_rows(cands, owner, engine)
    lap('a2a candidates')
    # sd accumulator: sum floor(a) over all ranks is the exact total when no record can need a correction (E == 0);
    # only otherwise the order-dependent replay is chained through the ranks
    f, e = engine.sd_fast_collect()
    fe = _reduce(torch.tensor([f, min(e, 1 << 40)], dtype=torch.int64, device=dev), dist.ReduceOp.SUM)
    t = torch.zeros(1, dtype=torch.int64, device=dev)
    if int(fe[1]) == 0:
        t[0] = int(fe[0])
    elif N:
        if hasattr(engine, "sd_prepare"):
            engine.sd_prepare(mean)      # block tables on all ranks at once
        for src in range(W):
            if r == src:
                t[0] = engine.sd_partial(mean, int(t[0]))
            if W > 1:
                dist.broadcast(t, src=src)
    lap('sd')
    sd = math.sqrt(int(t[0]) / float(N)) if N else float("nan")
    d = times * math.sqrt(times) * (mean + sd_mult * sd)
    engine.set_stats(mean, sd)
    pairs = engine.join(cands, d)
    lap('join')
    # pairs go to the owner of their chr-pair bucket
    bucket = pairs.view(torch.int32)[:, 12].to(torch.int64) if pairs.shape[0] else torch.zeros(0, dtype=torch.int64, device=dev)
    pairs = _all_to_all_rows(pairs, bucket % W, engine)
    lap('a2a pairs')
    engine.set_pairs(pairs)
    lap('set_pairs')
    # dense bucket ids of the reference = rank among ALL buckets that hold at least one pair, over all ranks
    br = engine.bucket_ranks().to(torch.int32).reshape(-1, 1).contiguous().view(torch.uint8)
    all_ranks = torch.sort(_all_gather_rows(br).view(torch.int32).reshape(-1)).values
    engine.cluster(d, mode)
    lap('mask+cluster')
    # every rank gets every cluster summary, in the reference's order (bucket name rank, cluster id)
    cl = _all_gather_rows(engine.clusters())
    if cl.shape[0]:
        v = cl.view(torch.int32)
        key = (v[:, 0].to(torch.int64) << 32) | v[:, 1].to(torch.int64)
        cl = torch.index_select(cl.view(torch.int64), 0, torch.sort(key, stable=True).indices).view(torch.uint8)
    engine.set_clusters(cl)
    lap('gather clusters')
    rows = _all_gather_rows(engine.sa_rows())
    if rows.shape[0] and W > 1:
        # the table is searched by (tid, pos): restore coordinate order when the slices are not genomic bins
        v = rows.view(torch.int32)
        key = ((v[:, 18].to(torch.int64) & 0xffffffff) << 32) | (v[:, 19].to(torch.int64) & 0xffffffff)
        rows = torch.index_select(rows.view(torch.int64), 0, torch.sort(key, stable=True).indices).view(torch.uint8)
    engine.set_sa_rows(rows)
    lap('gather sa rows')
    ms = _reduce(torch.tensor([engine.maxspan()], dtype=torch.int32, device=dev), dist.ReduceOp.MAX)
    engine.set_maxspan(int(ms[0]))
    engine.commit_coverage(_reduce(engine.coverage(d), dist.ReduceOp.SUM))
    lap('coverage + allreduce')
    engine.vote()
    lap('vote')
    engine.commit_depth(_reduce(engine.depth(), dist.ReduceOp.SUM))
    out = engine.finish()
    lap('depth + finish')
    if len(out):
        out = out.copy()
        out["bucket"] = np.searchsorted(all_ranks.cpu().numpy(), out["bucket"]).astype(np.int32)   # name rank -> dense id
    return mean, sd, d, out
