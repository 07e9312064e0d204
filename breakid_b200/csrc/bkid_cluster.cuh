// bkid_cluster.cuh -- K5a (AHC), K5b (-fast sweep), K6 (cluster summary).  Included by bkid_core.cu.
//
// AHC (reference src/util_cluster.cc:7-396; default mode, src/BreakID.cc:135): the reference builds
// an N x N double matrix and per-node sorted neighbour lists, O(N^3).  Device formulation (proved
// equal to the reference on the CPU, oracle/oracle.cc orc_model_ahc_tree):
//   * points split into COMPONENTS (x-gap > thr, then y-gap > thr): no linkage across components
//     can be <= thr, so each component keeps only in-component neighbour ROWS;
//   * the sorted list is replaced by a closed-form winner rule with the tail exception of
//     insert_sorted (src/util_cluster.cc:266-273) expressed through creation ranks;
//   * FP64 arithmetic is the reference's, op for op: sqrt(dx*dx+dy*dy) without FMA, average
//     linkage as the sequential sum over a.points x b.points divided by (double)(int)(m*n).
// Kernel A runs one warp per component (speculating that an out-of-component sentinel exists where
// that is not statically known, and flagging the component if a decision depended on it); kernel B
// replays the per-component merge events per bucket in the reference's global order to get the
// global creation ranks (cluster numbering); buckets with a flagged component are re-run by kernel
// C, the exact online form (one warp per bucket, true ranks).
#pragma once

struct AhcView {
  // bucket-level
  const uint32_t *seg_off;      // [nb+1] point ranges per bucket
  const uint32_t *X, *Y;        // [np] coordinates in leaf order
  // component-level
  const uint32_t *comp_off;     // [ncomp+1] into comp_leaf
  const uint32_t *comp_leaf;    // [np] global point index p (ascending inside a component)
  const uint32_t *comp_bucket;  // [ncomp]
  const int32_t *lo_oc;         // [np] per leaf slot (aligned with comp_leaf): highest bucket-local leaf index < leaf not in its component
  const int32_t *hi_oc;         // [ncomp] highest bucket-local leaf index not in the component
  const unsigned long long *pts_off, *row_off;   // [ncomp] pool offsets
  // node-level (index = 2*comp_off[c] + local id)
  uint8_t *node_root;
  uint32_t *node_npts;
  unsigned long long *node_pts;   // absolute offset into pts_pool (merged nodes)
  unsigned long long *node_row;   // absolute offset into row pools (merged nodes)
  uint32_t *node_rowlen;
  int32_t *node_best_t;
  double *node_best_d;
  int32_t *node_grank;
  int32_t *node_ma, *node_mb;     // merged children (local ids)
  // pools
  uint32_t *pts_pool;             // bucket-local leaf indices
  int32_t *row_t;
  double *row_d;
  // per component state
  uint32_t *comp_nnodes;
  unsigned long long *comp_pts_used, *comp_row_used;
  int32_t *comp_head_j;
  double *comp_head_d;
  int32_t *comp_flag;
  uint32_t *comp_cursor;          // replay cursor
  // events (index = comp_off[c] + lrank)
  double *ev_d;
  int32_t *ev_first;              // local id of `first`
  double thr;
};

__device__ __forceinline__ double ahc_euclid(const AhcView &v, uint32_t base, uint32_t a, uint32_t b)
{
  // src/util_cluster.cc:79-84 compiled without FMA: mulsd, mulsd, addsd, sqrtsd
  double dx = __dsub_rn((double)v.X[base + a], (double)v.X[base + b]);
  double dy = __dsub_rn((double)v.Y[base + a], (double)v.Y[base + b]);
  return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

struct CompCtx {
  uint32_t comp, c, lbase /* comp_off[comp] */, pbase /* seg_off[bucket] */, nbase /* node base */;
};

__device__ __forceinline__ uint32_t leaf_of(const AhcView &v, const CompCtx &cc, uint32_t local) { return v.comp_leaf[cc.lbase + local] - cc.pbase; }

// distance from node j to target t (both local ids), t < j, for entry e of j's row / leaf row
__device__ __forceinline__ void row_entry(const AhcView &v, const CompCtx &cc, uint32_t j, uint32_t e, int32_t &t, double &d)
{
  if (j < cc.c) { t = (int32_t)e; d = ahc_euclid(v, cc.pbase, leaf_of(v, cc, j), leaf_of(v, cc, e)); }
  else { unsigned long long o = v.node_row[cc.nbase + j] + e; t = v.row_t[o]; d = v.row_d[o]; }
}

// winner among the still-root targets of node j (warp-cooperative).  exact: use true creation ranks.
// g_j / l_j: global and in-component creation rank of merged node j (merged nodes only).
template <bool EXACT>
__device__ void ahc_find_best(const AhcView &v, const CompCtx &cc, uint32_t j)
{
  const unsigned lane = threadIdx.x & 31;
  uint32_t len = (j < cc.c) ? j : v.node_rowlen[cc.nbase + j];
  // pass 1: minimal distance among still-root targets, ties -> lowest index
  double dmin = 1.7976931348623157e308; int32_t tmin = -1;
  for (uint32_t e = lane; e < len; e += 32) {
    int32_t t; double d;
    row_entry(v, cc, j, e, t, d);
    if (v.node_root[cc.nbase + t] && (tmin < 0 || d < dmin || (d == dmin && t < tmin))) { dmin = d; tmin = t; }
  }
  for (int o = 16; o; o >>= 1) {
    double od = __shfl_xor_sync(0xffffffffu, dmin, o);
    int32_t ot = __shfl_xor_sync(0xffffffffu, tmin, o);
    if (ot >= 0 && (tmin < 0 || od < dmin || (od == dmin && ot < tmin))) { dmin = od; tmin = ot; }
  }
  int32_t best = tmin;
  if (tmin >= 0) {
    // pass 2: two highest indices t1 > t2 over the WHOLE row with D == dmin
    int32_t t1 = -1, t2 = -1;
    for (uint32_t e = lane; e < len; e += 32) {
      int32_t t; double d;
      row_entry(v, cc, j, e, t, d);
      if (d == dmin) { if (t > t1) { t2 = t1; t1 = t; } else if (t > t2) t2 = t; }
    }
    for (int o = 16; o; o >>= 1) {
      int32_t a1 = __shfl_xor_sync(0xffffffffu, t1, o), a2 = __shfl_xor_sync(0xffffffffu, t2, o);
      // merge (t1,t2) with (a1,a2)
      int32_t n1 = max(t1, a1);
      int32_t n2 = max(min(t1, a1), max(t2, a2));
      t1 = n1; t2 = n2;
    }
    if (t2 >= 0 && tmin == t2 && v.node_root[cc.nbase + t1]) {
      // pass 3: in-component part of the tail test
      bool blocked = false;
      for (uint32_t e = lane; e < len; e += 32) {
        int32_t t; double d;
        row_entry(v, cc, j, e, t, d);
        if (t > t2 && t != t1 && !(d < dmin)) blocked = true;
      }
      blocked = __any_sync(0xffffffffu, blocked);
      if (!blocked) {
        bool sentinel;
        if (j < cc.c) sentinel = v.lo_oc[cc.lbase + j] > (int32_t)leaf_of(v, cc, (uint32_t)t2);
        else if ((uint32_t)t2 < cc.c && v.hi_oc[cc.comp] > (int32_t)leaf_of(v, cc, (uint32_t)t2)) sentinel = true;   // static
        else if (EXACT) {
          int32_t gj = v.node_grank[cc.nbase + j], lj = (int32_t)(j - cc.c);
          if ((uint32_t)t2 >= cc.c) sentinel = (gj - v.node_grank[cc.nbase + t2]) > (lj - (int32_t)((uint32_t)t2 - cc.c));
          else sentinel = gj > lj;
        } else {
          sentinel = true;                         // speculation; the component is flagged for the exact pass
          if (lane == 0) v.comp_flag[cc.comp] = 1;
        }
        if (!sentinel) best = t1;
      }
    }
  }
  if (lane == 0) { v.node_best_t[cc.nbase + j] = best; v.node_best_d[cc.nbase + j] = dmin; }
  __syncwarp();
}

// head event of a component: min best_d over roots, ties -> highest local id (= highest global index inside a component)
__device__ void ahc_comp_head(const AhcView &v, const CompCtx &cc)
{
  const unsigned lane = threadIdx.x & 31;
  uint32_t nn = v.comp_nnodes[cc.comp];
  double hd = 1.7976931348623157e308; int32_t hj = -1;
  for (uint32_t j = lane; j < nn; j += 32)
    if (v.node_root[cc.nbase + j] && v.node_best_t[cc.nbase + j] >= 0) {
      double d = v.node_best_d[cc.nbase + j];
      if (hj < 0 || d < hd || (d == hd && (int32_t)j > hj)) { hd = d; hj = (int32_t)j; }
    }
  for (int o = 16; o; o >>= 1) {
    double od = __shfl_xor_sync(0xffffffffu, hd, o);
    int32_t oj = __shfl_xor_sync(0xffffffffu, hj, o);
    if (oj >= 0 && (hj < 0 || od < hd || (od == hd && oj > hj))) { hd = od; hj = oj; }
  }
  if (lane == 0) { v.comp_head_j[cc.comp] = hj; v.comp_head_d[cc.comp] = hd; }
  __syncwarp();
}

__device__ __forceinline__ uint32_t node_pt(const AhcView &v, const CompCtx &cc, uint32_t node, uint32_t k)
{
  return node < cc.c ? leaf_of(v, cc, node) : v.pts_pool[v.node_pts[cc.nbase + node] + k];
}

// merge head event of the component (warp-cooperative); grank = global creation rank or -1
template <bool EXACT>
__device__ void ahc_merge(const AhcView &v, const CompCtx &cc, int32_t grank)
{
  const unsigned lane = threadIdx.x & 31;
  uint32_t first = (uint32_t)v.comp_head_j[cc.comp];
  uint32_t second = (uint32_t)v.node_best_t[cc.nbase + first];
  uint32_t jn = v.comp_nnodes[cc.comp];
  uint32_t n1 = v.node_npts[cc.nbase + first], n2 = v.node_npts[cc.nbase + second];
  unsigned long long po = v.pts_off[cc.comp] + v.comp_pts_used[cc.comp];
  unsigned long long ro = v.row_off[cc.comp] + v.comp_row_used[cc.comp];
  __syncwarp();
  // points = first.points ++ second.points (src/util_cluster.cc:371-382)
  for (uint32_t k = lane; k < n1 + n2; k += 32)
    v.pts_pool[po + k] = k < n1 ? node_pt(v, cc, first, k) : node_pt(v, cc, second, k - n1);
  if (lane == 0) {
    v.node_root[cc.nbase + first] = 0; v.node_root[cc.nbase + second] = 0;
    v.node_root[cc.nbase + jn] = 1; v.node_npts[cc.nbase + jn] = n1 + n2;
    v.node_pts[cc.nbase + jn] = po; v.node_row[cc.nbase + jn] = ro;
    v.node_grank[cc.nbase + jn] = grank;
    v.node_ma[cc.nbase + jn] = (int32_t)first; v.node_mb[cc.nbase + jn] = (int32_t)second;
    v.comp_nnodes[cc.comp] = jn + 1;
    v.comp_pts_used[cc.comp] += n1 + n2;
  }
  __syncwarp();
  // row of the new node: average linkage to every in-component root (src/util_cluster.cc:201-215)
  uint32_t rl = 0;
  const uint32_t m = n1 + n2;
  for (uint32_t hi = jn; hi > 0; hi = hi > 32 ? hi - 32 : 0) {
    int32_t t = (int32_t)hi - 1 - (int32_t)lane;
    bool ok = t >= 0 && v.node_root[cc.nbase + t];
    double D = 0.0;
    if (ok) {
      uint32_t nt = v.node_npts[cc.nbase + t];
      double total = 0.0;
      for (uint32_t i = 0; i < m; ++i) {
        uint32_t a = v.pts_pool[po + i];
        for (uint32_t k = 0; k < nt; ++k) total = __dadd_rn(total, ahc_euclid(v, cc.pbase, a, node_pt(v, cc, (uint32_t)t, k)));
      }
      D = __ddiv_rn(total, (double)(int)(m * nt));
    }
    unsigned mk = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      unsigned long long o = ro + rl + __popc(mk & ((1u << lane) - 1u));
      v.row_t[o] = t; v.row_d[o] = D;
    }
    rl += __popc(mk);
  }
  if (lane == 0) { v.node_rowlen[cc.nbase + jn] = rl; v.comp_row_used[cc.comp] += rl; }
  __syncwarp();
  ahc_find_best<EXACT>(v, cc, jn);
  // roots whose winner was one of the merged nodes look again
  for (uint32_t b = 0; b < jn; b += 32) {
    uint32_t j = b + lane;
    bool need = j < jn && v.node_root[cc.nbase + j] &&
                (v.node_best_t[cc.nbase + j] == (int32_t)first || v.node_best_t[cc.nbase + j] == (int32_t)second);
    unsigned mk = __ballot_sync(0xffffffffu, need);
    while (mk) {
      int src = __ffs(mk) - 1;
      mk &= mk - 1;
      ahc_find_best<EXACT>(v, cc, b + (uint32_t)src);
    }
  }
  ahc_comp_head(v, cc);
}

__device__ void ahc_comp_init(const AhcView &v, const CompCtx &cc)
{
  const unsigned lane = threadIdx.x & 31;
  for (uint32_t j = lane; j < 2 * cc.c; j += 32) {
    v.node_root[cc.nbase + j] = j < cc.c ? 1 : 0;
    v.node_npts[cc.nbase + j] = j < cc.c ? 1u : 0u;
    v.node_rowlen[cc.nbase + j] = 0;
    v.node_best_t[cc.nbase + j] = -1;
    v.node_grank[cc.nbase + j] = -1;
    v.node_ma[cc.nbase + j] = -1; v.node_mb[cc.nbase + j] = -1;
  }
  if (lane == 0) { v.comp_nnodes[cc.comp] = cc.c; v.comp_pts_used[cc.comp] = 0; v.comp_row_used[cc.comp] = 0; v.comp_cursor[cc.comp] = 0; }
  __syncwarp();
}

__device__ __forceinline__ CompCtx make_cc(const AhcView &v, uint32_t comp)
{
  CompCtx cc;
  cc.comp = comp; cc.lbase = v.comp_off[comp]; cc.c = v.comp_off[comp + 1] - cc.lbase;
  cc.pbase = v.seg_off[v.comp_bucket[comp]]; cc.nbase = 2 * cc.lbase;
  return cc;
}

// Kernel A: one warp per component, speculative
__global__ void __launch_bounds__(128) ahc_components(AhcView v, uint32_t ncomp)
{
  uint32_t comp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (comp >= ncomp) return;
  const unsigned lane = threadIdx.x & 31;
  CompCtx cc = make_cc(v, comp);
  if (lane == 0) v.comp_flag[comp] = 0;
  ahc_comp_init(v, cc);
  for (uint32_t j = 1; j < cc.c; ++j) ahc_find_best<false>(v, cc, j);
  ahc_comp_head(v, cc);
  uint32_t lr = 0;
  while (true) {
    int32_t hj = v.comp_head_j[comp];
    double hd = v.comp_head_d[comp];
    if (hj < 0 || !(hd <= v.thr)) break;
    if (lane == 0) { v.ev_d[cc.lbase + lr] = hd; v.ev_first[cc.lbase + lr] = hj; }
    ahc_merge<false>(v, cc, -1);
    ++lr;
  }
}

// Kernel B (shared-memory form): one warp per bucket with every event, cursor and component head of the
// bucket staged in shared memory, so a replay step is a handful of LDS + shuffles instead of a chain
// of dependent global loads.  Components that kernel A flagged (a tie decision depended on global
// creation ranks -- common, because the isolated-pair mask duplicates one pair per bucket) are not
// replayed from their speculative events: they are re-run ONLINE inside this loop in exact mode, with
// the true ranks, while every other component of the bucket keeps its precomputed events.
// Buckets with n_lo <= points < n_hi are handled; smem = 49 B per point.
__global__ void __launch_bounds__(32) ahc_replay_smem(AhcView v, const uint32_t *__restrict__ bucket_comp_off, uint32_t nb, const int32_t *__restrict__ bucket_flag,
                                                      uint32_t n_lo, uint32_t n_hi)
{
  uint32_t b = blockIdx.x;
  if (b >= nb || !bucket_flag[b]) return;       // buckets without flagged components take the rank form
  uint32_t nleaf = v.seg_off[b + 1] - v.seg_off[b];
  if (nleaf < n_lo || nleaf >= n_hi) return;
  extern __shared__ unsigned char dynsm[];
  const unsigned lane = threadIdx.x;
  uint32_t c0 = bucket_comp_off[b], c1 = bucket_comp_off[b + 1], K = c1 - c0;
  double *ev_d = reinterpret_cast<double *>(dynsm);                    // [cap = nleaf]
  double *hd = ev_d + nleaf;
  int32_t *ev_first = reinterpret_cast<int32_t *>(hd + nleaf);
  int32_t *ev_g = ev_first + nleaf;
  int32_t *hg = ev_g + nleaf;
  uint32_t *cb = reinterpret_cast<uint32_t *>(hg + nleaf);              // event base of comp
  uint32_t *cn = cb + nleaf;                                             // number of precomputed events (0 for flagged comps)
  uint32_t *cc = cn + nleaf;                                             // leaf count
  uint32_t *cl = cc + nleaf;                                             // lbase
  uint32_t *cur = cl + nleaf;                                            // cursor
  uint8_t *cf = reinterpret_cast<uint8_t *>(cur + nleaf);               // flagged -> online exact mode
  const uint32_t pbase = v.seg_off[b];
  uint32_t run = 0;
  for (uint32_t q0 = 0; q0 < K; q0 += 32) {
    uint32_t q = q0 + lane, nev = 0;
    if (q < K) {
      uint32_t comp = c0 + q, lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase;
      uint8_t fl = v.comp_flag[comp] ? 1 : 0;
      nev = fl ? 0u : v.comp_nnodes[comp] - c;
      cc[q] = c; cl[q] = lbase; cn[q] = nev; cur[q] = 0; cf[q] = fl;
    }
    uint32_t inc = bk::warp_incl_scan(nev);
    if (q < K) cb[q] = run + inc - nev;
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  // events: `first` is resolved to a global node index here when it is a leaf (>= 0); a merged `first` is
  // stored as -1 - (its creating event) and resolved through ev_g once that event has its rank
  for (uint32_t q = lane; q < K; q += 32)
    for (uint32_t e = 0; e < cn[q]; ++e) {
      int32_t f = v.ev_first[cl[q] + e];
      ev_d[cb[q] + e] = v.ev_d[cl[q] + e];
      ev_first[cb[q] + e] = (uint32_t)f < cc[q] ? (int32_t)(v.comp_leaf[cl[q] + f] - pbase) : -1 - (int32_t)((uint32_t)f - cc[q]);
      ev_g[cb[q] + e] = -1;
    }
  __syncwarp();
  // flagged components start over in exact mode (warp-cooperative, one after the other)
  for (uint32_t q = 0; q < K; ++q)
    if (cf[q]) {
      CompCtx cx = make_cc(v, c0 + q);
      ahc_comp_init(v, cx);
      for (uint32_t j = 1; j < cx.c; ++j) ahc_find_best<true>(v, cx, j);
      ahc_comp_head(v, cx);
    }
  auto head_of = [&](uint32_t q) {
    if (cf[q]) {
      uint32_t comp = c0 + q;
      int32_t hj = v.comp_head_j[comp];
      double d = v.comp_head_d[comp];
      if (hj < 0 || !(d <= v.thr)) { hg[q] = -1; hd[q] = 0.0; return; }
      hd[q] = d;
      hg[q] = (uint32_t)hj < cc[q] ? (int32_t)(v.comp_leaf[cl[q] + hj] - pbase) : (int32_t)nleaf + v.node_grank[2 * cl[q] + hj];
      return;
    }
    uint32_t k = cur[q];
    if (k >= cn[q]) { hg[q] = -1; hd[q] = 0.0; return; }
    int32_t f = ev_first[cb[q] + k];
    hd[q] = ev_d[cb[q] + k];
    hg[q] = f >= 0 ? f : (int32_t)nleaf + ev_g[cb[q] + (uint32_t)(-1 - f)];
  };
  for (uint32_t q = lane; q < K; q += 32) head_of(q);
  __syncwarp();
  int32_t g = 0;
  while (true) {
    double bd = 0.0; int32_t bg = -1; uint32_t bq = 0;
    for (uint32_t q = lane; q < K; q += 32) {
      int32_t gi = hg[q];
      if (gi < 0) continue;
      double d = hd[q];
      if (bg < 0 || d < bd || (d == bd && gi > bg)) { bd = d; bg = gi; bq = q; }
    }
    for (int o = 16; o; o >>= 1) {
      double od = __shfl_xor_sync(0xffffffffu, bd, o);
      int32_t og = __shfl_xor_sync(0xffffffffu, bg, o);
      uint32_t oq = __shfl_xor_sync(0xffffffffu, bq, o);
      if (og >= 0 && (bg < 0 || od < bd || (od == bd && og > bg))) { bd = od; bg = og; bq = oq; }
    }
    if (bg < 0) break;
    if (cf[bq]) {
      CompCtx cx = make_cc(v, c0 + bq);
      ahc_merge<true>(v, cx, g);
      if (lane == 0) head_of(bq);
    } else if (lane == 0) { ev_g[cb[bq] + cur[bq]] = g; cur[bq] += 1; head_of(bq); }
    ++g;
    __syncwarp();
  }
  for (uint32_t q = lane; q < K; q += 32)
    for (uint32_t e = 0; e < cn[q]; ++e) v.node_grank[2 * cl[q] + cc[q] + e] = ev_g[cb[q] + e];
}

// Kernel B (rank form) for buckets without flagged components: the global pop order of the reference's
// merge loop is a merge of the per-component event lists, and an event pops before every event of another
// component whose PREFIX-MAX distance is larger (a component's later, smaller distances pop immediately
// after the event that set the prefix max).  So the global creation rank of an event is its rank under the
// key (prefix-max distance, component, index) -- computed by direct counting in shared memory, no
// sequential replay -- except inside groups of events from different components that share one exact
// prefix-max value; those (small, rare) groups are resolved by a literal heap walk over the group.
constexpr int RK_THREADS = 512;
__global__ void __launch_bounds__(RK_THREADS) ahc_replay_rank(AhcView v, const uint32_t *__restrict__ bucket_comp_off, uint32_t nb, const int32_t *__restrict__ bucket_flag,
                                                              uint32_t n_lo, uint32_t n_hi)
{
  uint32_t b = blockIdx.x;
  if (b >= nb || bucket_flag[b]) return;
  uint32_t nleaf = v.seg_off[b + 1] - v.seg_off[b];
  if (nleaf < n_lo || nleaf >= n_hi) return;
  extern __shared__ unsigned char dynsm[];
  __shared__ unsigned sh32[33];
  __shared__ unsigned sh_M;
  uint32_t c0 = bucket_comp_off[b], c1 = bucket_comp_off[b + 1], K = c1 - c0;
  double *ev_d = reinterpret_cast<double *>(dynsm);                    // [cap = nleaf]
  double *ev_pm = ev_d + nleaf;
  int32_t *ev_first = reinterpret_cast<int32_t *>(ev_pm + nleaf);
  uint32_t *ev_qi = reinterpret_cast<uint32_t *>(ev_first + nleaf);     // comp (high 16) | index (low 16)
  uint32_t *rank = ev_qi + nleaf;
  uint32_t *order = rank + nleaf;
  uint32_t *cb = order + nleaf, *cn = cb + nleaf, *cc = cn + nleaf, *cl = cc + nleaf;
  uint8_t *tie = reinterpret_cast<uint8_t *>(cl + nleaf);
  const uint32_t pbase = v.seg_off[b];
  unsigned run = 0;
  for (uint32_t q0 = 0; q0 < K; q0 += RK_THREADS) {
    uint32_t q = q0 + threadIdx.x;
    unsigned nev = 0;
    if (q < K) {
      uint32_t comp = c0 + q, lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase;
      nev = v.comp_nnodes[comp] - c;
      cc[q] = c; cl[q] = lbase; cn[q] = nev;
    }
    unsigned tot;
    unsigned ex = bk::block_excl_scan<unsigned>(nev, sh32, tot);
    if (q < K) cb[q] = run + ex;
    run += tot;
  }
  if (threadIdx.x == 0) sh_M = run;
  __syncthreads();
  const uint32_t M = sh_M;
  for (uint32_t q = threadIdx.x; q < K; q += RK_THREADS) {
    double pm = 0.0;
    for (uint32_t e = 0; e < cn[q]; ++e) {
      int32_t f = v.ev_first[cl[q] + e];
      double d = v.ev_d[cl[q] + e];
      if (d > pm) pm = d;
      uint32_t o = cb[q] + e;
      ev_d[o] = d; ev_pm[o] = pm;
      ev_first[o] = (uint32_t)f < cc[q] ? (int32_t)(v.comp_leaf[cl[q] + f] - pbase) : -1 - (int32_t)((uint32_t)f - cc[q]);
      ev_qi[o] = (q << 16) | e;
    }
  }
  __syncthreads();
  // rank by counting; events are stored in (component, index) order, so ev_qi is increasing with the slot
  for (uint32_t e = threadIdx.x; e < M; e += RK_THREADS) {
    double pm = ev_pm[e];
    uint32_t qe = ev_qi[e] >> 16, cnt = 0;
    bool t = false;
    for (uint32_t o = 0; o < M; ++o) {
      double p2 = ev_pm[o];
      cnt += (p2 < pm || (p2 == pm && o < e)) ? 1u : 0u;
      t |= (p2 == pm) && ((ev_qi[o] >> 16) != qe);
    }
    rank[e] = cnt; order[cnt] = e; tie[e] = t ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t r = 0;
    while (r < M) {
      uint32_t e0 = order[r];
      if (!tie[e0]) { ++r; continue; }
      uint32_t r2 = r + 1;
      while (r2 < M && ev_pm[order[r2]] == ev_pm[e0]) ++r2;
      // literal heap walk over the group [r, r2): heads = first unpopped event of each component's run
      for (uint32_t s = 0; s < r2 - r; ++s) {
        int best = -1; double bd = 0.0; int32_t bg = -1;
        for (uint32_t p = r; p < r2; ++p) {
          uint32_t e = order[p];
          if (tie[e] == 2) continue;
          if (p > r) { uint32_t ep = order[p - 1]; if ((ev_qi[ep] >> 16) == (ev_qi[e] >> 16) && tie[ep] != 2) continue; }
          int32_t f = ev_first[e];
          uint32_t q = ev_qi[e] >> 16;
          int32_t gi = f >= 0 ? f : (int32_t)nleaf + (int32_t)rank[cb[q] + (uint32_t)(-1 - f)];
          double d = ev_d[e];
          if (best < 0 || d < bd || (d == bd && gi > bg)) { best = (int)e; bd = d; bg = gi; }
        }
        rank[best] = r + s;
        tie[best] = 2;
      }
      r = r2;
    }
  }
  __syncthreads();
  for (uint32_t e = threadIdx.x; e < M; e += RK_THREADS) {
    uint32_t q = ev_qi[e] >> 16, i = ev_qi[e] & 0xffffu;
    v.node_grank[2 * cl[q] + cc[q] + i] = (int32_t)rank[e];
  }
}

// Kernel B (rank form, global memory) for buckets too large for shared memory: same algorithm as
// ahc_replay_rank, spread over the whole GPU (one thread per event slot for the counting pass).
struct RankGlobal {
  double *pm;            // [n] prefix-max distance per event slot, +inf for unused slots
  uint32_t *slot_comp;   // [n]
  int32_t *first;        // [n] leaf index (>= 0) or -1 - creating event
  uint32_t *rank, *order;
  uint8_t *tie;
  uint32_t *is_head;
};

__global__ void ahc_rg_prepare(AhcView v, RankGlobal g, uint32_t ncomp, const int32_t *__restrict__ bucket_flag, uint32_t n_lo)
{
  uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= ncomp) return;
  uint32_t b = v.comp_bucket[comp];
  if (bucket_flag[b] || v.seg_off[b + 1] - v.seg_off[b] < n_lo) return;
  uint32_t lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase, nev = v.comp_nnodes[comp] - c;
  const uint32_t pbase = v.seg_off[b];
  double pm = 0.0;
  for (uint32_t e = 0; e < c; ++e) {
    uint32_t q = lbase + e;
    g.slot_comp[q] = comp;
    if (e < nev) {
      double d = v.ev_d[q];
      if (d > pm) pm = d;
      int32_t f = v.ev_first[q];
      g.pm[q] = pm;
      g.first[q] = (uint32_t)f < c ? (int32_t)(v.comp_leaf[lbase + f] - pbase) : -1 - (int32_t)((uint32_t)f - c);
    } else g.pm[q] = __longlong_as_double(0x7ff0000000000000ll);
    g.tie[q] = 0; g.is_head[q] = 0;
  }
}

// The rank of an event under (prefix-max distance, slot) inside its bucket is its position after two stable radix
// sorts -- by the bit pattern of the (non-negative) prefix max, then by bucket -- instead of O(M^2) counting: buckets
// of tens of thousands of points (deep coverage, multi-GPU bucket owners) made the counting pass the longest kernel
// of the whole step.
__global__ void ahc_rg_sortkeys(AhcView v, RankGlobal g, const uint32_t *__restrict__ point_bucket, long long n, const int32_t *__restrict__ bucket_flag, uint32_t n_lo,
                                uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  uint32_t b = point_bucket[e];
  bool live = !bucket_flag[b] && v.seg_off[b + 1] - v.seg_off[b] >= n_lo;
  key[e] = live ? (uint64_t)__double_as_longlong(g.pm[e]) : 0x7ff0000000000000ull;
  val[e] = (uint32_t)e;
}
__global__ void ahc_rg_bucketkeys(const uint32_t *__restrict__ point_bucket, const uint32_t *__restrict__ val, long long n, uint64_t *__restrict__ key)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) key[p] = point_bucket[val[p]];
}
__global__ void __launch_bounds__(256) ahc_rg_rank_sorted(AhcView v, RankGlobal g, const uint32_t *__restrict__ point_bucket, long long n, const int32_t *__restrict__ bucket_flag,
                                                          uint32_t n_lo, const uint32_t *__restrict__ sorted)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t e = sorted[p];
  uint32_t b = point_bucket[e];
  uint32_t s0 = v.seg_off[b], s1 = v.seg_off[b + 1];
  if (bucket_flag[b] || s1 - s0 < n_lo) return;
  double pm = g.pm[e];
  if (!(pm < 1.0e300)) return;                         // unused slot (sorted behind every event of its bucket)
  uint32_t ce = g.slot_comp[e];
  bool t = false;
  for (long long q = p - 1; q >= (long long)s0 && !t; --q) { uint32_t o = sorted[q]; if (g.pm[o] != pm) break; t = g.slot_comp[o] != ce; }
  for (long long q = p + 1; q < (long long)s1 && !t; ++q) { uint32_t o = sorted[q]; if (g.pm[o] != pm) break; t = g.slot_comp[o] != ce; }
  g.rank[e] = (uint32_t)(p - s0); g.order[p] = e; g.tie[e] = t ? 1 : 0;
}

// group heads in rank order (a group = maximal run of one exact prefix-max value touching >= 2 components)
__global__ void ahc_rg_heads(AhcView v, RankGlobal g, const uint32_t *__restrict__ point_bucket, long long n, const int32_t *__restrict__ bucket_flag, uint32_t n_lo,
                             const uint32_t *__restrict__ bucket_events)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t b = point_bucket[p];
  uint32_t s0 = v.seg_off[b], s1 = v.seg_off[b + 1];
  if (bucket_flag[b] || s1 - s0 < n_lo) return;
  uint32_t r = (uint32_t)p - s0;
  if (r >= bucket_events[b]) return;
  uint32_t e = g.order[p];
  if (!g.tie[e]) return;
  if (r == 0 || g.pm[g.order[p - 1]] != g.pm[e]) g.is_head[p] = 1;
}

// Tie groups (one exact prefix-max value touching >= 2 components): the literal heap walk over the group decides the
// order of its events.  ONE WARP PER GROUP, groups handed out by an atomic ticket in rank order.  A group only depends on
// earlier groups through the rank of the event that created a candidate's first node (same component, smaller or equal
// prefix max): the warp waits for exactly those events (tie == 2) before it walks, everything else runs concurrently.
// Earlier groups hold earlier tickets, so whatever a warp waits for is already running.  Inside a group every step picks
// the minimum over the live candidates, which the lanes search in parallel (lexicographic warp reduction on (distance,
// -global index)); sqrt(dx^2+dy^2) of integer offsets takes few distinct values, so a bucket has hundreds of such groups.
// Measured on B200 before this form: one warp per BUCKET walking its groups in order took 0.65 ms (30x genome, global
// memory) / 0.30 ms (events staged in shared memory), and 3.6 - 6.7 ms per rank on the 8-GPU workload whose buckets no
// longer fit shared memory.
__global__ void __launch_bounds__(128) ahc_rg_ties(AhcView v, RankGlobal g, const uint32_t *__restrict__ point_bucket, const int32_t *__restrict__ bucket_flag, uint32_t n_lo,
                                                   const uint32_t *__restrict__ bucket_events, const uint32_t *__restrict__ head_pos, const unsigned long long *__restrict__ n_heads_p,
                                                   unsigned *__restrict__ ticket)
{
  const unsigned lane = threadIdx.x & 31;
  const unsigned n_heads = (unsigned)*n_heads_p;
  volatile uint8_t *tie = g.tie;
  volatile uint32_t *rank = g.rank;
  for (;;) {
    unsigned h = 0;
    if (lane == 0) h = atomicAdd(ticket, 1u);
    h = __shfl_sync(0xffffffffu, h, 0);
    if (h >= n_heads) return;
    const uint32_t p0 = head_pos[h], b = point_bucket[p0];
    const uint32_t s0 = v.seg_off[b], s1 = v.seg_off[b + 1];
    if (bucket_flag[b] || s1 - s0 < n_lo) continue;
    const uint32_t nleaf = s1 - s0, M = bucket_events[b], r = p0 - s0;
    const double pm0 = g.pm[g.order[p0]];
    // group end: first position whose prefix max differs (lanes probe 32 positions at a time)
    uint32_t r2 = r + 1;
    for (;;) {
      uint32_t q = r2 + lane;
      bool same = q < M && g.pm[g.order[s0 + q]] == pm0;
      unsigned m = __ballot_sync(0xffffffffu, same);
      if (m == 0xffffffffu) { r2 += 32; continue; }
      r2 += (uint32_t)__ffs(~m) - 1u;
      break;
    }
    // wait for the creators that sit in EARLIER groups (a creator inside this group is ranked by this walk before its
    // dependant becomes eligible: events of one component are taken in order)
    for (uint32_t p = r + lane; p < r2; p += 32) {
      uint32_t e = g.order[s0 + p];
      int32_t f = g.first[e];
      if (f >= 0) continue;
      uint32_t ce = v.comp_off[g.slot_comp[e]] + (uint32_t)(-1 - f);
      if (g.pm[ce] == pm0) continue;
      while (tie[ce] == 1) { }
    }
    __threadfence();
    __syncwarp();
    for (uint32_t s = 0; s < r2 - r; ++s) {
      long long best = -1; double bd = 0.0; int32_t bg = -1;
      for (uint32_t p = r + lane; p < r2; p += 32) {
        uint32_t e = g.order[s0 + p];
        if (tie[e] == 2) continue;
        if (p > r) { uint32_t ep = g.order[s0 + p - 1]; if (g.slot_comp[ep] == g.slot_comp[e] && tie[ep] != 2) continue; }
        int32_t f = g.first[e];
        uint32_t comp = g.slot_comp[e];
        int32_t gi = f >= 0 ? f : (int32_t)nleaf + (int32_t)rank[v.comp_off[comp] + (uint32_t)(-1 - f)];
        double d = v.ev_d[e];
        if (best < 0 || d < bd || (d == bd && gi > bg)) { best = (long long)e; bd = d; bg = gi; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        double od = __shfl_xor_sync(0xffffffffu, bd, o);
        int32_t og = __shfl_xor_sync(0xffffffffu, bg, o);
        if (ob >= 0 && (best < 0 || od < bd || (od == bd && og > bg))) { best = ob; bd = od; bg = og; }
      }
      if (lane == 0) { rank[best] = r + s; __threadfence(); tie[best] = 2; }
      __syncwarp();
    }
  }
}

__global__ void ahc_rg_head_list(const uint32_t *__restrict__ is_head, const uint32_t *__restrict__ head_excl, long long n, uint32_t *__restrict__ head_pos)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n && is_head[p]) head_pos[head_excl[p]] = (uint32_t)p;
}

__global__ void ahc_rg_bucket_events(AhcView v, uint32_t ncomp, uint32_t *__restrict__ bucket_events)
{
  uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= ncomp) return;
  uint32_t c = v.comp_off[comp + 1] - v.comp_off[comp], nev = v.comp_nnodes[comp] - c;
  if (nev) atomicAdd(&bucket_events[v.comp_bucket[comp]], nev);
}

__global__ void ahc_rg_write(AhcView v, RankGlobal g, const uint32_t *__restrict__ point_bucket, long long n, const int32_t *__restrict__ bucket_flag, uint32_t n_lo)
{
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  uint32_t b = point_bucket[e];
  if (bucket_flag[b] || v.seg_off[b + 1] - v.seg_off[b] < n_lo) return;
  if (!(g.pm[e] < 1.0e300)) return;
  uint32_t comp = g.slot_comp[e], lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase;
  v.node_grank[2 * lbase + c + ((uint32_t)e - lbase)] = (int32_t)g.rank[e];
}

// Kernel C: exact online form for buckets with a flagged component (one warp per bucket)
__global__ void __launch_bounds__(32) ahc_bucket_exact(AhcView v, const uint32_t *__restrict__ bucket_comp_off, uint32_t nb, const int32_t *__restrict__ bucket_flag, uint32_t n_lo)
{
  uint32_t b = blockIdx.x;
  if (b >= nb || !bucket_flag[b]) return;
  if (v.seg_off[b + 1] - v.seg_off[b] < n_lo) return;     // small buckets: flagged components are re-run inside ahc_replay_smem
  const unsigned lane = threadIdx.x & 31;
  uint32_t c0 = bucket_comp_off[b], c1 = bucket_comp_off[b + 1];
  uint32_t nleaf = v.seg_off[b + 1] - v.seg_off[b];
  for (uint32_t comp = c0; comp < c1; ++comp) {
    CompCtx cc = make_cc(v, comp);
    ahc_comp_init(v, cc);
    for (uint32_t j = 1; j < cc.c; ++j) ahc_find_best<true>(v, cc, j);
    ahc_comp_head(v, cc);
  }
  int32_t g = 0;
  while (true) {
    double bd = 1.7976931348623157e308; int32_t bg = -1; uint32_t bc = 0xffffffffu;
    for (uint32_t comp = c0 + lane; comp < c1; comp += 32) {
      int32_t hj = v.comp_head_j[comp];
      if (hj < 0) continue;
      uint32_t lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase;
      double d = v.comp_head_d[comp];
      int32_t gi = (uint32_t)hj < c ? (int32_t)(v.comp_leaf[lbase + hj] - v.seg_off[b]) : (int32_t)nleaf + v.node_grank[2 * lbase + hj];
      if (bg < 0 || d < bd || (d == bd && gi > bg)) { bd = d; bg = gi; bc = comp; }
    }
    for (int o = 16; o; o >>= 1) {
      double od = __shfl_xor_sync(0xffffffffu, bd, o);
      int32_t og = __shfl_xor_sync(0xffffffffu, bg, o);
      uint32_t oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (og >= 0 && (bg < 0 || od < bd || (od == bd && og > bg))) { bd = od; bg = og; bc = oc; }
    }
    if (bg < 0 || !(bd <= v.thr)) break;
    CompCtx cc = make_cc(v, bc);
    ahc_merge<true>(v, cc, g);
    ++g;
  }
}

// ---- component construction helpers -----------------------------------------------------------
__global__ void ahc_key_bx(const uint32_t *__restrict__ bucket_of, const uint32_t *__restrict__ X, long long np, uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < np) { key[p] = ((uint64_t)bucket_of[p] << 32) | X[p]; val[p] = (uint32_t)p; }
}
// head flags in sorted order: new group when the high word changes or the low-word gap exceeds thr
__global__ void ahc_gap_heads(const uint64_t *__restrict__ key, long long np, long long thr, uint32_t *__restrict__ head)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  unsigned h = 1;
  if (p > 0 && (key[p] >> 32) == (key[p - 1] >> 32)) {
    long long gap = (long long)(key[p] & 0xffffffffull) - (long long)(key[p - 1] & 0xffffffffull);
    h = gap > thr ? 1u : 0u;
  }
  head[p] = h;
}
// key2 = (group id << 32) | lowval[val]
__global__ void ahc_key_group(const uint32_t *__restrict__ head, const uint32_t *__restrict__ head_excl, const uint32_t *__restrict__ val,
                              const uint32_t *__restrict__ low, long long np, uint64_t *__restrict__ key)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  uint32_t g = head_excl[p] - (head[p] ? 0u : 1u);
  key[p] = ((uint64_t)g << 32) | (low ? low[val[p]] : val[p]);
}
// after the final sort by (comp, p): comp_leaf = val, comp heads where the high word changes
__global__ void ahc_comp_heads(const uint64_t *__restrict__ key, long long np, uint32_t *__restrict__ head)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < np) head[p] = (p == 0 || (key[p] >> 32) != (key[p - 1] >> 32)) ? 1u : 0u;
}
__global__ void ahc_comp_fill(const uint64_t *__restrict__ key, const uint32_t *__restrict__ head, const uint32_t *__restrict__ val, long long np,
                              const uint32_t *__restrict__ bucket_of, uint32_t *__restrict__ comp_off, uint32_t *__restrict__ comp_bucket,
                              uint32_t *__restrict__ comp_of_point)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  uint32_t comp = (uint32_t)(key[p] >> 32);
  comp_of_point[val[p]] = comp;
  if (head[p]) { comp_off[comp] = (uint32_t)p; comp_bucket[comp] = bucket_of[val[p]]; }
}
// per component: lo_oc per leaf slot, hi_oc, pool sizes; one thread per component
__global__ void ahc_comp_static(const uint32_t *__restrict__ comp_off, const uint32_t *__restrict__ comp_leaf, const uint32_t *__restrict__ comp_bucket,
                                const uint32_t *__restrict__ comp_of_point, const uint32_t *__restrict__ seg_off, uint32_t ncomp,
                                int32_t *__restrict__ lo_oc, int32_t *__restrict__ hi_oc, unsigned long long *__restrict__ pts_sz, unsigned long long *__restrict__ row_sz)
{
  uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= ncomp) return;
  uint32_t s = comp_off[comp], e = comp_off[comp + 1], c = e - s;
  uint32_t pb = seg_off[comp_bucket[comp]], pe = seg_off[comp_bucket[comp] + 1];
  for (uint32_t q = s; q < e; ++q) {
    int32_t leaf = (int32_t)(comp_leaf[q] - pb);
    if (q > s && comp_leaf[q - 1] + 1 == comp_leaf[q]) lo_oc[q] = lo_oc[q - 1];
    else lo_oc[q] = leaf - 1;
  }
  long long r = (long long)pe - 1;
  while (r >= (long long)pb && comp_of_point[r] == comp) --r;
  hi_oc[comp] = (int32_t)(r - (long long)pb);
  pts_sz[comp] = (unsigned long long)c * (c + 1) / 2 + 1;
  row_sz[comp] = (unsigned long long)c * (c > 0 ? c - 1 : 0) / 2 + 1;
}
__global__ void ahc_bucket_comp_off(const uint32_t *__restrict__ comp_bucket, uint32_t ncomp, uint32_t nb, uint32_t *__restrict__ bucket_comp_off)
{
  // comp ids ascend with bucket; bucket_comp_off[b] = first comp with bucket >= b
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  uint32_t lo = 0, hi = ncomp;
  while (lo < hi) { uint32_t m = (lo + hi) / 2; if (comp_bucket[m] < b) lo = m + 1; else hi = m; }
  bucket_comp_off[b] = lo;
}
__global__ void ahc_bucket_flags(const int32_t *__restrict__ comp_flag, const uint32_t *__restrict__ comp_bucket, uint32_t ncomp, int32_t *__restrict__ bucket_flag)
{
  uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp < ncomp && comp_flag[comp]) bucket_flag[comp_bucket[comp]] = 1;
}

// final clusters: merged roots; key = (bucket << 32) | grank
__global__ void ahc_final_roots(AhcView v, uint32_t ncomp, uint64_t *__restrict__ key, uint32_t *__restrict__ val /* node id */, unsigned *__restrict__ count,
                                unsigned *__restrict__ merges_per_bucket)
{
  uint32_t comp = blockIdx.x * blockDim.x + threadIdx.x;
  if (comp >= ncomp) return;
  uint32_t lbase = v.comp_off[comp], c = v.comp_off[comp + 1] - lbase, nn = v.comp_nnodes[comp];
  if (nn > c) atomicAdd(&merges_per_bucket[v.comp_bucket[comp]], nn - c);
  for (uint32_t j = c; j < nn; ++j)
    if (v.node_root[2 * lbase + j]) {
      unsigned o = atomicAdd(count, 1u);
      key[o] = ((uint64_t)v.comp_bucket[comp] << 32) | (uint32_t)v.node_grank[2 * lbase + j];
      val[o] = 2 * lbase + j;
    }
}
// per final root (sorted by bucket, grank): size + cluster id inside the bucket
__global__ void ahc_root_sizes(AhcView v, const uint64_t *__restrict__ key, const uint32_t *__restrict__ node, uint32_t nroot, uint32_t *__restrict__ size,
                               uint32_t *__restrict__ head)
{
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nroot) return;
  size[r] = v.node_npts[node[r]];
  head[r] = (r == 0 || (key[r] >> 32) != (key[r - 1] >> 32)) ? 1u : 0u;
}
__global__ void ahc_root_firsts(const uint32_t *__restrict__ head, uint32_t nroot, const uint64_t *__restrict__ key, uint32_t *__restrict__ bucket_first_root)
{
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nroot && head[r]) bucket_first_root[(uint32_t)(key[r] >> 32)] = r;
}
// emit members: one warp per final root
__global__ void __launch_bounds__(128) ahc_emit(AhcView v, const uint64_t *__restrict__ key, const uint32_t *__restrict__ node, const uint32_t *__restrict__ out_off,
                                                const uint32_t *__restrict__ bucket_first_root, uint32_t nroot,
                                                uint32_t *__restrict__ out_point, int32_t *__restrict__ out_cluster, uint32_t *__restrict__ out_bucket)
{
  uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nroot) return;
  const unsigned lane = threadIdx.x & 31;
  uint32_t b = (uint32_t)(key[r] >> 32);
  uint32_t nd = node[r], n = v.node_npts[nd], o = out_off[r];
  unsigned long long po = v.node_pts[nd];
  uint32_t pbase = v.seg_off[b];
  int32_t k = (int32_t)(r - bucket_first_root[b]);
  for (uint32_t i = lane; i < n; i += 32) {
    out_point[o + i] = pbase + v.pts_pool[po + i];
    out_cluster[o + i] = k;
    out_bucket[o + i] = b;
  }
}

// =============================================================================================
// K5b: -fast anchored-window sweeps (src/BreakID.cc:1046-1160).  One thread per bucket walks its
// (small, post-mask) segment; windows are anchored at their first element, the element at index n-1
// always closes the open window and is itself lost (the reference never flushes the last window).
// =============================================================================================
__global__ void fast_sweep(const uint32_t *__restrict__ cur, const uint32_t *__restrict__ seg_off, uint32_t nb, const uint32_t *__restrict__ coord,
                           double w, int min_reads, uint32_t *__restrict__ kout /* per pair id */, uint32_t *__restrict__ keep /* per position */)
{
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t s = seg_off[b], e = seg_off[b + 1];
  uint32_t n = e - s;
  for (uint32_t i = s; i < e; ++i) keep[i] = 0;
  if (n == 0) return;
  uint32_t k = 1, ws = 0;                    // current window = [ws, i)
  long long pre = coord[cur[s]];
  for (uint32_t i = 1; i < n; ++i) {
    if ((double)coord[cur[s + i]] <= (double)pre + w && i != n - 1) continue;
    if ((int)(i - ws) >= min_reads) {
      for (uint32_t j = ws; j < i; ++j) { kout[cur[s + j]] = k; keep[s + j] = 1; }
      ++k;
    }
    pre = coord[cur[s + i]];
    ws = i;
  }
}

__global__ void compact_write(const uint32_t *__restrict__ cur, const uint32_t *__restrict__ bucket_of, const uint32_t *__restrict__ keep,
                              const uint32_t *__restrict__ off, long long np, uint32_t *__restrict__ cur_out, uint32_t *__restrict__ bucket_of_out)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < np && keep[p]) { cur_out[off[p]] = cur[p]; bucket_of_out[off[p]] = bucket_of[p]; }
}

// final numbering (src/BreakID.cc:1129-1157): in p1 order the k1 groups are contiguous; inside a
// group count equal (k1,k2), keep ids seen >= min_reads times, number by first appearance (from 1).
__global__ void fast_number(const uint32_t *__restrict__ cur, const uint32_t *__restrict__ seg_off, uint32_t nb, const uint32_t *__restrict__ k1,
                            const uint32_t *__restrict__ k2, int min_reads, int32_t *__restrict__ cl /* per position, 0 = dropped */, uint32_t *__restrict__ keep,
                            int32_t *__restrict__ nclusters)
{
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  uint32_t s = seg_off[b], e = seg_off[b + 1];
  int32_t k = 0;
  uint32_t gs = s;
  while (gs < e) {
    uint32_t ge = gs + 1;
    while (ge < e && k1[cur[ge]] == k1[cur[gs]]) ++ge;
    for (uint32_t i = gs; i < ge; ++i) {
      uint32_t id = k2[cur[i]], first = i;
      int cnt = 0;
      for (uint32_t j = gs; j < ge; ++j)
        if (k2[cur[j]] == id) { if (cnt == 0) first = j; ++cnt; }
      if (cnt >= min_reads) {
        if (first == i) { ++k; cl[i] = k; } else cl[i] = cl[first];
        keep[i] = 1;
      } else { cl[i] = 0; keep[i] = 0; }
    }
    gs = ge;
  }
  nclusters[b] = k;
}

// =============================================================================================
// K6: per-cluster summary (src/BreakID.cc:297-352).  Members arrive grouped by (bucket, cluster);
// one thread per cluster (clusters are small; sums are exact integer sums so order is irrelevant).
// =============================================================================================
__global__ void k6_cluster_heads(const uint32_t *__restrict__ mb, const int32_t *__restrict__ mc, long long nm, uint32_t *__restrict__ head)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nm) head[p] = (p == 0 || mb[p] != mb[p - 1] || mc[p] != mc[p - 1]) ? 1u : 0u;
}
__global__ void k6_cluster_starts(const uint32_t *__restrict__ head, const uint32_t *__restrict__ head_excl, long long nm, uint32_t *__restrict__ start)
{
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nm && head[p]) start[head_excl[p]] = (uint32_t)p;
}
__global__ void k6_summarize(const bkid_pair *__restrict__ pairs, const uint32_t *__restrict__ mp /* pair id per member */, const int32_t *__restrict__ mc,
                             const uint32_t *__restrict__ start, uint32_t ncl, long long nm, double w, bkid_cluster_rec *__restrict__ out, uint32_t *__restrict__ keep)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  uint32_t s = start[c], e = (c + 1 < ncl) ? start[c + 1] : (uint32_t)nm;
  bkid_cluster_rec R;
  memset(&R, 0, sizeof R);
  const bkid_pair &f = pairs[mp[s]];
  R.bucket = f.bucket; R.id = mc[s]; R.p1_tid = f.p1_tid; R.p2_tid = f.p2_tid;
  unsigned long long s1 = 0, s2 = 0;
  uint32_t mn1 = 0xffffffffu, mx1 = 0, mn2 = 0xffffffffu, mx2 = 0;
  unsigned types = 0;
  for (uint32_t i = s; i < e; ++i) {
    const bkid_pair &p = pairs[mp[i]];
    s1 += p.p1_pos; s2 += p.p2_pos;
    mn1 = min(mn1, p.p1_pos); mx1 = max(mx1, p.p1_pos); mn2 = min(mn2, p.p2_pos); mx2 = max(mx2, p.p2_pos);
    if (p.p1_tid != p.p2_tid) types |= 1u;                                   // diff_chr (:231-253)
    else {
      if (p.p1_strand == '-' && p.p2_strand == '+') types |= 4u;             // absolute_reverse
      if (p.p1_strand == p.p2_strand) types |= 2u;                           // same_orientation
      if (p.p1_strand == '+' && p.p2_strand == '-') types |= 8u;             // default_orientation
    }
  }
  long long n = (long long)(e - s);
  R.n_discordant_pair = n;
  R.p1_mean_pos = (uint32_t)__ddiv_rn((double)s1, (double)n);              // :342-343
  R.p2_mean_pos = (uint32_t)__ddiv_rn((double)s2, (double)n);
  R.p1_min_pos = mn1; R.p1_max_pos = mx1; R.p2_min_pos = mn2; R.p2_max_pos = mx2;
  long long md = (long long)(R.p1_mean_pos - R.p2_mean_pos);                 // :345
  bool close = (R.p1_tid == R.p2_tid) && ((double)md <= 2.0 * w) && ((double)md >= -2.0 * w);   // :348
  int ft = 0;                                                                // :1888-1907 (overwrite order)
  if (types & 1u) ft = 1;
  if (types & 2u) ft = 2;
  if (types & 4u) ft = 3;
  if (types & 8u) ft = 4;
  R.fusion_type = ft;
  R.p1_exact_pos = 0xffffffffu; R.p2_exact_pos = -1;
  out[c] = R;
  keep[c] = close ? 0u : 1u;
}
