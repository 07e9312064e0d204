/* oracle/oracle.cc -- TEST INFRASTRUCTURE ONLY: CPU restatement of the BreakID hot path.
 *
 * Every function restates one reference function on the struct-of-arrays record batch the
 * product's C-ABI consumes, and cites the reference lines it follows (paths relative to
 * /root/reference).  PINNING: the reference ships no tests or golden vectors, so
 * this restatement is pinned against the reference ITSELF, compiled here from its own sources
 * (oracle/Makefile -> oracle/_ref/libbreakid_ref.so + BreakID_ref): tests/test_oracle_vs_ref.py
 * diffs every orc_* entry point against the corresponding ref_* wrapper (oracle/ref_shim.cc) in
 * this container, and the committed fixtures under tests/golden/ (made by
 * tests/golden/make_golden.py with the reference binary) pin it where /root/reference is absent.
 *
 * Two kinds of functions live here:
 *   orc_*        literal restatements (sorted linked lists, N x N matrix, std::sort ...).
 *   orc_model_*  CPU models of the *device formulations* (parallel introsort replay, AHC by
 *                components with the closed-form tie rule).  They exist so that the formulation
 *                is proven equal to the literal restatement / the reference on the CPU before it
 *                is trusted on the GPU.  The CUDA kernels in breakid_b200/csrc follow them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

enum { F_PAIRED = 0x1, F_PROPER = 0x2, F_UNMAP = 0x4, F_REVERSE = 0x10, F_SECONDARY = 0x100,
       F_QCFAIL = 0x200, F_DUP = 0x400 };

struct H128 {
  uint64_t lo, hi;
  bool operator==(const H128 &o) const { return lo == o.lo && hi == o.hi; }
};
struct H128Hash { size_t operator()(const H128 &h) const { return (size_t)(h.lo ^ (h.hi * 0x9E3779B97F4A7C15ULL)); } };

/* src/util_bam.cc:57-68 -- uint32 wrap-around sum of target_len[0..tid) + pos */
uint32_t genome_pos(const uint32_t *target_len, int tid, int32_t pos)
{
  uint32_t p = 0;
  for (int i = 0; i < tid; ++i) p += target_len[i];
  p += (uint32_t)pos;
  return p;
}

struct KeyIdx { uint32_t key; uint32_t idx; };
bool cmp_key(KeyIdx a, KeyIdx b) { return a.key < b.key; }   /* src/BreakID.h:170-183 shape */

}  // namespace

extern "C" {

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------
 * a1. insert-size statistics -- src/BreakID.cc:1909-1954.
 * The sd accumulator is a `long` that is converted to double, added to d*d and truncated back on
 * every element (src/BreakID.cc:1913,1944); sd = sqrt(total/(double)n) (population sd).
 * Also returns the integer intermediates so the device path can be compared exactly. */
void orc_insert_stats(long n, const uint16_t *flag, const int32_t *isize, double *mean, double *sd,
                      int64_t *sum_out, int64_t *count_out, int64_t *sd_total_out)
{
  const uint32_t filter = F_UNMAP | F_SECONDARY | F_QCFAIL | F_DUP;
  long total = 0, cnt = 0;
  for (long i = 0; i < n; ++i)
    if ((flag[i] & F_PAIRED) && (flag[i] & F_PROPER) && !(flag[i] & filter)) { total += abs(isize[i]); ++cnt; }
  double m = (double)total / (double)cnt;
  long sd_total = 0;
  for (long i = 0; i < n; ++i)
    if ((flag[i] & F_PAIRED) && (flag[i] & F_PROPER) && !(flag[i] & filter)) {
      double x = (double)abs(isize[i]);
      sd_total += (x - m) * (x - m);      /* long += double: convert, add, truncate */
    }
  *mean = m;
  *sd = sqrt(sd_total / (double)cnt);
  if (sum_out) *sum_out = total;
  if (count_out) *count_out = cnt;
  if (sd_total_out) *sd_total_out = sd_total;
}

/* src/BreakID.cc:103 -- one distance used by scan, mask, span and cluster stages */
double orc_dist(double mean, double sd, int times) { return times * sqrt(times) * (mean + 3 * sd); }

/* ------------------------------------------------------------------------------------------
 * a2/a3. discordant-pair scan -- src/BreakID.cc:1362-1515, src/util_bam.cc:7-47,57-68.
 * names[t] = header target name of tid t.  Output: pairs bucket by bucket in std::map<string>
 * order of "p1chr_p2chr", inside a bucket in emission order. */
long orc_scan(long n, const uint16_t *flag, const uint8_t *mapq, const int32_t *tid, const int32_t *pos,
              const int32_t *mtid, const int32_t *mpos, const uint64_t *name_hash /* [2n] lo,hi */,
              int n_targets, const uint32_t *target_len, const char *const *names,
              long qual, double w, orc_pair **out)
{
  std::unordered_map<H128, long, H128Hash> store;      /* readname_2_alignment (:1379) */
  std::vector<orc_pair> emitted;
  for (long i = 0; i < n; ++i) {
    /* :1419-1420 -- SUPPLEMENTARY, UNMAP, QCFAIL are NOT tested */
    if (!((long)mapq[i] >= qual && !(flag[i] & F_DUP) && !(flag[i] & F_SECONDARY) &&
          (flag[i] & F_PAIRED) && !(flag[i] & F_PROPER)))
      continue;
    H128 h{name_hash[2 * i], name_hash[2 * i + 1]};
    auto it = store.find(h);
    if (it == store.end()) { store[h] = i; continue; }       /* first mate (:1485-1494) */
    long j = it->second;
    int ti = tid[i] < 0 ? -1 : tid[i], tj = tid[j] < 0 ? -1 : tid[j];   /* rname "*" for tid<0 */
    long pi = (long)pos[i] + 1, pj = (long)pos[j] + 1;
    if (ti != tj || (double)labs(pi - pj) >= w) {             /* :1428 */
      uint32_t c1 = genome_pos(target_len, tid[i], pos[i]);   /* :1431 -- the CURRENT record's own fields */
      uint32_t c2 = genome_pos(target_len, mtid[i], mpos[i]); /* :1432 -- and its MATE fields           */
      orc_pair p;
      memset(&p, 0, sizeof p);
      p.name_lo = h.lo; p.name_hi = h.hi;
      if (c1 <= c2) {                                          /* :1434-1448 */
        p.p1_flag = flag[i]; p.p1_tid = ti; p.p1_pos = (uint32_t)pi; p.p1_mapq = mapq[i];
        p.p1_chr_pos = c1; p.p2_chr_pos = c2;
        p.p2_flag = flag[j]; p.p2_tid = tj; p.p2_pos = (uint32_t)pj; p.p2_mapq = mapq[j];
      } else {                                                 /* :1449-1465 */
        p.p2_flag = flag[i]; p.p2_tid = ti; p.p2_pos = (uint32_t)pi; p.p2_mapq = mapq[i];
        p.p1_chr_pos = c2; p.p2_chr_pos = c1;
        p.p1_flag = flag[j]; p.p1_tid = tj; p.p1_pos = (uint32_t)pj; p.p1_mapq = mapq[j];
      }
      p.p1_strand = (p.p1_flag & F_REVERSE) ? '-' : '+';       /* :1467-1478 */
      p.p2_strand = (p.p2_flag & F_REVERSE) ? '-' : '+';
      p.cluster = -1;
      emitted.push_back(p);
    }
    store.erase(it);                                           /* :1482 -- always */
  }
  /* :1500-1512 -- bucket by "p1chr_p2chr", std::map<string> iteration order */
  auto nm = [&](int t) { return t < 0 ? std::string("*") : std::string(names[t]); };
  std::map<std::string, std::vector<long>> buckets;
  for (size_t k = 0; k < emitted.size(); ++k)
    buckets[nm(emitted[k].p1_tid) + "_" + nm(emitted[k].p2_tid)].push_back((long)k);
  orc_pair *o = (orc_pair *)calloc(emitted.size() ? emitted.size() : 1, sizeof(orc_pair));
  long k = 0; int b = 0;
  for (auto &kv : buckets) {
    for (long e : kv.second) { o[k] = emitted[e]; o[k].bucket = b; o[k].orig = (uint32_t)k; ++k; }
    ++b;
  }
  *out = o;
  return (long)emitted.size();
}

/* rank of every possible "chrA_chrB" bucket name in std::map<string> order (what the product's
 * host computes once per header and hands to the device).  rank[(a+1)*(n_targets+1)+(b+1)],
 * index 0 = "*". */
void orc_bucket_rank_table(int n_targets, const char *const *names, int32_t *rank)
{
  int m = n_targets + 1;
  std::vector<std::pair<std::string, int>> v;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b) {
      std::string sa = a ? names[a - 1] : "*", sb = b ? names[b - 1] : "*";
      v.push_back({sa + "_" + sb, a * m + b});
    }
  std::sort(v.begin(), v.end());
  for (size_t r = 0; r < v.size(); ++r) rank[v[r].second] = (int32_t)r;
}

/* ------------------------------------------------------------------------------------------
 * A2. the unstable std::sort the pipeline depends on (src/BreakID.cc:1274,1278,1282,144,1091,
 * 1127): literal = libstdc++ std::sort with a key-only comparator.  perm[i] = original index. */
void orc_std_sort_perm(long n, const uint32_t *key, uint32_t *perm)
{
  std::vector<KeyIdx> v(n);
  for (long i = 0; i < n; ++i) v[i] = KeyIdx{key[i], (uint32_t)i};
  std::sort(v.begin(), v.end(), cmp_key);
  for (long i = 0; i < n; ++i) perm[i] = v[i].idx;
}

/* MODEL of the device formulation of the same sort ("parallel introsort replay"):
 * libstdc++ introsort = recursive {median-of-3 to front, unguarded Hoare partition} until a
 * segment has <= 16 elements or the depth budget 2*floor(log2 n) is spent (then heapsort), followed
 * by one insertion sort over the whole array.  Facts used:
 *  (1) a partition step is data-parallel: with L = positions (ascending) holding key >= pivot and
 *      R = positions (descending) holding key <= pivot inside [first+1,last), the sequential
 *      pointer walk swaps L[k] <-> R[k] for k < K = #{k : L[k] < R[k]} and returns
 *      cut = min(L[K], R[K-1]) (R[-1] = last);
 *  (2) the final insertion sort is stable and never moves an element across a partition cut, so
 *      it equals an independent stable sort of every terminal segment.
 * So the permutation is: level-synchronous partitions over all segments, then a stable sort of
 * each terminal (<=16) segment.  Depth-exhausted segments fall back to the literal heapsort. */
static long g_model_heapsorts = 0;
static void model_heapsort(KeyIdx *a, long n)
{
  ++g_model_heapsorts;
  /* std::__partial_sort(first,last,last) = __heap_select (make_heap; nothing to select) +
   * __sort_heap -- libstdc++ bits/stl_heap.h semantics */
  auto adjust = [&](long hole, long len, KeyIdx value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (a[child].key < a[child - 1].key) child--;
      a[hole] = a[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      a[hole] = a[child - 1];
      hole = child - 1;
    }
    long parent = (hole - 1) / 2;                      /* __push_heap */
    while (hole > top && a[parent].key < value.key) {
      a[hole] = a[parent];
      hole = parent;
      parent = (hole - 1) / 2;
    }
    a[hole] = value;
  };
  if (n < 2) return;
  for (long parent = (n - 2) / 2;; --parent) {         /* __make_heap */
    KeyIdx v = a[parent];
    adjust(parent, n, v);
    if (parent == 0) break;
  }
  for (long last = n - 1; last > 0; --last) {          /* __sort_heap / __pop_heap */
    KeyIdx v = a[last];
    a[last] = a[0];
    adjust(0, last, v);
  }
}

long orc_model_heapsort_count(void) { return g_model_heapsorts; }

void orc_model_sort_perm(long n, const uint32_t *key, uint32_t *perm)
{
  std::vector<KeyIdx> a(n);
  for (long i = 0; i < n; ++i) a[i] = KeyIdx{key[i], (uint32_t)i};
  struct Seg { long f, l; int depth; };
  std::vector<Seg> cur, nxt, terminal;
  if (n > 1) {
    int lg = 0;
    for (long t = n; t > 1; t >>= 1) ++lg;              /* std::__lg(n) */
    cur.push_back({0, n, 2 * lg});
  }
  std::vector<long> L, R;
  while (!cur.empty()) {
    nxt.clear();
    for (const Seg &s : cur) {
      long f = s.f, l = s.l;
      if (l - f <= 16) { terminal.push_back(s); continue; }
      if (s.depth == 0) { model_heapsort(&a[f], l - f); continue; }
      /* __move_median_to_first(first, first+1, mid, last-1) */
      long mid = f + (l - f) / 2, A = f + 1, B = mid, C = l - 1;
      uint32_t ka = a[A].key, kb = a[B].key, kc = a[C].key;
      long med;
      if (ka < kb) { if (kb < kc) med = B; else if (ka < kc) med = C; else med = A; }
      else if (ka < kc) med = A;
      else if (kb < kc) med = C;
      else med = B;
      std::swap(a[f], a[med]);
      uint32_t p = a[f].key;
      /* data-parallel form of __unguarded_partition(first+1, last, first) */
      L.clear(); R.clear();
      for (long i = f + 1; i < l; ++i) if (!(a[i].key < p)) L.push_back(i);
      for (long i = l - 1; i > f; --i) if (!(p < a[i].key)) R.push_back(i);
      size_t K = 0;
      while (K < L.size() && K < R.size() && L[K] < R[K]) ++K;
      for (size_t k = 0; k < K; ++k) std::swap(a[L[k]], a[R[k]]);
      long rprev = K ? R[K - 1] : l;
      long cut = (K < L.size() && L[K] < rprev) ? L[K] : rprev;
      nxt.push_back({cut, l, s.depth - 1});
      nxt.push_back({f, cut, s.depth - 1});
    }
    cur.swap(nxt);
  }
  if (n <= 16 && n > 0) terminal.push_back({0, n, 0});
  for (const Seg &s : terminal)                          /* stable sort of each terminal segment */
    std::stable_sort(a.begin() + s.f, a.begin() + s.l, cmp_key);
  for (long i = 0; i < n; ++i) perm[i] = a[i].idx;
}

/* ------------------------------------------------------------------------------------------
 * a4. isolated-pair removal -- src/BreakID.cc:1271-1285, mask_pairs_chr_pos :1813-1877.
 * Works on indices; returns surviving original indices in final order (may contain the index
 * that sat at position 1 twice; never positions 0 and np-1 of either pass). */
static void mask_pass(std::vector<KeyIdx> &v, const uint32_t *p1, const uint32_t *p2, long distance)
{
  long np = (long)v.size();
  if (np <= 2) { v.clear(); return; }
  auto gap = [](uint32_t a, uint32_t b) { return (long)abs((int32_t)(a - b)); };
  std::vector<KeyIdx> out;
  long Lx = gap(p1[v[1].idx], p1[v[2].idx]), Ly = gap(p2[v[1].idx], p2[v[2].idx]);
  if (!(Lx > distance || Ly > distance)) out.push_back(v[1]);            /* :1830-1835 */
  for (long i = 1; i < np - 1; ++i) {                                    /* :1845-1870 */
    long ll = gap(p1[v[i - 1].idx], p1[v[i].idx]), lr = gap(p1[v[i + 1].idx], p1[v[i].idx]);
    Lx = ll < lr ? ll : lr;
    ll = gap(p2[v[i - 1].idx], p2[v[i].idx]); lr = gap(p2[v[i + 1].idx], p2[v[i].idx]);
    Ly = ll < lr ? ll : lr;
    if (!(Lx > distance || Ly > distance)) out.push_back(v[i]);
  }
  v.swap(out);
}

long orc_remove_isolated(long n, const uint32_t *p1, const uint32_t *p2, double w, uint32_t *out_idx)
{
  std::vector<KeyIdx> v(n);
  for (long i = 0; i < n; ++i) v[i] = KeyIdx{p1[i], (uint32_t)i};
  std::sort(v.begin(), v.end(), cmp_key);                                /* :1274 */
  mask_pass(v, p1, p2, (long)w);                                         /* double -> long at the call */
  if (!v.empty()) {
    for (auto &e : v) e.key = p2[e.idx];
    std::sort(v.begin(), v.end(), cmp_key);                              /* :1278 */
    mask_pass(v, p1, p2, (long)w);
    if (!v.empty()) {
      for (auto &e : v) e.key = p1[e.idx];
      std::sort(v.begin(), v.end(), cmp_key);                            /* :1282 */
    }
  }
  for (size_t i = 0; i < v.size(); ++i) out_idx[i] = v[i].idx;
  return (long)v.size();
}

/* ------------------------------------------------------------------------------------------
 * a6. -fast clustering -- src/BreakID.cc:1046-1160.  Input order = order after a4.
 * Output: surviving original indices + cluster id (1-based, first appearance in p1 order). */
long orc_cluster_fast(long n, const uint32_t *p1, const uint32_t *p2, double w,
                      uint32_t *out_idx, int32_t *out_cluster, int *roots)
{
  struct E { uint32_t idx; int k1, k2; };
  std::vector<E> e(n), tmp;
  for (long i = 0; i < n; ++i) e[i] = E{(uint32_t)i, 0, 0};
  const size_t min_reads = 2;
  auto sweep = [&](const uint32_t *pp, bool second) {
    tmp.clear();
    long m = (long)e.size();
    if (m == 0) return;                     /* the reference reads enspan[0] of an empty vector here (:1095) */
    std::vector<long> cl{0};
    int k = 1;
    long pre = pp[e[0].idx];
    for (long i = 1; i < m; ++i) {
      if ((double)pp[e[i].idx] <= (double)pre + w && i != m - 1) cl.push_back(i);   /* :1064,:1100 */
      else {
        if (cl.size() >= min_reads) {
          for (long j : cl) { if (second) e[j].k2 = k; else e[j].k1 = k; tmp.push_back(e[j]); }
          ++k;
        }
        pre = pp[e[i].idx];
        cl.clear(); cl.push_back(i);
      }
    }
    e = tmp;
  };
  auto sort_by = [&](const uint32_t *pp) {
    std::vector<KeyIdx> v(e.size());
    for (size_t i = 0; i < e.size(); ++i) v[i] = KeyIdx{pp[e[i].idx], (uint32_t)i};
    std::sort(v.begin(), v.end(), cmp_key);
    std::vector<E> r(e.size());
    for (size_t i = 0; i < e.size(); ++i) r[i] = e[v[i].idx];
    e.swap(r);
  };
  sweep(p1, false);
  sort_by(p2);                                                             /* :1091 */
  sweep(p2, true);
  sort_by(p1);                                                             /* :1127 */
  std::map<std::pair<int, int>, int> cnt, cl;                              /* key / key_cl (:1129-1157) */
  for (auto &x : e) cnt[{x.k1, x.k2}]++;
  int k = 0; long m = 0;
  for (auto &x : e) {
    if (cnt[{x.k1, x.k2}] >= (int)min_reads) {
      auto it = cl.find({x.k1, x.k2});
      int id;
      if (it == cl.end()) { id = ++k; cl[{x.k1, x.k2}] = id; } else id = it->second;
      out_idx[m] = x.idx; out_cluster[m] = id; ++m;
    }
  }
  *roots = k;
  return m;
}

/* ------------------------------------------------------------------------------------------
 * a5. AHC, literal -- src/util_cluster.cc:7-396 + src/BreakID.cc:1304-1352.
 * Average linkage (distance_type hard-wired to 1, src/BreakID.cc:33,135). */
namespace {
struct Nb { int target; double d; Nb *prev, *next; };
struct LNode { int is_root; std::vector<int> pts; Nb *nbs; int ma, mb; };
struct Lit {
  long N; std::vector<LNode> nodes; int roots; std::vector<double> mat; std::vector<Nb *> pool;
  double m(int a, int b) const { return mat[(size_t)a * N + b]; }
  double dist(int cur, int tgt) const {                                     /* get_distance :158-199 */
    if (cur < N && tgt < N) return m(cur, tgt);
    const std::vector<int> &a = nodes[cur].pts, &b = nodes[tgt].pts;
    double total = 0.0;                                                      /* average_linkage :201-215 */
    for (size_t i = 0; i < a.size(); ++i) for (size_t j = 0; j < b.size(); ++j) total += m(a[i], b[j]);
    return total / (int)(a.size() * b.size());
  }
  void insert_sorted(Nb *cn, int cur) {                                      /* :249-297 */
    Nb *t = nodes[cur].nbs;
    auto before = [&](Nb *tt) {
      cn->next = tt;
      if (tt->prev) { tt->prev->next = cn; cn->prev = tt->prev; } else nodes[cur].nbs = cn;
      tt->prev = cn;
    };
    while (t->next) { if (t->d >= cn->d) { before(t); return; } t = t->next; }
    if (t->d > cn->d) before(t); else { cn->prev = t; t->next = cn; }
  }
  void update_neighbours() {                                                 /* :112-134 */
    int cur = (int)nodes.size() - 1, seen = 1, tgt = cur;
    while (seen < roots) {
      --tgt;
      if (nodes[tgt].is_root) {
        ++seen;
        Nb *cn = (Nb *)calloc(1, sizeof(Nb)); pool.push_back(cn);
        cn->target = tgt; cn->d = dist(cur, tgt);
        if (nodes[cur].nbs) insert_sorted(cn, cur); else nodes[cur].nbs = cn;
      }
    }
  }
};
}  // namespace

/* returns number of nodes; fills is_root / merged children (size 2n) */
long orc_ahc_tree(long n, const double *x, const double *y, long thr,
                  int32_t *is_root, int32_t *merged_a, int32_t *merged_b)
{
  Lit c; c.N = n; c.roots = 0;
  c.mat.resize((size_t)n * n);
  for (long i = 0; i < n; ++i) for (long j = 0; j < n; ++j) {
    double dx = x[i] - x[j], dy = y[i] - y[j];
    c.mat[(size_t)i * n + j] = sqrt(dx * dx + dy * dy);                      /* :79-84, pow(.,2) folds to a product */
  }
  for (long i = 0; i < n; ++i) {                                             /* add_leaves :86-110 */
    c.nodes.push_back(LNode{1, {(int)i}, nullptr, -1, -1}); c.roots++;
    c.update_neighbours();
  }
  while (c.roots > 1) {                                                      /* merge_clusters :299-318 */
    double best = DBL_MAX; int first = -1, second = 0;
    int seen = 0, j = (int)c.nodes.size();
    while (seen < c.roots) {                                                 /* :320-334 */
      --j;
      if (!c.nodes[j].is_root) continue;
      ++seen;
      for (Nb *q = c.nodes[j].nbs; q; q = q->next)                           /* :337-355 */
        if (c.nodes[q->target].is_root) {
          if (first == -1 || q->d < best) { first = j; second = q->target; best = q->d; }
          break;
        }
    }
    if (first != -1 && best <= (double)thr) {                                /* :308 */
      LNode nn{1, {}, nullptr, first, second};                               /* merge :357-396 */
      c.nodes[first].is_root = 0; c.nodes[second].is_root = 0;
      nn.pts = c.nodes[first].pts;
      nn.pts.insert(nn.pts.end(), c.nodes[second].pts.begin(), c.nodes[second].pts.end());
      c.nodes.push_back(nn); c.roots--;
      c.update_neighbours();
    } else break;
  }
  long nn = (long)c.nodes.size();
  for (long i = 0; i < nn; ++i) { is_root[i] = c.nodes[i].is_root; merged_a[i] = c.nodes[i].ma; merged_b[i] = c.nodes[i].mb; }
  for (Nb *p : c.pool) free(p);
  return nn;
}

/* from a merge tree to the reference's pair re-emission (src/BreakID.cc:1328-1352): roots with
 * >= 2 points in node order get k = 0,1,..; points in `points` order (first.points ++ second.points) */
static long tree_to_clusters(long n, long nn, const int32_t *is_root, const int32_t *ma, const int32_t *mb,
                             uint32_t *out_idx, int32_t *out_cluster, int *roots)
{
  std::vector<std::vector<int>> pts(nn);
  for (long i = 0; i < nn; ++i) {
    if (i < n) pts[i] = {(int)i};
    else { pts[i] = pts[ma[i]]; pts[i].insert(pts[i].end(), pts[mb[i]].begin(), pts[mb[i]].end()); }
  }
  long m = 0; int k = 0, r = 0;
  for (long i = 0; i < nn; ++i) {
    if (is_root[i]) ++r;
    if (is_root[i] && pts[i].size() >= 2) {
      for (int p : pts[i]) { out_idx[m] = (uint32_t)p; out_cluster[m] = k; ++m; }
      ++k;
    }
  }
  *roots = r;                                                                 /* print_root_nodes :398-417 */
  return m;
}

long orc_cluster_ahc(long n, const uint32_t *p1, const uint32_t *p2, double thr,
                     uint32_t *out_idx, int32_t *out_cluster, int *roots)
{
  std::vector<double> x(n), y(n);
  for (long i = 0; i < n; ++i) { x[i] = p1[i]; y[i] = p2[i]; }             /* build_pair_array :1795-1806 */
  std::vector<int32_t> r(2 * n + 1), a(2 * n + 1), b(2 * n + 1);
  long nn = orc_ahc_tree(n, x.data(), y.data(), (long)thr, r.data(), a.data(), b.data());
  return tree_to_clusters(n, nn, r.data(), a.data(), b.data(), out_idx, out_cluster, roots);
}

/* ------------------------------------------------------------------------------------------
 * MODEL of the device AHC formulation (SURVEY.md App. A4 + the closed-form tie rule below).
 *
 *  * Components: sort by x, cut at x-gaps > thr; inside a piece sort by y, cut at y-gaps > thr.
 *    Points of different components are farther apart than thr, so no linkage across components
 *    can ever be <= thr and only in-component neighbour rows are kept.
 *  * Node j keeps a ROW D(j,t) over the in-component roots t < j that existed when j was created
 *    (leaf rows are recomputed from coordinates).  The reference's sorted list is replaced by:
 *      winner(j) = among still-root t with minimal D: let T = {t : D(j,t) == dmin} over the whole
 *      row (root or not), t1 > t2 its two highest indices.  The list holds T in ascending index
 *      order, except that t1 precedes t2 iff, when t2 was inserted, t1 was the list tail:
 *      every row entry with index in (t2, j) other than t1 has D < dmin AND no out-of-component
 *      root with index in (t2, j) existed when j was created ("sentinel", insert_sorted's tail
 *      exception src/util_cluster.cc:266-273).  So winner = min(S) unless min(S)==t2 && t1 in S &&
 *      exception -> t1.
 *  * Sentinel arithmetic: every merged node stores its global creation rank g and its rank l
 *      inside its component.  For merged j and t2: if t2 is merged, an out-of-component root in
 *      (t2,j) exists iff (g_j-g_t2) > (l_j-l_t2); if t2 is a leaf it exists iff g_j > l_j (some
 *      other component has merged before) or an out-of-component leaf index > t2 exists.
 *      For leaf j: static -- an out-of-component leaf with index in (t2, j).
 *  * Global loop: repeatedly take the component whose head event is minimal (ties -> higher global
 *      index of `first`), stop when it exceeds thr.
 * Output identical in shape to orc_ahc_tree (global node numbering).
 */
namespace {
struct MNode {
  int is_root, comp, lrank;       /* lrank: creation rank inside the component (merged only) */
  long grank;                     /* global creation rank (merged only) */
  int gidx;                       /* global node index */
  std::vector<int> pts;           /* global leaf indices */
  std::vector<int> row_t;         /* local target ids (descending) */
  std::vector<double> row_d;
  int best_t; double best_d;      /* cached winner (local id) or -1 */
};
struct Comp {
  std::vector<int> leaves;        /* global leaf indices ascending = local ids 0..c-1 */
  std::vector<MNode> nodes;       /* local id order */
  int merges;
  int hi_oc_leaf;                 /* highest global leaf index NOT in this component (-1 if none) */
  int head_j; double head_d;      /* current best event */
};
}  // namespace

static long g_model_exc_taken = 0, g_model_exc_blocked = 0;
void orc_model_ahc_counters(long *taken, long *blocked) { *taken = g_model_exc_taken; *blocked = g_model_exc_blocked; }

long orc_model_ahc_tree(long n, const double *x, const double *y, long thr,
                        int32_t *is_root, int32_t *merged_a, int32_t *merged_b)
{
  /* ---- components ---- */
  std::vector<int> comp_of(n, -1);
  std::vector<Comp> comps;
  {
    std::vector<int> ox(n);
    for (long i = 0; i < n; ++i) ox[i] = (int)i;
    std::stable_sort(ox.begin(), ox.end(), [&](int a, int b) { return x[a] < x[b]; });
    long s = 0;
    while (s < n) {
      long e = s + 1;
      while (e < n && !(x[ox[e]] - x[ox[e - 1]] > (double)thr)) ++e;
      std::vector<int> oy(ox.begin() + s, ox.begin() + e);
      std::stable_sort(oy.begin(), oy.end(), [&](int a, int b) { return y[a] < y[b]; });
      size_t s2 = 0;
      while (s2 < oy.size()) {
        size_t e2 = s2 + 1;
        while (e2 < oy.size() && !(y[oy[e2]] - y[oy[e2 - 1]] > (double)thr)) ++e2;
        Comp c; c.merges = 0; c.head_j = -1; c.head_d = DBL_MAX; c.hi_oc_leaf = -1;
        c.leaves.assign(oy.begin() + s2, oy.begin() + e2);
        std::sort(c.leaves.begin(), c.leaves.end());
        for (int l : c.leaves) comp_of[l] = (int)comps.size();
        comps.push_back(c);
        s2 = e2;
      }
      s = e;
    }
  }
  auto euclid = [&](int a, int b) { double dx = x[a] - x[b], dy = y[a] - y[b]; return sqrt(dx * dx + dy * dy); };
  /* lo_oc[j] = highest leaf index < j not in j's component (-1 if none) */
  std::vector<int> lo_oc(n, -1);
  for (long j = 0; j < n; ++j) { long r = j - 1; while (r >= 0 && comp_of[r] == comp_of[j]) --r; lo_oc[j] = (int)r; }
  for (size_t ci = 0; ci < comps.size(); ++ci) {
    Comp &c = comps[ci];
    long r = n - 1; while (r >= 0 && comp_of[r] == (int)ci) --r;
    c.hi_oc_leaf = (int)r;
    for (size_t k = 0; k < c.leaves.size(); ++k) {
      MNode nd; nd.is_root = 1; nd.comp = (int)ci; nd.lrank = -1; nd.grank = -1; nd.gidx = c.leaves[k];
      nd.pts = {c.leaves[k]}; nd.best_t = -1; nd.best_d = 0;
      c.nodes.push_back(nd);
    }
  }
  long total_merges = 0;
  /* distance from node j (local) to target t (local): row lookup or recompute */
  auto rowd = [&](Comp &c, int j, int t) -> double {
    int cl = (int)c.leaves.size();
    if (j < cl) return euclid(c.leaves[j], c.leaves[t]);
    MNode &nj = c.nodes[j];
    for (size_t k = 0; k < nj.row_t.size(); ++k) if (nj.row_t[k] == t) return nj.row_d[k];
    return DBL_MAX;
  };
  /* winner among still-root targets of node j */
  auto find_best = [&](Comp &c, int j) {
    int cl = (int)c.leaves.size();
    MNode &nj = c.nodes[j];
    std::vector<int> ts;                    /* row targets (roots at creation), any order */
    if (j < cl) { for (int t = j - 1; t >= 0; --t) ts.push_back(t); }
    else ts = nj.row_t;
    double dmin = DBL_MAX; int tmin = -1;
    for (int t : ts) if (c.nodes[t].is_root) { double d = rowd(c, j, t); if (d < dmin || (d == dmin && t < tmin)) { dmin = d; tmin = t; } }
    nj.best_t = tmin; nj.best_d = dmin;
    if (tmin < 0) return;
    int t1 = -1, t2 = -1;                   /* two highest indices in the row with D == dmin */
    for (int t : ts) if (rowd(c, j, t) == dmin) { if (t > t1) { t2 = t1; t1 = t; } else if (t > t2) t2 = t; }
    if (t2 < 0 || tmin != t2 || !c.nodes[t1].is_root) return;
    /* in-component part of the tail test */
    for (int t : ts) if (t > t2 && t != t1 && !(rowd(c, j, t) < dmin)) return;
    /* out-of-component sentinel */
    bool sentinel;
    if (j < cl) sentinel = lo_oc[c.leaves[j]] > c.leaves[t2];
    else if (t2 >= cl) sentinel = (nj.grank - c.nodes[t2].grank) > (long)(nj.lrank - c.nodes[t2].lrank);
    else sentinel = (nj.grank > (long)nj.lrank) || (c.hi_oc_leaf > c.leaves[t2]);
    if (!sentinel) { nj.best_t = t1; ++g_model_exc_taken; } else ++g_model_exc_blocked;
  };
  auto comp_head = [&](Comp &c) {
    c.head_j = -1; c.head_d = DBL_MAX;
    for (int j = (int)c.nodes.size() - 1; j >= 0; --j) {
      MNode &nj = c.nodes[j];
      if (!nj.is_root || nj.best_t < 0) continue;
      if (c.head_j == -1 || nj.best_d < c.head_d) { c.head_j = j; c.head_d = nj.best_d; }
    }
  };
  for (Comp &c : comps) {
    for (size_t j = 0; j < c.nodes.size(); ++j) find_best(c, (int)j);
    comp_head(c);
  }
  /* ---- global loop ---- */
  struct GNode { int a, b, root; };
  std::vector<GNode> g(n, GNode{-1, -1, 1});
  for (;;) {
    int bc = -1; double bd = DBL_MAX; int bg = -1;
    for (size_t ci = 0; ci < comps.size(); ++ci) {
      Comp &c = comps[ci];
      if (c.head_j < 0) continue;
      int gi = c.nodes[c.head_j].gidx;
      if (bc < 0 || c.head_d < bd || (c.head_d == bd && gi > bg)) { bc = (int)ci; bd = c.head_d; bg = gi; }
    }
    if (bc < 0 || !(bd <= (double)thr)) break;
    Comp &c = comps[bc];
    int first = c.head_j, second = c.nodes[first].best_t;
    MNode nn; nn.is_root = 1; nn.comp = bc; nn.lrank = c.merges; nn.grank = total_merges;
    nn.gidx = (int)(n + total_merges); nn.best_t = -1; nn.best_d = 0;
    c.merges++; total_merges++;
    nn.pts = c.nodes[first].pts;
    nn.pts.insert(nn.pts.end(), c.nodes[second].pts.begin(), c.nodes[second].pts.end());
    c.nodes[first].is_root = 0; c.nodes[second].is_root = 0;
    g[c.nodes[first].gidx].root = 0; g[c.nodes[second].gidx].root = 0;
    g.push_back(GNode{c.nodes[first].gidx, c.nodes[second].gidx, 1});
    int jn = (int)c.nodes.size();
    for (int t = jn - 1; t >= 0; --t) {
      if (!c.nodes[t].is_root) continue;
      const std::vector<int> &a = nn.pts, &b = c.nodes[t].pts;
      double total = 0.0;
      for (size_t i = 0; i < a.size(); ++i) for (size_t k = 0; k < b.size(); ++k) total += euclid(a[i], b[k]);
      nn.row_t.push_back(t); nn.row_d.push_back(total / (int)(a.size() * b.size()));
    }
    c.nodes.push_back(nn);
    find_best(c, jn);
    for (int j = 0; j < jn; ++j)
      if (c.nodes[j].is_root && (c.nodes[j].best_t == first || c.nodes[j].best_t == second)) find_best(c, j);
    comp_head(c);
  }
  long nn = (long)g.size();
  for (long i = 0; i < nn; ++i) { is_root[i] = g[i].root; merged_a[i] = g[i].a; merged_b[i] = g[i].b; }
  return nn;
}

long orc_model_cluster_ahc(long n, const uint32_t *p1, const uint32_t *p2, double thr,
                           uint32_t *out_idx, int32_t *out_cluster, int *roots)
{
  std::vector<double> x(n), y(n);
  for (long i = 0; i < n; ++i) { x[i] = p1[i]; y[i] = p2[i]; }
  std::vector<int32_t> r(2 * n + 1), a(2 * n + 1), b(2 * n + 1);
  long nn = orc_model_ahc_tree(n, x.data(), y.data(), (long)thr, r.data(), a.data(), b.data());
  return tree_to_clusters(n, nn, r.data(), a.data(), b.data(), out_idx, out_cluster, roots);
}

}  /* extern "C" */
