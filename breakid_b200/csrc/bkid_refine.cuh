// bkid_refine.cuh -- K7 split-read evidence, K8 breakpoint vote, K9 depth/AF, K10 nib 41-mer.
// Included by bkid_core.cu.
//
// The reference answers every per-cluster question with a BAM index query + inflate
// (src/BreakID.cc:436-474).  Here the records are resident in HBM in coordinate order, so a region
// query is two binary searches on (tid,pos) plus a window scan with the index iterator's exact
// overlap rule (htslib hts.c:1776-1777,1963-1965: same tid, pos < end, bam_endpos > max(beg,0)).
#pragma once

// ---- CIGAR algebra (src/CigarRoller.cc:26-136,187-205,316-346; src/Cigar.cc:80-144) -----------------
// A "roller" keeps adjacent equal operations merged; op codes follow the reference enum
// (src/Cigar.h:66-77): 1 match, 3 insert, 4 del, 5 skip, 6 softClip, 7 hardClip, 8 pad.
struct Roller {
  // only what the path needs: first two merged ops, op count, totals
  int nops;
  int op0, op1;
  uint32_t c0, c1;
  int last_op; uint32_t last_cnt;        // last merged op (for end clips)
  int matches, ref_count;
  int begin_clips, end_clips;            // accumulated while scanning
  bool in_begin;
  bool bad;                              // more structure than we summarise and it matters not: regex fails when nops != 2
};

__device__ __forceinline__ void roller_init(Roller &r)
{
  r.nops = 0; r.op0 = r.op1 = 0; r.c0 = r.c1 = 0; r.last_op = 0; r.last_cnt = 0;
  r.matches = 0; r.ref_count = 0; r.begin_clips = 0; r.end_clips = 0; r.in_begin = true; r.bad = false;
}
__device__ __forceinline__ void roller_add(Roller &r, int op, int count)
{
  if ((uint32_t)count == 0) return;                                 // operator+= :33-36
  if (r.nops == 0 || r.last_op != op) {
    if (r.nops == 0) { r.op0 = op; r.c0 = (uint32_t)count; }
    else if (r.nops == 1) { r.op1 = op; r.c1 = (uint32_t)count; }
    r.nops++;
    r.last_op = op; r.last_cnt = (uint32_t)count;
  } else {
    r.last_cnt += (uint32_t)count;
    if (r.nops == 1) r.c0 = r.last_cnt; else if (r.nops == 2) r.c1 = r.last_cnt;
  }
  if (op == 1) r.matches += count;
  if (op == 1 || op == 2 || op == 4 || op == 5) r.ref_count += count;
  bool clip = (op == 6 || op == 7);
  if (clip && r.in_begin) r.begin_clips += count;
  if (!clip) r.in_begin = false;
  if (clip) r.end_clips += count; else r.end_clips = 0;
}
__device__ __forceinline__ void roller_add_char(Roller &r, int ch, int count)   // Add(char,int) :67-117
{
  switch (ch) {
    case 0: case 'M': roller_add(r, 1, count); break;
    case 1: case 'I': roller_add(r, 3, count); break;
    case 2: case 'D': roller_add(r, 4, count); break;
    case 3: case 'N': roller_add(r, 5, count); break;
    case 4: case 'S': roller_add(r, 6, count); break;
    case 5: case 'H': roller_add(r, 7, count); break;
    case 6: case 'P': roller_add(r, 8, count); break;
    case 7: case '=': roller_add(r, 1, count); break;
    case 8: case 'X': roller_add(r, 1, count); break;
    default: break;
  }
}
// all-clip cigars: begin clips count every op, end clips too (both loops run over everything)
__device__ __forceinline__ void roller_set_text(Roller &r, const uint8_t *s, uint32_t len)   // Add(const char*) :120-136
{
  roller_init(r);
  int cnt = 0;
  uint32_t i = 0;
  while (i < len) {
    uint8_t ch = s[i];
    if (ch >= '0' && ch <= '9') {
      // (int) strtol: saturate at LONG_MAX then truncate
      unsigned long long v = 0; bool sat = false;
      while (i < len && s[i] >= '0' && s[i] <= '9') {
        if (!sat) { v = v * 10 + (s[i] - '0'); if (v > 0x7fffffffffffffffull) { sat = true; v = 0x7fffffffffffffffull; } }
        ++i;
      }
      cnt = (int)(long long)v;
    } else { roller_add_char(r, ch, cnt); ++i; }
  }
}
__device__ __forceinline__ void roller_set_bam(Roller &r, const uint32_t *ops, uint32_t n)
{
  roller_init(r);
  for (uint32_t i = 0; i < n; ++i) roller_add_char(r, (int)(ops[i] & 0xF), (int)(ops[i] >> 4));
}
__device__ __forceinline__ char roller_op_char(int op)
{
  switch (op) { case 1: case 2: return 'M'; case 3: return 'I'; case 4: return 'D'; case 5: return 'N'; case 6: return 'S'; case 7: return 'H'; case 8: return 'P'; }
  return '?';
}
// the merged string matches ([0-9]+[MS]){2}  <=>  exactly two merged ops, each M or S
__device__ __forceinline__ bool roller_regex_2ms(const Roller &r)
{
  return r.nops == 2 && (r.op0 == 1 || r.op0 == 6) && (r.op1 == 1 || r.op1 == 6);
}
// full match of ([0-9]+[MS]){2} on raw text
__device__ __forceinline__ bool text_regex_2ms(const uint8_t *s, uint32_t len)
{
  uint32_t i = 0;
  for (int g = 0; g < 2; ++g) {
    uint32_t d = i;
    while (i < len && s[i] >= '0' && s[i] <= '9') ++i;
    if (i == d || i >= len || (s[i] != 'M' && s[i] != 'S')) return false;
    ++i;
  }
  return i == len;
}
__device__ __forceinline__ uint64_t fnv_bytes(uint64_t h, const uint8_t *s, uint32_t len)
{
  for (uint32_t i = 0; i < len; ++i) h = (h ^ s[i]) * 0x100000001b3ULL;
  return h;
}
__device__ __forceinline__ uint64_t fnv_uint(uint64_t h, uint32_t v)     // std::to_string(unsigned)
{
  char buf[10]; int n = 0;
  do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) h = (h ^ (uint8_t)buf[--n]) * 0x100000001b3ULL;
  return h;
}
// hash of the roller's string form (valid for the <= 2-op rollers that pass the regex)
__device__ __forceinline__ uint64_t roller_str_hash(const Roller &r)
{
  uint64_t h = 0xcbf29ce484222325ULL;
  h = fnv_uint(h, r.c0); h = (h ^ (uint8_t)roller_op_char(r.op0)) * 0x100000001b3ULL;
  h = fnv_uint(h, r.c1); h = (h ^ (uint8_t)roller_op_char(r.op1)) * 0x100000001b3ULL;
  return h;
}
// chromosome-name code: t in 0..23 when the string equals chromID2ChrName(t) (src/util_bam.cc:128-142),
// ~0 for "", otherwise FNV hash with the top bit set.  The host applies the same rule to header names.
__device__ __host__ inline uint64_t chr_code(const uint8_t *s, uint32_t len)
{
  if (len == 0) return ~0ull;
  if (len >= 4 && s[0] == 'c' && s[1] == 'h' && s[2] == 'r') {
    if (len == 4 && s[3] == 'X') return 22;
    if (len == 4 && s[3] == 'Y') return 23;
    if (len == 4 && s[3] >= '1' && s[3] <= '9') return (uint64_t)(s[3] - '1');
    if (len == 5 && s[3] >= '1' && s[3] <= '2' && s[4] >= '0' && s[4] <= '9') {
      int v = (s[3] - '0') * 10 + (s[4] - '0');
      if (v >= 10 && v <= 22) return (uint64_t)(v - 1);
    }
  }
  uint64_t h = 0xcbf29ce484222325ULL;
  for (uint32_t i = 0; i < len; ++i) h = (h ^ s[i]) * 0x100000001b3ULL;
  return h | (1ull << 63);
}

struct EvRow {        // == bkid_sarow (include/breakid_b200.h), 96 bytes
  uint64_t pchr, schr, pcig, scig;
  uint64_t name_lo, name_hi;
  uint32_t pstart, sstart, pend, send, pbp, sbp;
  int32_t tid, pos, endpos;      // the SA-tagged record itself (region membership test)
  uint8_t ok;        // complementary cigars (counts as evidence)
  uint8_t fatal;     // the reference would exit(-1) "error cigar" on this record
  uint8_t secondary;
  uint8_t _pad;
};
static_assert(sizeof(EvRow) == 88 || sizeof(EvRow) == 96, "EvRow layout");

// K7a: one thread per SA-tagged record -- src/BreakID.cc:896-1016
__global__ void k7_evidence_rows(const uint32_t *__restrict__ sa_rec, long long n_sa, const uint8_t *__restrict__ cls, const uint16_t *__restrict__ flag, const int32_t *__restrict__ tid,
                                 const int32_t *__restrict__ pos, const int32_t *__restrict__ endpos, const uint16_t *__restrict__ span16, const uint32_t *__restrict__ x_rec, long long n_x,
                                 const uint64_t *__restrict__ x_nh, int *__restrict__ missing, const uint32_t *__restrict__ cig_off, const uint32_t *__restrict__ cig_ops,
                                 const uint32_t *__restrict__ sa_off, const uint8_t *__restrict__ sa_txt, const uint32_t *__restrict__ oc_off,
                                 const uint8_t *__restrict__ oc_txt, int mismatch, EvRow *__restrict__ rows)
{
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_sa) return;
  EvRow R;
  memset(&R, 0, sizeof R);
  uint32_t i = sa_rec[k];
  unsigned fl = flag[i];
  R.tid = tid[i]; R.pos = pos[i]; R.endpos = endpos ? endpos[i] : pos[i] + (int32_t)span16[i];
  {
    long long x = x_slot_of(x_rec, n_x, i);
    if (x < 0) atomicExch(missing, 1);
    else { R.name_lo = x_nh[2 * (size_t)x]; R.name_hi = x_nh[2 * (size_t)x + 1]; }
  }
  const uint8_t *sa = sa_txt + sa_off[k]; uint32_t sal = sa_off[k + 1] - sa_off[k];
  const uint8_t *oc = oc_txt + oc_off[k]; uint32_t ocl = oc_off[k + 1] - oc_off[k];
  // split_string(sa, ",") drops empty fields (src/util_bed.cc:194-222): fields 0,1,3 of the first entry
  uint32_t fs[4], fe[4]; int nf = 0;
  {
    uint32_t p = 0;
    while (p < sal && nf < 4) {
      while (p < sal && sa[p] == ',') ++p;
      if (p >= sal) break;
      uint32_t q = p;
      while (q < sal && sa[q] != ',') ++q;
      fs[nf] = p; fe[nf] = q; ++nf;
      p = q;
    }
  }
  if (sal == 0 || nf < 4 || (fl & F_DUP) || !(fl & F_PAIRED) || (cls[i] & CL_EXCL)) { rows[k] = R; return; }
  Roller sa_c, rec_c, c1;
  roller_set_text(sa_c, sa + fs[3], fe[3] - fs[3]);
  roller_set_bam(rec_c, cig_ops + cig_off[k], cig_off[k + 1] - cig_off[k]);
  if (ocl) roller_set_text(c1, oc, ocl); else c1 = rec_c;
  // is_complementary_cigar (src/CigarRoller.cc:323-346)
  bool ok = roller_regex_2ms(c1) && text_regex_2ms(sa + fs[3], fe[3] - fs[3]);
  if (ok) {
    int c1_m = c1.matches, c2_m = sa_c.matches;
    int c1_s = c1.begin_clips + c1.end_clips, c2_s = sa_c.end_clips + sa_c.begin_clips;
    ok = (c1_m <= c2_s + mismatch && c1_m >= c2_s - mismatch) && (c1_m + c1_s == c2_m + c2_s);
  }
  if (!ok) { rows[k] = R; return; }
  R.ok = 1;
  R.secondary = (fl & F_SECONDARY) ? 1 : 0;
  // stoi(field 1)
  long long sv = 0; bool neg = false;
  {
    uint32_t p = fs[1];
    while (p < fe[1] && (sa[p] == ' ' || (sa[p] >= 9 && sa[p] <= 13))) ++p;
    if (p < fe[1] && (sa[p] == '+' || sa[p] == '-')) { neg = sa[p] == '-'; ++p; }
    while (p < fe[1] && sa[p] >= '0' && sa[p] <= '9') { sv = sv * 10 + (sa[p] - '0'); if (sv > 0x7fffffffll) sv = 0x7fffffffll; ++p; }
    if (neg) sv = -sv;
  }
  uint32_t sa_start = (uint32_t)(int)sv;
  uint32_t sa_end = sa_start + (uint32_t)sa_c.ref_count - 1u;
  uint32_t a_start = (uint32_t)((long long)pos[i] + 1);
  int alen = rec_c.ref_count;
  uint32_t a_end = (uint32_t)((long long)(alen == 0 ? pos[i] : pos[i] + alen - 1) + 1);
  int t = tid[i];
  uint64_t own_chr = (t >= 0 && t < 24) ? (uint64_t)t : ~0ull;
  uint64_t sa_chr = chr_code(sa + fs[0], fe[0] - fs[0]);
  uint32_t own_end = ocl ? (a_start + (uint32_t)c1.ref_count - 1u) : a_end;
  uint64_t own_cig = ocl ? fnv_bytes(0xcbf29ce484222325ULL, oc, ocl) : roller_str_hash(rec_c);
  uint64_t sa_cig = fnv_bytes(0xcbf29ce484222325ULL, sa + fs[3], fe[3] - fs[3]);
  uint32_t own_bp = 0, sa_bp = 0;
  if (c1.begin_clips != 0) own_bp = a_start; else if (c1.end_clips != 0) own_bp = a_end; else R.fatal = 1;
  if (sa_c.begin_clips != 0) sa_bp = sa_start; else if (sa_c.end_clips != 0) sa_bp = sa_end; else R.fatal = 1;
  if (!R.secondary) {
    R.pchr = own_chr; R.pstart = a_start; R.pend = own_end; R.pcig = own_cig; R.pbp = own_bp;
    R.schr = sa_chr; R.sstart = sa_start; R.send = sa_end; R.scig = sa_cig; R.sbp = sa_bp;
  } else {
    R.pchr = sa_chr; R.pstart = sa_start; R.pend = sa_end; R.pcig = sa_cig; R.pbp = sa_bp;
    R.schr = own_chr; R.sstart = a_start; R.send = own_end; R.scig = own_cig; R.sbp = own_bp;
  }
  rows[k] = R;
}

// ---- record window helpers --------------------------------------------------------------------
// first record index with (tid,pos) >= (qt,qp); records are coordinate sorted, tid < 0 sorts last
__device__ __forceinline__ long long rec_lower_bound(const int32_t *__restrict__ tid, const int32_t *__restrict__ pos, long long n, int qt, long long qp)
{
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long m = (lo + hi) >> 1;
    uint32_t tm = (uint32_t)tid[m];
    bool less = tm < (uint32_t)qt || (tm == (uint32_t)qt && (long long)pos[m] < qp);
    if (less) lo = m + 1; else hi = m;
  }
  return lo;
}
__device__ __forceinline__ long long u32_lower_bound(const uint32_t *__restrict__ a, long long n, long long q)
{
  long long lo = 0, hi = n;
  while (lo < hi) { long long m = (lo + hi) >> 1; if ((long long)a[m] < q) lo = m + 1; else hi = m; }
  return lo;
}

struct RegionQ {        // one side of one cluster
  int tid; int beg, end;            // iterator bounds after clamping (beg >= 0); end < beg -> empty
  long long s_lo, s_hi;             // SA-row window: rows with (tid, pos in [beg - maxspan, end))
};

struct RefineView {
  // local record shard (coverage / depth partial counts)
  long long n; const uint8_t *cls; const int32_t *tid, *pos, *endpos; const uint16_t *span16;   // bam_endpos: endpos[i], or pos[i] + span16[i] (narrow column)
  __device__ __forceinline__ int32_t end_of(long long i) const { return endpos ? endpos[i] : pos[i] + (int32_t)span16[i]; }
  // global SA-row table in coordinate order (evidence)
  long long n_sa; const EvRow *rows;
  const uint64_t *name_key; const uint32_t *name_row;   // rows ordered by the low 32 bits of name_lo (evidence pairing)
  int maxspan;
  const uint64_t *canon;            // [nt] chr_code of the header names
  int nt;
  const uint8_t *const *nib; const uint64_t *nib_len;   // per tid, may be null
};

__device__ __forceinline__ long long row_lower_bound(const EvRow *__restrict__ rows, long long n, int qt, long long qp)
{
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long m = (lo + hi) >> 1;
    uint32_t tm = (uint32_t)rows[m].tid;
    bool less = tm < (uint32_t)qt || (tm == (uint32_t)qt && (long long)rows[m].pos < qp);
    if (less) lo = m + 1; else hi = m;
  }
  return lo;
}

__device__ __forceinline__ void make_region(const RefineView &v, int tid, uint32_t start_u, uint32_t end_u, RegionQ &q)
{
  q.tid = tid;
  int beg = (int)start_u, end = (int)end_u;
  if (beg < 0) beg = 0;                                       // hts.c:1776
  q.beg = beg; q.end = end;
  if (end < beg || tid < 0) { q.end = beg - 1; q.s_lo = q.s_hi = 0; return; }
  q.s_lo = row_lower_bound(v.rows, v.n_sa, tid, (long long)beg - v.maxspan);
  q.s_hi = row_lower_bound(v.rows, v.n_sa, tid, (long long)end);
}

struct ClusterWork {
  RegionQ q1, q2;
  uint32_t n_ev1, n_ev2;     // evidence rows that survive the gate (0 when the side is empty)
  uint32_t n_entries;
  uint32_t fatal;
};

__global__ void k7_regions(RefineView v, const bkid_cluster_rec *__restrict__ cl, uint32_t ncl, int w, ClusterWork *__restrict__ work, uint32_t *__restrict__ ev_cap)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  const bkid_cluster_rec &R = cl[c];
  ClusterWork W;
  memset(&W, 0, sizeof W);
  // src/BreakID.cc:430-433: uint64 mean -/+ int w, truncated to uint32, then passed as int
  make_region(v, R.p1_tid, (uint32_t)(R.p1_mean_pos - (unsigned long long)(long long)w), (uint32_t)(R.p1_mean_pos + (unsigned long long)(long long)w), W.q1);
  make_region(v, R.p2_tid, (uint32_t)(R.p2_mean_pos - (unsigned long long)(long long)w), (uint32_t)(R.p2_mean_pos + (unsigned long long)(long long)w), W.q2);
  work[c] = W;
  ev_cap[c] = (uint32_t)((W.q1.s_hi - W.q1.s_lo) + (W.q2.s_hi - W.q2.s_lo));
}

constexpr int RF_THREADS = 128;

// number of LOCAL records the index iterator would return for [beg,end) on tid (every record counts,
// src/BreakID.cc:894); warp-cooperative, result valid in lane 0 after the reduction
__device__ unsigned count_overlaps(const RefineView &v, int tid, int beg, int end, bool depth_only)
{
  if (end < beg || tid < 0 || v.n == 0) return 0;
  long long lo = rec_lower_bound(v.tid, v.pos, v.n, tid, (long long)beg - v.maxspan);
  long long hi = rec_lower_bound(v.tid, v.pos, v.n, tid, (long long)end);
  unsigned c = 0;
  for (long long i = lo + (threadIdx.x & 31); i < hi; i += 32)
    if (v.end_of(i) > beg && (depth_only ? (v.cls[i] & CL_DEPTH) : !(v.cls[i] & CL_EXCL))) ++c;
  return bk::warp_sum(c);
}

// partial coverage of both regions of every cluster on the local record shard: one warp per cluster
__global__ void __launch_bounds__(128) k7_coverage(RefineView v, const ClusterWork *__restrict__ work, uint32_t ncl, uint32_t *__restrict__ cov /* [2*ncl] */)
{
  uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= ncl) return;
  unsigned a = count_overlaps(v, work[c].q1.tid, work[c].q1.beg, work[c].q1.end, false);
  unsigned b = count_overlaps(v, work[c].q2.tid, work[c].q2.beg, work[c].q2.end, false);
  if ((threadIdx.x & 31) == 0) { cov[2 * c] = a; cov[2 * c + 1] = b; }
}

// evidence rows of one region in record order (CTA-cooperative ordered compaction)
__device__ void region_collect(const RefineView &v, const RegionQ &q, uint32_t *__restrict__ list, unsigned *sh_ev, unsigned *sh_fatal)
{
  __shared__ unsigned sh32[33];
  unsigned base = 0;
  for (long long s0 = q.s_lo; s0 < q.s_hi; s0 += blockDim.x) {
    long long s = s0 + threadIdx.x;
    unsigned is = 0;
    if (s < q.s_hi) {
      const EvRow &r = v.rows[s];
      if (r.endpos > q.beg && r.ok) { is = 1; if (r.fatal) atomicExch(sh_fatal, 1u); }
    }
    unsigned tot;
    unsigned rr = bk::block_excl_scan<unsigned>(is, sh32, tot);
    if (is) list[base + rr] = (uint32_t)s;
    base += tot;
  }
  if (threadIdx.x == 0) *sh_ev = base;
  __syncthreads();
}

__device__ __forceinline__ bool ev_match(const RefineView &v, uint32_t sa, uint32_t sb)
{
  const EvRow &a = v.rows[sa], &b = v.rows[sb];
  return a.name_lo == b.name_lo && a.name_hi == b.name_hi &&
         a.secondary != b.secondary && a.pchr == b.pchr && a.schr == b.schr && a.pstart == b.pstart && a.sstart == b.sstart &&
         a.pend == b.pend && a.send == b.send && a.pcig == b.pcig && a.scig == b.scig && a.pbp == b.pbp && a.sbp == b.sbp;   // new_condition :627-637
}

// pass 1 (count) / pass 2 (write) of find_sa_reads x2 + the pairing half of find_bp_pair; cov = TOTAL coverage
template <bool WRITE>
__global__ void __launch_bounds__(RF_THREADS)
k7_collect(RefineView v, const bkid_cluster_rec *__restrict__ cl, uint32_t ncl, ClusterWork *__restrict__ work, const uint32_t *__restrict__ cov,
           const uint32_t *__restrict__ ev_off, uint32_t *__restrict__ ev_list, const uint32_t *__restrict__ ent_off, int2 *__restrict__ entries)
{
  uint32_t c = blockIdx.x;
  if (c >= ncl) return;
  __shared__ unsigned sh_ev, sh_fatal, sh_cnt;
  ClusterWork &W = work[c];
  uint32_t *l1 = ev_list + ev_off[c];
  uint32_t *l2 = l1 + (uint32_t)(W.q1.s_hi - W.q1.s_lo);
  uint32_t n1, n2;
  if (!WRITE) {
    if (threadIdx.x == 0) sh_fatal = 0;
    __syncthreads();
    region_collect(v, W.q1, l1, &sh_ev, &sh_fatal);
    n1 = (cov[2 * c] < 5 || sh_ev < 2) ? 0 : sh_ev;          // gate :1032-1035
    __syncthreads();
    n2 = 0;
    if (n1 > 0) {                                            // side 2 only if side 1 has reads (:438-439)
      region_collect(v, W.q2, l2, &sh_ev, &sh_fatal);
      n2 = (cov[2 * c + 1] < 5 || sh_ev < 2) ? 0 : sh_ev;
    }
    __syncthreads();
    if (threadIdx.x == 0) { W.n_ev1 = n1; W.n_ev2 = n2; W.fatal = sh_fatal; }
  } else { n1 = W.n_ev1; n2 = W.n_ev2; }
  if (threadIdx.x == 0) sh_cnt = 0;
  __syncthreads();
  uint64_t p1code = (cl[c].p1_tid >= 0 && cl[c].p1_tid < v.nt) ? v.canon[cl[c].p1_tid] : chr_code((const uint8_t *)"*", 1);
  if (n1 > 0 && n2 > 0) {
    // the reference walks names in map order, then i over side 1, j over side 2 (src/BreakID.cc:603-760); the vote
    // only needs the multiset of matching (a in side 1, b in side 2) pairs.  A match needs equal read names, so
    // b is looked up through the name-ordered row index instead of scanning all of side 2: side-2 membership is
    // the region predicate itself (row window, iterator overlap rule, complementary cigars).
    const RegionQ &Q2 = W.q2;
    for (uint32_t a = threadIdx.x; a < n1; a += blockDim.x) {
      uint32_t ra = l1[a];
      const EvRow &A = v.rows[ra];
      uint32_t key = (uint32_t)A.name_lo;
      long long lo = 0, hi = v.n_sa;
      while (lo < hi) { long long mid = (lo + hi) >> 1; if ((uint32_t)v.name_key[mid] < key) lo = mid + 1; else hi = mid; }
      for (long long j = lo; j < v.n_sa && (uint32_t)v.name_key[j] == key; ++j) {
        uint32_t rb = v.name_row[j];
        if ((long long)rb < Q2.s_lo || (long long)rb >= Q2.s_hi) continue;
        const EvRow &B = v.rows[rb];
        if (!(B.endpos > Q2.beg && B.ok)) continue;
        if (ev_match(v, ra, rb)) {
          unsigned o = atomicAdd(&sh_cnt, 1u);
          if (WRITE) {
            int2 e;
            if (A.pchr == p1code) { e.x = (int)A.pbp; e.y = (int)A.sbp; }       // :647,671-672
            else { e.x = (int)A.sbp; e.y = (int)A.pbp; }                         // :717-718
            entries[ent_off[c] + o] = e;
          }
        }
      }
    }
  }
  __syncthreads();
  if (!WRITE && threadIdx.x == 0) W.n_entries = sh_cnt;
}

__global__ void k7_entry_counts(const ClusterWork *__restrict__ work, uint32_t ncl, uint32_t *__restrict__ cnt, int *__restrict__ fatal)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl) return;
  cnt[c] = work[c].n_entries;
  if (work[c].fatal) atomicExch(fatal, 1);
}

// lexicographic order of to_string(x)+","+to_string(y) (std::map<string,int>, src/BreakID.cc:599,806)
__device__ __forceinline__ int key_str(int x, int y, char *buf)
{
  int n = 0;
  auto put = [&](int v) {
    char t[12]; int m = 0;
    long long a = v; bool neg = a < 0; if (neg) a = -a;
    do { t[m++] = (char)('0' + a % 10); a /= 10; } while (a);
    if (neg) buf[n++] = '-';
    while (m) buf[n++] = t[--m];
  };
  put(x); buf[n++] = ','; put(y);
  return n;
}
__device__ __forceinline__ bool key_less(int x1, int y1, int x2, int y2)
{
  char a[26], b[26];
  int na = key_str(x1, y1, a), nb = key_str(x2, y2, b);
  int m = na < nb ? na : nb;
  for (int i = 0; i < m; ++i) {
    unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[i];
    if (ca != cb) return ca < cb;
  }
  return na < nb;
}

__device__ __forceinline__ char nib_base(const uint8_t *packed, uint64_t nbases, long long pos, char prev)
{
  if (pos < 0 || (uint64_t)pos >= nbases) return prev;                  // src/nibtools.cc:45-46 leaves the byte untouched
  int b = packed[pos >> 1];
  int x = (pos & 1) ? (b & 0xf) : (b >> 4);                             // high nibble first (:55-61)
  switch (x) { case 0: case 8: return 'T'; case 1: case 9: return 'C'; case 2: case 10: return 'A'; case 3: case 11: return 'G'; default: return 'N'; }
}

constexpr uint32_t K8_SMALL = 256;      // clusters with more evidence entries go through k8_vote_big
constexpr int K8_BIG_THREADS = 256;
constexpr uint32_t K8_BIG_SMEM_KEYS = 8192;

// K8: vote.  One CTA per cluster; writes exact positions / votes and valid[c].
__global__ void __launch_bounds__(RF_THREADS)
k8_vote(bkid_cluster_rec *__restrict__ cl, uint32_t ncl, const ClusterWork *__restrict__ work, const uint32_t *__restrict__ ent_off,
        const int2 *__restrict__ entries, int bp_err, uint32_t *__restrict__ valid)
{
  uint32_t c = blockIdx.x;
  if (c >= ncl) return;
  uint32_t m = work[c].n_entries;
  if (m > K8_SMALL) return;                                               // k8_vote_big
  const int2 *E = entries + ent_off[c];
  // votes: entries within +-bp_err of each key, mixed int32/uint32 compares (:820-821)
  int my_best = 0, mx = -1, my = -1;
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
    uint32_t k1 = (uint32_t)E[i].x, k2 = (uint32_t)E[i].y;
    int cnt = 0;
    for (uint32_t j = 0; j < m; ++j) {
      uint32_t a = (uint32_t)E[j].x, b = (uint32_t)E[j].y;
      if (a <= k1 + (uint32_t)bp_err && a >= k1 - (uint32_t)bp_err && b <= k2 + (uint32_t)bp_err && b >= k2 - (uint32_t)bp_err) ++cnt;
    }
    if (cnt > my_best || (cnt == my_best && cnt > 0 && key_less(E[i].x, E[i].y, mx, my))) { my_best = cnt; mx = E[i].x; my = E[i].y; }
  }
  // block arg-max: highest count, ties -> lexicographically smallest key string (first strict max in map order, :841-855)
  for (int o = 16; o; o >>= 1) {
    int oc = __shfl_xor_sync(0xffffffffu, my_best, o), ox = __shfl_xor_sync(0xffffffffu, mx, o), oy = __shfl_xor_sync(0xffffffffu, my, o);
    if (oc > my_best || (oc == my_best && oc > 0 && key_less(ox, oy, mx, my))) { my_best = oc; mx = ox; my = oy; }
  }
  __shared__ int wb[RF_THREADS / 32], wx[RF_THREADS / 32], wy[RF_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { wb[threadIdx.x >> 5] = my_best; wx[threadIdx.x >> 5] = mx; wy[threadIdx.x >> 5] = my; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0, x = -1, y = -1;
    for (int q = 0; q < RF_THREADS / 32; ++q)
      if (wb[q] > b || (wb[q] == b && b > 0 && key_less(wx[q], wy[q], x, y))) { b = wb[q]; x = wx[q]; y = wy[q]; }
    bool ok = b >= 2;                                                     // :446
    valid[c] = ok ? 1u : 0u;
    if (ok) { cl[c].p1_exact_pos = (uint32_t)x; cl[c].p2_exact_pos = y; cl[c].n_split_read = b; }
  }
}

// K8 for heavy clusters (split-read hotspots): the reference counts identical "x,y" keys in a std::map and
// then sums, for every key, the counts of all keys within +-bp_err (src/BreakID.cc:806-855).  Same structure
// here: CTA-wide bitonic sort of the entries (shared memory up to 8192 entries, in place in global memory
// beyond), unique keys with multiplicities, then the vote over unique keys with an x-window binary search.
__global__ void __launch_bounds__(K8_BIG_THREADS)
k8_vote_big(bkid_cluster_rec *__restrict__ cl, uint32_t ncl, const ClusterWork *__restrict__ work, const uint32_t *__restrict__ ent_off,
            int2 *__restrict__ entries, uint64_t *__restrict__ uniq, uint32_t *__restrict__ ustart, int bp_err, uint32_t *__restrict__ valid)
{
  extern __shared__ uint64_t k8_sh[];
  uint32_t c = blockIdx.x;
  if (c >= ncl) return;
  uint32_t m = work[c].n_entries;
  if (m <= K8_SMALL) return;
  uint64_t *G = reinterpret_cast<uint64_t *>(entries + ent_off[c]);
  uint64_t *U = uniq + ent_off[c];
  uint32_t *S0 = ustart + ent_off[c];
  const bool in_sh = m <= K8_BIG_SMEM_KEYS;
  uint64_t *K = in_sh ? k8_sh : G;
  // key = (uint32 x) << 32 | (uint32 y)
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
    int2 e = reinterpret_cast<const int2 *>(G)[i];
    K[i] = ((uint64_t)(uint32_t)e.x << 32) | (uint32_t)e.y;
  }
  __syncthreads();
  uint32_t P = 1; while (P < m) P <<= 1;
  for (uint32_t k = 2; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      bool flip = (j == (k >> 1));
      for (uint32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        uint32_t i = 2 * t - (t & (j - 1));
        uint32_t p = flip ? (i ^ (2 * j - 1)) : (i + j);
        if (p < m) {                                   // out-of-range slots are +inf: never swapped down
          uint64_t a = K[i], b = K[p];
          if (a > b) { K[i] = b; K[p] = a; }
        }
      }
      __syncthreads();
    }
  }
  // unique keys in order + first index of each run
  __shared__ unsigned sh32[33];
  __shared__ unsigned sh_nu;
  unsigned base = 0;
  for (uint32_t i0 = 0; i0 < m; i0 += blockDim.x) {
    uint32_t i = i0 + threadIdx.x;
    unsigned head = (i < m && (i == 0 || K[i] != K[i - 1])) ? 1u : 0u;
    unsigned tot;
    unsigned r = bk::block_excl_scan<unsigned>(head, sh32, tot);
    if (head) { U[base + r] = K[i]; S0[base + r] = i; }
    base += tot;
  }
  if (threadIdx.x == 0) sh_nu = base;
  __syncthreads();
  uint32_t nu = sh_nu;
  uint32_t e = (uint32_t)bp_err;
  int my_best = 0, mx = -1, my = -1;
  for (uint32_t i = threadIdx.x; i < nu; i += blockDim.x) {
    uint32_t k1 = (uint32_t)(U[i] >> 32), k2 = (uint32_t)U[i];
    uint32_t jlo = 0, jhi = nu;
    if (k1 >= e && k1 <= 0xffffffffu - e) {             // no uint32 wrap: the x condition is a contiguous key range
      uint64_t lo_key = (uint64_t)(k1 - e) << 32, hi_key = ((uint64_t)(k1 + e) << 32) | 0xffffffffull;
      uint32_t lo = 0, hi = nu;
      while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (U[mid] < lo_key) lo = mid + 1; else hi = mid; }
      jlo = lo; hi = nu;
      while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (U[mid] <= hi_key) lo = mid + 1; else hi = mid; }
      jhi = lo;
    }
    int cnt = 0;
    for (uint32_t j = jlo; j < jhi; ++j) {
      uint32_t a = (uint32_t)(U[j] >> 32), b = (uint32_t)U[j];
      if (a <= k1 + e && a >= k1 - e && b <= k2 + e && b >= k2 - e) cnt += (int)(((j + 1 < nu) ? S0[j + 1] : m) - S0[j]);
    }
    if (cnt > my_best || (cnt == my_best && cnt > 0 && key_less((int)k1, (int)k2, mx, my))) { my_best = cnt; mx = (int)k1; my = (int)k2; }
  }
  for (int o = 16; o; o >>= 1) {
    int oc = __shfl_xor_sync(0xffffffffu, my_best, o), ox = __shfl_xor_sync(0xffffffffu, mx, o), oy = __shfl_xor_sync(0xffffffffu, my, o);
    if (oc > my_best || (oc == my_best && oc > 0 && key_less(ox, oy, mx, my))) { my_best = oc; mx = ox; my = oy; }
  }
  __shared__ int wb[K8_BIG_THREADS / 32], wx[K8_BIG_THREADS / 32], wy[K8_BIG_THREADS / 32];
  if ((threadIdx.x & 31) == 0) { wb[threadIdx.x >> 5] = my_best; wx[threadIdx.x >> 5] = mx; wy[threadIdx.x >> 5] = my; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0, x = -1, y = -1;
    for (int q = 0; q < K8_BIG_THREADS / 32; ++q)
      if (wb[q] > b || (wb[q] == b && b > 0 && key_less(wx[q], wy[q], x, y))) { b = wb[q]; x = wx[q]; y = wy[q]; }
    bool ok = b >= 2;                                                     // :446
    valid[c] = ok ? 1u : 0u;
    if (ok) { cl[c].p1_exact_pos = (uint32_t)x; cl[c].p2_exact_pos = y; cl[c].n_split_read = b; }
  }
}

// name index: key = name_lo of every evidence row, value = row number (sorted on the low 32 bits afterwards)
__global__ void k7_name_keys(const EvRow *__restrict__ rows, long long n, uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { key[i] = rows[i].name_lo; val[i] = (uint32_t)i; }
}

// K9 partial: depth at both break points of every valid cluster on the local record shard
// (src/util_bed.cc:154-192: iterator over [pos-1, pos); qual>0 && !DUP && PAIRED = class bit CL_DEPTH)
__global__ void __launch_bounds__(128) k9_depth(RefineView v, const bkid_cluster_rec *__restrict__ cl, const uint32_t *__restrict__ valid, uint32_t ncl, uint32_t *__restrict__ depth)
{
  uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= ncl) return;
  unsigned a = 0, b = 0;
  if (valid[c]) {
    unsigned long long p1 = (unsigned long long)cl[c].p1_exact_pos, p2 = (unsigned long long)(long long)cl[c].p2_exact_pos;
    int b1 = (int)(p1 - 1), e1 = (int)p1, b2 = (int)(p2 - 1), e2 = (int)p2;
    if (b1 < 0) b1 = 0;
    if (b2 < 0) b2 = 0;
    a = count_overlaps(v, cl[c].p1_tid, b1, e1, true);
    b = count_overlaps(v, cl[c].p2_tid, b2, e2, true);
  }
  if ((threadIdx.x & 31) == 0) { depth[2 * c] = a; depth[2 * c + 1] = b; }
}

// K9/K10 finish: depth (TOTAL over shards), AF, 41-mers, homopolymer flag; one thread per cluster
__global__ void k10_finish(RefineView v, bkid_cluster_rec *__restrict__ cl, const uint32_t *__restrict__ valid, const uint32_t *__restrict__ depth, uint32_t ncl)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncl || !valid[c]) return;
  bkid_cluster_rec &R = cl[c];
  R.p1_bp_depth = (double)depth[2 * c]; R.p2_bp_depth = (double)depth[2 * c + 1];
  R.p1_alle_freq = __fdiv_rn((float)(long long)R.n_split_read, (float)R.p1_bp_depth);   // :475-478
  R.p2_alle_freq = __fdiv_rn((float)(long long)R.n_split_read, (float)R.p2_bp_depth);
  int rpt = 0;
  for (int side = 0; side < 2; ++side) {
    // 41-mer = 1-based [bp-20, bp+20] (src/BreakID.cc:554-557, src/util_bam.cc:78-122); sequential because an
    // out-of-range base repeats the previous one
    int t = side ? R.p2_tid : R.p1_tid;
    long long bp = side ? (long long)R.p2_exact_pos : (long long)(int32_t)R.p1_exact_pos;
    char *out = side ? R.p2_rpt : R.p1_rpt;
    for (int k = 0; k < 44; ++k) out[k] = 0;
    if (v.nib && t >= 0 && t < v.nt && v.nib[t]) {
      char prev = 'N';
      for (int k = 0; k < 41; ++k) { prev = nib_base(v.nib[t], v.nib_len[t], bp - 21 + k, prev); out[k] = prev; }
    }
    int best = 0;
    for (int i = 0; out[i];) { int j = i; while (out[j] == out[i]) ++j; if (j - i > best) best = j - i; i = j; }
    if (best > 10) rpt = 1;                                                               // src/BreakID.cc:560-561
  }
  R.is_rpt = rpt;
}

__global__ void max_span_kernel(const int32_t *__restrict__ pos, const int32_t *__restrict__ endpos, long long n, int *__restrict__ out)
{
  int m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = max(m, endpos[i] - pos[i]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

__global__ void compact_clusters(const bkid_cluster_rec *__restrict__ src, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ off, uint32_t n,
                                 bkid_cluster_rec *__restrict__ dst)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n && keep[c]) dst[off[c]] = src[c];
}
