/* oracle/ref_shim.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * C-linkage wrappers around the REAL reference functions so that pytest can call them through
 * ctypes and diff them against oracle/oracle.cc and against the CUDA path.  The reference's
 * BreakID.cc is included textually (with main renamed) because src/BreakID.h:170-188 defines
 * non-inline functions in the header; the other reference .cc files are linked as objects.
 * Nothing from the reference is copied into this repository: the include path points at
 * /root/reference/src (see oracle/Makefile).  Built only when /root/reference exists.
 */
#define main breakid_reference_main
#include "BreakID.cc"
#undef main
#include "oracle.h"

static int tid_of(bam_header_t *h, const std::string &name)
{
  for (int i = 0; i < h->n_targets; ++i)
    if (name == h->target_name[i]) return i;
  return -1;
}

static int chrname_id(const std::string &s, std::vector<std::string> &others)
{
  if (s.empty()) return -1;
  for (int t = 0; t < 24; ++t) if (chromID2ChrName(t) == s) return t;
  for (size_t k = 0; k < others.size(); ++k) if (others[k] == s) return -2 - (int)k;
  others.push_back(s);
  return -2 - (int)(others.size() - 1);
}

extern "C" {

int ref_main(int argc, char **argv) { return breakid_reference_main(argc, argv); }

void ref_free(void *p) { free(p); }

/* src/BreakID.cc:1909-1954 */
void ref_insert_stats(const char *bam, double *mean, double *sd)
{
  std::vector<double> v;
  get_mean_insert_size(bam, v);
  *mean = v[0]; *sd = v[1];
}

/* src/BreakID.cc:1362-1515; pairs are returned bucket by bucket in std::map order, inside a
 * bucket in emission order -- exactly the order main() consumes them (src/BreakID.cc:119). */
long ref_scan(const char *bam, int qual, double w, const char *nib_dir, orc_pair **out)
{
  std::map<std::string, std::vector<discordant_pair>> m;
  scan_discordant_pairs(bam, "hg19", qual, w, m, nib_dir);
  samfile_t *fp = samopen(bam, "rb", 0);
  long n = 0;
  for (auto &kv : m) n += (long)kv.second.size();
  orc_pair *o = (orc_pair *)calloc(n ? n : 1, sizeof(orc_pair));
  long k = 0; int b = 0;
  for (auto &kv : m) {
    for (auto &d : kv.second) {
      orc_pair &p = o[k];
      orc_name_hash(d.qname.c_str(), &p.name_lo, &p.name_hi);
      p.p1_tid = tid_of(fp->header, d.p1_chr); p.p2_tid = tid_of(fp->header, d.p2_chr);
      p.p1_pos = d.p1_pos; p.p2_pos = d.p2_pos;
      p.p1_chr_pos = d.p1_chr_pos; p.p2_chr_pos = d.p2_chr_pos;
      p.p1_flag = (uint16_t)d.p1_flag; p.p2_flag = (uint16_t)d.p2_flag;
      p.p1_mapq = (uint8_t)d.p1_mapq; p.p2_mapq = (uint8_t)d.p2_mapq;
      p.p1_strand = d.p1_strand; p.p2_strand = d.p2_strand;
      p.bucket = b; p.cluster = -1; p.orig = (uint32_t)k;
      ++k;
    }
    ++b;
  }
  samclose(fp);
  *out = o;
  return n;
}

static void to_vec(long n, const uint32_t *p1, const uint32_t *p2, std::vector<discordant_pair> &v)
{
  v.resize(n);
  for (long i = 0; i < n; ++i) {
    v[i].p1_chr_pos = p1[i]; v[i].p2_chr_pos = p2[i];
    v[i].p1_flag = i;            /* carries the original index through the reference code */
    v[i].is_isolated = 0; v[i].cluster = -1;
  }
}

/* src/BreakID.cc:1271-1285 (+ mask_pairs_chr_pos :1813-1877) */
long ref_remove_isolated(long n, const uint32_t *p1, const uint32_t *p2, double w, uint32_t *out_idx)
{
  std::vector<discordant_pair> v; to_vec(n, p1, p2, v);
  remove_isolated_pairs(v, w);
  for (size_t i = 0; i < v.size(); ++i) out_idx[i] = (uint32_t)v[i].p1_flag;
  return (long)v.size();
}

/* src/BreakID.cc:1304-1352 + src/util_cluster.cc, followed by the sort of src/BreakID.cc:144 */
long ref_cluster_ahc(long n, const uint32_t *p1, const uint32_t *p2, double thr,
                     uint32_t *out_idx, int32_t *out_cluster, int *roots, int sort_by_id)
{
  std::vector<discordant_pair> v; to_vec(n, p1, p2, v);
  *roots = find_cluster_pairs_enspan_ahc(v, thr, 1, 2);
  if (sort_by_id) sort(v.begin(), v.end(), cmp_enspan_id);
  for (size_t i = 0; i < v.size(); ++i) { out_idx[i] = (uint32_t)v[i].p1_flag; out_cluster[i] = v[i].cluster; }
  return (long)v.size();
}

/* src/BreakID.cc:1046-1160 */
long ref_cluster_fast(long n, const uint32_t *p1, const uint32_t *p2, double w,
                      uint32_t *out_idx, int32_t *out_cluster, int *roots, int sort_by_id)
{
  std::vector<discordant_pair> v; to_vec(n, p1, p2, v);
  *roots = find_cluster_pairs_enspan_fast(v, w, 2);
  if (sort_by_id) sort(v.begin(), v.end(), cmp_enspan_id);
  for (size_t i = 0; i < v.size(); ++i) { out_idx[i] = (uint32_t)v[i].p1_flag; out_cluster[i] = v[i].cluster; }
  return (long)v.size();
}

/* the three std::sort call shapes the pipeline depends on (src/BreakID.cc:1274,1278,144):
 * which: 0 = cmp_p1, 1 = cmp_p2, 2 = cmp_enspan_id (key in p1). perm[i] = original index. */
void ref_std_sort_perm(long n, const uint32_t *key, int which, uint32_t *perm)
{
  std::vector<discordant_pair> v(n);
  for (long i = 0; i < n; ++i) { v[i].p1_chr_pos = key[i]; v[i].p2_chr_pos = key[i]; v[i].cluster = (int)key[i]; v[i].p1_flag = i; }
  if (which == 0) sort(v.begin(), v.end(), cmp_p1_enspan_pairs);
  else if (which == 1) sort(v.begin(), v.end(), cmp_p2_enspan_pairs);
  else sort(v.begin(), v.end(), cmp_enspan_id);
  for (long i = 0; i < n; ++i) perm[i] = (uint32_t)v[i].p1_flag;
}

/* src/BreakID.cc:868-1037; rows in std::map<qname> order, inside a name in push order.
 * Returns -1-k when the region gate (coverage<5 or evidence<2) cleared the map (k rows seen is lost). */
long ref_find_sa_reads(const char *bam, const char *chr, uint32_t start, uint32_t end, orc_evidence **out)
{
  samfile_t *fp = samopen(bam, "rb", 0);
  bam_index_t *idx = bam_index_load(bam);
  std::map<std::string, std::vector<split_align_pair>> m;
  find_sa_reads(fp, chr, start, end, m, idx);
  long n = 0;
  for (auto &kv : m) n += (long)kv.second.size();
  orc_evidence *o = (orc_evidence *)calloc(n ? n : 1, sizeof(orc_evidence));
  std::vector<std::string> others;
  long k = 0;
  for (auto &kv : m)
    for (auto &s : kv.second) {
      orc_evidence &e = o[k++];
      orc_name_hash(s.read_name.c_str(), &e.name_lo, &e.name_hi);
      e.primary_chr = chrname_id(s.primary_chr, others);
      e.secondary_chr = chrname_id(s.secondary_chr, others);
      e.primary_start = s.primary_start; e.secondary_start = s.secondary_start;
      e.primary_end = s.primary_end; e.secondary_end = s.secondary_end;
      e.primary_bp = s.primary_bp; e.secondary_bp = s.secondary_bp;
      e.primary_cigar_h = orc_str_hash(s.primary_cigar_str.c_str());
      e.secondary_cigar_h = orc_str_hash(s.secondary_cigar_str.c_str());
      e.flag = (uint16_t)s.flag; e.secondary = s.secondary;
    }
  samclose(fp);
  *out = o;
  return n;
}

/* src/BreakID.cc:436-446 (two region pulls + vote).  Returns encompass_num. */
int ref_find_bp(const char *bam, const char *chr1, uint32_t s1, uint32_t e1,
                const char *chr2, uint32_t s2, uint32_t e2, int32_t *p1_bp, int32_t *p2_bp)
{
  samfile_t *fp = samopen(bam, "rb", 0);
  bam_index_t *idx = bam_index_load(bam);
  std::map<std::string, std::vector<split_align_pair>> m1, m2;
  breakpoint_pair bp; bp.encompass_num = 0; bp.p1_bp = -1; bp.p2_bp = -1;
  std::vector<bam1_t *> sr;
  find_sa_reads(fp, chr1, s1, e1, m1, idx);
  if (m1.size() > 0) find_sa_reads(fp, chr2, s2, e2, m2, idx);
  if (m1.size() > 0 && m2.size() > 0) find_bp_pair(m1, m2, bp, chr1, chr2, sr, 2);
  samclose(fp);
  *p1_bp = bp.p1_bp; *p2_bp = bp.p2_bp;
  return bp.encompass_num;
}

/* src/util_bed.cc:154-192 */
double ref_single_base_depth(const char *bam, const char *chr, uint64_t pos)
{
  samfile_t *fp = samopen(bam, "rb", 0);
  bam_index_t *idx = bam_index_load(bam);
  double d = cal_single_base_depth(chr, pos, fp, idx);
  samclose(fp);
  return d;
}

/* src/CigarRoller.cc:323-346 with c1 set from a string (as at src/BreakID.cc:906-915) */
int ref_is_complementary(const char *c1, const char *c2, int err)
{
  CigarRoller r; r.Set(c1);
  return r.is_complementary_cigar(c2, err) ? 1 : 0;
}

/* src/BreakID.cc:554-561 */
void ref_neighbor_41(const char *nib_dir, const char *chr, int32_t bp, char *out42)
{
  std::string l = get_left_neighbor_sequence_nib(chr, bp, 20, nib_dir);
  std::string r = get_right_neighbor_sequence_nib(chr, bp - 1, 21, nib_dir);
  std::string s = l + r;
  memset(out42, 0, 42);
  memcpy(out42, s.data(), s.size() < 41 ? s.size() : 41);
}
int ref_longest_repeat(const char *s) { return find_longest_repeat_substring(s); }

/* raw util_cluster: returns number of nodes; node_pts/node_off give each node's point list,
 * is_root flags; lets tests diff the merge tree itself (src/util_cluster.cc:7-396). */
long ref_ahc_tree(long n, const double *x, const double *y, long thr,
                  int32_t *is_root, int32_t *merged_a, int32_t *merged_b)
{
  std::vector<point> pts(n);
  for (long i = 0; i < n; ++i) { pts[i].pos.x = x[i]; pts[i].pos.y = y[i]; }
  cluster_struct c;
  init_cluster(c, thr, pts, 1);
  for (int i = 0; i < c.num_nodes; ++i) {
    is_root[i] = c.nodes[i].is_root;
    merged_a[i] = c.nodes[i].merged.size() == 2 ? c.nodes[i].merged[0] : -1;
    merged_b[i] = c.nodes[i].merged.size() == 2 ? c.nodes[i].merged[1] : -1;
  }
  return c.num_nodes;
}

} /* extern "C" */
