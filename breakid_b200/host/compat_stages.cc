// compat_stages.cc -- the reference's stage functions (include/compat/BreakID_stages.h, util_cluster.h, util_bam.h) over the
// B200 C ABI.  One session per input BAM keeps the decoded records resident on the device between the stage calls, the way
// the reference re-opens the same BAM in every stage (src/BreakID.cc:1909, :1363, :390).
#include <sys/stat.h>

#include <algorithm>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>

#include "../../include/breakid_b200.h"
#include "../../include/compat/BreakID_stages.h"
#include "../../include/compat/util_bed.h"
#include "annotate.h"
#include "bam_reader.h"
#include "nibtools.h"

namespace {

[[noreturn]] void die(const std::string &what)
{
  std::cerr << "Error: " << what << std::endl;
  exit(1);
}

struct Session {
  std::string bam;
  bkid_host_bgzf *file = nullptr;
  bkid_ctx *ctx = nullptr;
  bkid_params prm;
  std::vector<std::string> names;
  std::vector<uint32_t> lens;
  std::map<std::string, int> tid_of;
  std::string nib_dir;                 // directory whose nib files are on the device
  std::vector<Transcript> tx;
  RefGeneIndex tx_index;
  bool have_tx = false;
  ~Session()
  {
    if (ctx) bkid_destroy(ctx);
    if (file) bkid_host_bgzf_close(file);
  }
  void check(int rc, const char *what)
  {
    if (rc) die(std::string(what) + ": " + bkid_last_error(ctx));
  }
};
std::unique_ptr<Session> g_session;

// the session of `bam`: opens the file, inflates + decodes it on the device (once)
Session &session(const std::string &bam)
{
  if (g_session && g_session->bam == bam) return *g_session;
  g_session.reset(new Session());
  Session &s = *g_session;
  s.bam = bam;
  char err[512] = {0};
  s.file = bkid_host_bgzf_open(bam.c_str(), err, sizeof err);
  if (!s.file) die("can not open bam-file: " + bam + " (" + err + ")");                  // src/BreakID.cc:1917-1921
  const bkid_header *hdr = bkid_host_bgzf_header(s.file);
  for (int t = 0; t < hdr->n_targets; ++t) {
    s.names.emplace_back(hdr->target_name[t]);
    s.lens.push_back(hdr->target_len[t]);
    s.tid_of[s.names.back()] = t;
  }
  bkid_default_params(&s.prm);
  const char *g = getenv("BREAKID_GPU");
  s.ctx = bkid_create(g ? atoi(g) : 0, hdr, &s.prm);
  if (!s.ctx) die(std::string("cannot create the device context: ") + bkid_last_error(nullptr));
  int64_t n = 0;
  if (bkid_push_bgzf(s.ctx, bkid_host_bgzf_data(s.file), bkid_host_bgzf_size(s.file), bkid_host_bgzf_blocks(s.file), bkid_host_bgzf_n_blocks(s.file),
                     bkid_host_bgzf_first_record(s.file), &n))
    die("can not read bam-file: " + bam + " (" + bkid_last_error(s.ctx) + ")");
  return s;
}

Session &current()
{
  if (!g_session) die("no input BAM has been opened yet (call get_mean_insert_size or scan_discordant_pairs first)");
  return *g_session;
}

void set_params(Session &s, const bkid_params &p)
{
  if (!memcmp(&p, &s.prm, sizeof p)) return;
  s.check(bkid_set_params(s.ctx, &p), "set_params");
  s.prm = p;
}

void load_nibs(Session &s, const std::string &dir)
{
  if (s.nib_dir == dir) return;
  for (size_t t = 0; t < s.names.size(); ++t) {
    nib nb;
    if (nb.open(dir + "/hg19_" + s.names[t] + ".nib") == 0) s.check(bkid_set_nib(s.ctx, (int32_t)t, nb.payload(), nb.size()), "set_nib");
  }
  s.nib_dir = dir;
}

std::string hex128(uint64_t lo, uint64_t hi)
{
  char b[40];
  snprintf(b, sizeof b, "%016llx%016llx", (unsigned long long)hi, (unsigned long long)lo);
  return b;
}

const char *kFusion[] = {"Unknown", "Translocation", "Inversion", "Duplication", "Deletion"};

std::string refgene_path()
{
  if (const char *e = getenv("BREAKID_REFGENE")) return e;
  if (const char *e = getenv("BREAKID_INSTALLDIR")) return std::string(e) + "/ref_files/refGene.txt";
  return "ref_files/refGene.txt";
}

void ensure_refgene(Session &s)
{
  if (s.have_tx) return;
  std::string p = refgene_path();
  if (!load_refgene(p, s.tx)) die("cannot open \t" + p);                                  // src/RefSeqTranscript.cc:205-209
  s.tx_index.build(s.tx);
  s.have_tx = true;
}

// pair coordinates of one bucket as the two device operand columns
void coords(const std::vector<discordant_pair> &v, std::vector<uint32_t> &x, std::vector<uint32_t> &y)
{
  x.resize(v.size()); y.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) { x[i] = v[i].p1_chr_pos; y[i] = v[i].p2_chr_pos; }
}

// device clustering of one point set: members (index, cluster id) of the clusters with >= 2 points, cluster ids in the
// reference's root order, plus the root count
void cluster_points(Session &s, int mode, const std::vector<uint32_t> &x, const std::vector<uint32_t> &y, double thr, std::vector<uint32_t> &idx,
                    std::vector<int32_t> &cl, int32_t &n_roots)
{
  int64_t n = (int64_t)x.size(), n_out = 0;
  idx.assign((size_t)n + 1, 0); cl.assign((size_t)n + 1, 0);
  n_roots = 0;
  s.check(bkid_op_cluster(s.ctx, mode, n, x.data(), y.data(), thr, idx.data(), cl.data(), &n_out, &n_roots), mode ? "find_cluster_pairs_enspan_fast" : "init_cluster");
  idx.resize((size_t)n_out); cl.resize((size_t)n_out);
}

// keep the clusters with at least min_reads members, renumbered densely in their old order (src/BreakID.cc:1328-1352)
void apply_clusters(std::vector<discordant_pair> &enspan, const std::vector<uint32_t> &idx, const std::vector<int32_t> &cl, int min_reads)
{
  std::vector<discordant_pair> out;
  size_t i = 0;
  int k = 0;
  while (i < idx.size()) {
    size_t j = i;
    while (j < idx.size() && cl[j] == cl[i]) ++j;
    if ((int)(j - i) >= min_reads) {
      std::string label = "cluster_No_" + std::to_string(k);
      for (size_t m = i; m < j; ++m) {
        discordant_pair p = enspan[idx[m]];
        p.cluster = k; p.cluster_id = label;
        out.push_back(p);
      }
      ++k;
    }
    i = j;
  }
  enspan.swap(out);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// src/BreakID.h:157-228
// ---------------------------------------------------------------------------------------------------------------
void get_mean_insert_size(std::string input_bam, std::vector<double> &insert)
{
  Session &s = session(input_bam);
  double mean = 0, sd = 0;
  s.check(bkid_insert_stats(s.ctx, &mean, &sd), "get_mean_insert_size");
  insert.clear();
  insert.push_back(mean);
  insert.push_back(sd);
}

void scan_discordant_pairs(const std::string &inp_file, const std::string &build, long qual, double w,
                           std::map<std::string, std::vector<discordant_pair>> &enspan_map, std::string nib_dir)
{
  (void)build;
  Session &s = session(inp_file);
  bkid_params p = s.prm;
  p.qual = (int32_t)qual;
  set_params(s, p);
  if (!nib_dir.empty()) load_nibs(s, nib_dir);
  int64_t np = 0, got = 0;
  s.check(bkid_scan(s.ctx, w, &np), "scan_discordant_pairs");
  std::vector<bkid_pair> pairs((size_t)std::max<int64_t>(np, 1));
  s.check(bkid_fetch_pairs(s.ctx, 0, pairs.data(), (int64_t)pairs.size(), &got), "scan_discordant_pairs");
  pairs.resize((size_t)got);
  // emission order of the reference (the record that completes a pair, in file order) inside every bucket
  std::vector<uint32_t> order(pairs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return pairs[a].orig < pairs[b].orig; });
  auto nm = [&](int t) { return t >= 0 && t < (int)s.names.size() ? s.names[t] : std::string("*"); };
  for (uint32_t i : order) {
    const bkid_pair &q = pairs[i];
    discordant_pair d;
    d.qname = hex128(q.name_lo, q.name_hi);
    d.p1_flag = q.p1_flag; d.p2_flag = q.p2_flag;
    d.p1_chr = nm(q.p1_tid); d.p2_chr = nm(q.p2_tid);
    d.p1_pos = q.p1_pos; d.p2_pos = q.p2_pos;
    d.p1_mapq = q.p1_mapq; d.p2_mapq = q.p2_mapq;
    d.p1_strand = (char)q.p1_strand; d.p2_strand = (char)q.p2_strand;
    d.p1_chr_pos = q.p1_chr_pos; d.p2_chr_pos = q.p2_chr_pos;
    enspan_map[d.p1_chr + "_" + d.p2_chr].push_back(d);                                 // src/BreakID.cc:1500-1512
  }
}

void add_enspan_point_id(std::vector<discordant_pair> &enspan_vec)
{
  for (size_t i = 0; i < enspan_vec.size(); ++i) enspan_vec[i].id = "pair_No_" + std::to_string(i);
}

void remove_isolated_pairs(std::vector<discordant_pair> &enspans, double w)
{
  Session &s = current();
  std::vector<uint32_t> x, y, keep(enspans.size() + 1);
  coords(enspans, x, y);
  int64_t n_out = 0;
  s.check(bkid_op_remove_isolated(s.ctx, (int64_t)enspans.size(), x.data(), y.data(), w, keep.data(), &n_out), "remove_isolated_pairs");
  std::vector<discordant_pair> out;
  out.reserve((size_t)n_out);
  for (int64_t i = 0; i < n_out; ++i) out.push_back(enspans[keep[(size_t)i]]);
  enspans.swap(out);
}

void build_pair_array(std::vector<discordant_pair> &enspan, std::vector<point> &points)
{
  for (const discordant_pair &d : enspan) {
    point p;
    p.pos.x = d.p1_chr_pos; p.pos.y = d.p2_chr_pos;
    p.label = "x=" + std::to_string(d.p1_chr_pos) + ",y=" + std::to_string(d.p2_chr_pos);
    points.push_back(p);
  }
}

void add_cluster_id_for_enspan_vec(cluster_struct &main_cluster, std::vector<discordant_pair> &enspan, int min_reads_per_cluster)
{
  std::vector<discordant_pair> out;
  int k = 0;
  for (int i = 0; i < main_cluster.num_nodes; ++i) {
    const node &nd = main_cluster.nodes[(size_t)i];
    if (!nd.is_root || nd.num_points < min_reads_per_cluster) continue;
    std::string label = "cluster_No_" + std::to_string(k);
    for (int j : nd.points) {
      discordant_pair p = enspan[(size_t)j];
      p.cluster = k; p.cluster_id = label;
      out.push_back(p);
    }
    ++k;
  }
  enspan.swap(out);
}

int find_cluster_pairs_enspan_ahc(std::vector<discordant_pair> &enspan, double distance_threshold, int distance_type, int min_reads_per_cluster)
{
  if (distance_type != 1) die("find_cluster_pairs_enspan_ahc: only distance_type 1 is implemented (the one src/BreakID.cc:32 uses)");
  Session &s = current();
  std::vector<uint32_t> x, y, idx;
  std::vector<int32_t> cl;
  int32_t roots = 0;
  coords(enspan, x, y);
  cluster_points(s, 0, x, y, distance_threshold, idx, cl, roots);
  apply_clusters(enspan, idx, cl, std::max(min_reads_per_cluster, 2));
  return roots;
}

int find_cluster_pairs_enspan_fast(std::vector<discordant_pair> &enspan, double w, int min_reads)
{
  Session &s = current();
  bkid_params p = s.prm;
  p.min_reads = min_reads;
  set_params(s, p);
  std::vector<uint32_t> x, y, idx;
  std::vector<int32_t> cl;
  int32_t roots = 0;
  coords(enspan, x, y);
  cluster_points(s, 1, x, y, w, idx, cl, roots);
  apply_clusters(enspan, idx, cl, 1);
  return roots;
}

std::string determine_fusion_type_from_drp(cluster_info &cluster)
{
  // later classes overwrite earlier ones (src/BreakID.cc:1888-1907)
  const char *order[4][2] = {{"diff_chr", "Translocation"}, {"same_chr_with_same_orientation", "Inversion"}, {"same_chr_with_absolute_reverse", "Duplication"},
                             {"same_chr_with_default_orientation", "Deletion"}};
  std::string t = "Unknown";
  for (auto &o : order) if (cluster.drp_type_set.count(o[0])) t = o[1];
  return t;
}

void annotate_cluster_for_sa_tag(std::vector<cluster_info> &clusters, std::string nib_dir)
{
  (void)nib_dir;
  Session &s = current();
  ensure_refgene(s);
  for (cluster_info &c : clusters) {
    long p1 = c.p1_exact_pos == (uint32_t)-1 ? (long)c.p1_mean_pos : (long)c.p1_exact_pos;            // src/BreakID.cc:518-534
    long p2 = c.p2_exact_pos == -1 ? (long)c.p2_mean_pos : (long)c.p2_exact_pos;
    SideAnnotation a1 = annotate_side(s.tx, s.tx_index, c.p1_chr, p1), a2 = annotate_side(s.tx, s.tx_index, c.p2_chr, p2);
    c.p1_behalf_gene = a1.gene; c.p1_exon_info = a1.exon_info; c.p1_strand = a1.strand;
    c.p2_behalf_gene = a2.gene; c.p2_exon_info = a2.exon_info; c.p2_strand = a2.strand;
  }
}

void findClusterBreakPointInfoSaTag(std::string bam_file, std::vector<discordant_pair> &enspan, double w, std::vector<cluster_info> &cluster_vec,
                                    std::vector<bam1_t *> &split_reads, std::string nib_dir)
{
  (void)split_reads;
  if (enspan.empty()) return;                                                            // src/BreakID.cc:222 (cluster_vec is left alone)
  Session &s = session(bam_file);
  if (!nib_dir.empty()) load_nibs(s, nib_dir);
  {
    // the reference opens the index as soon as one bucket reaches this stage (src/BreakID.cc:412-416)
    struct stat st;
    std::string a = bam_file + ".bai", b = bam_file.substr(0, bam_file.size() > 4 ? bam_file.size() - 4 : 0) + ".bai";
    if (stat(a.c_str(), &st) != 0 && stat(b.c_str(), &st) != 0) die("please index bam-file first:\t" + bam_file);
  }
  std::vector<bkid_pair> pairs(enspan.size());
  for (size_t i = 0; i < enspan.size(); ++i) {
    const discordant_pair &d = enspan[i];
    bkid_pair &q = pairs[i];
    memset(&q, 0, sizeof q);
    auto t1 = s.tid_of.find(d.p1_chr), t2 = s.tid_of.find(d.p2_chr);
    q.p1_tid = t1 == s.tid_of.end() ? -1 : t1->second;
    q.p2_tid = t2 == s.tid_of.end() ? -1 : t2->second;
    q.p1_pos = d.p1_pos; q.p2_pos = d.p2_pos;
    q.p1_chr_pos = d.p1_chr_pos; q.p2_chr_pos = d.p2_chr_pos;
    q.p1_flag = (uint16_t)d.p1_flag; q.p2_flag = (uint16_t)d.p2_flag;
    q.p1_mapq = (uint8_t)d.p1_mapq; q.p2_mapq = (uint8_t)d.p2_mapq;
    q.p1_strand = (uint8_t)d.p1_strand; q.p2_strand = (uint8_t)d.p2_strand;
    q.cluster = d.cluster;
    q.orig = (uint32_t)i;
  }
  // std::map<long, ...> keyed by cluster id (src/BreakID.cc:213-298): members grouped by ascending id, whatever order they came in
  std::stable_sort(pairs.begin(), pairs.end(), [](const bkid_pair &a, const bkid_pair &b) { return a.cluster < b.cluster; });
  int64_t ncl = 0, n_called = 0, got = 0;
  s.check(bkid_op_summarize(s.ctx, (int64_t)pairs.size(), pairs.data(), w, &ncl), "findClusterBreakPointInfoSaTag");
  std::cout << "there is " << ncl << " cluster after produce cluster data\n";
  s.check(bkid_refine(s.ctx, w, &n_called), "findClusterBreakPointInfoSaTag");
  std::vector<bkid_cluster_rec> rec((size_t)std::max<int64_t>(n_called, 1));
  s.check(bkid_fetch_clusters(s.ctx, rec.data(), (int64_t)rec.size(), &got), "findClusterBreakPointInfoSaTag");
  rec.resize((size_t)got);
  cluster_vec.clear();
  auto nm = [&](int t) { return t >= 0 && t < (int)s.names.size() ? s.names[t] : std::string("*"); };
  for (const bkid_cluster_rec &r : rec) {
    cluster_info c;
    c.id = r.id;
    c.p1_chr = nm(r.p1_tid); c.p2_chr = nm(r.p2_tid);
    c.p1_mean_pos = r.p1_mean_pos; c.p2_mean_pos = r.p2_mean_pos;
    c.p1_min_pos = r.p1_min_pos; c.p1_max_pos = r.p1_max_pos; c.p2_min_pos = r.p2_min_pos; c.p2_max_pos = r.p2_max_pos;
    c.p1_exact_pos = r.p1_exact_pos; c.p2_exact_pos = r.p2_exact_pos;
    c.n_split_read = (long)r.n_split_read; c.n_discordant_pair = (long)r.n_discordant_pair;
    c.fusion_type = kFusion[r.fusion_type >= 0 && r.fusion_type <= 4 ? r.fusion_type : 0];
    c.p1_rpt = r.p1_rpt; c.p2_rpt = r.p2_rpt;
    c.is_rpt = r.is_rpt != 0;
    c.p1_bp_depth = r.p1_bp_depth; c.p2_bp_depth = r.p2_bp_depth;
    c.p1_alle_freq = r.p1_alle_freq; c.p2_alle_freq = r.p2_alle_freq;
    cluster_vec.push_back(c);
  }
  annotate_cluster_for_sa_tag(cluster_vec, nib_dir);
}

void write_enspan_params(std::string inp_file, std::string out_file, std::string build, double w, long qual)
{
  std::ofstream p((out_file + "_params.txt").c_str());
  p << "ENSPAN" << std::endl;
  p << "inp_file\t" << inp_file << std::endl;
  p << "out_file\t" << out_file << std::endl;
  p << "qual\t" << qual << std::endl;
  p << "w\t" << w << std::endl;
  p << "build\t" << build << std::endl;
}

void write_enspan_out(std::string out_file, std::vector<cluster_info> &cluster, bool filter)
{
  static const char *head = "Fusion_Type\tBreakPoint1\tBreakPoint2\tGene1\tBreakPoint_Info_Pair1\tGene2\tBreakPoint_Info_Pair2\tN_DRP\tN_SR\t"
                            "BreakPoint1_Depth\tBreakPoint2_Depth\tBreakPoint1_AF\tBreakPoint2_AF\tBP1_Neighbour_Seq\tBP2_Neighbour_Seq\n";
  std::sort(cluster.begin(), cluster.end(), cmp_cluster);          // the same unstable library sort on the same order (src/BreakID.cc:1188)
  auto row = [](std::ofstream &o, const cluster_info &c) {
    o << c.fusion_type << "\t" << c.p1_chr << ":" << c.p1_exact_pos << "\t" << c.p2_chr << ":" << c.p2_exact_pos << "\t";
    o << c.p1_behalf_gene << "\t" << c.p1_strand << ":" << c.p1_exon_info << "\t" << c.p2_behalf_gene << "\t" << c.p2_strand << ":" << c.p2_exon_info << "\t";
    o << c.n_discordant_pair << "\t" << c.n_split_read << "\t" << c.p1_bp_depth << "\t" << c.p2_bp_depth << "\t";
    o << c.p1_alle_freq << "\t" << c.p2_alle_freq << "\t" << c.p1_rpt << "\t" << c.p2_rpt << "\n";
  };
  std::ofstream all, flt;
  if (!filter) { all.open((out_file + "_fusion_all.txt").c_str()); all << head; }
  flt.open((out_file + "_fusion.txt").c_str());
  flt << head;
  for (const cluster_info &c : cluster) {
    bool called = c.n_split_read > 0 && c.p1_exact_pos != (uint32_t)-1 && c.p2_exact_pos != -1;
    bool keep = called && !(c.p1_behalf_gene == "intergenic" && c.p2_behalf_gene == "intergenic") && c.p1_behalf_gene != c.p2_behalf_gene && !c.is_rpt;
    if (keep) row(flt, c);
    if (!filter && called) row(all, c);
  }
}

void breakid_compat_close() { g_session.reset(); }

// ---------------------------------------------------------------------------------------------------------------
// src/util_cluster.h:75
// ---------------------------------------------------------------------------------------------------------------
double euclidean_distance(coordinate &a, coordinate &b) { return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y)); }

void init_cluster(cluster_struct &main_cluster, long distance_threshold, std::vector<point> &points, int linkage_type)
{
  if (linkage_type != 1) throw std::invalid_argument("init_cluster: only linkage type 1 is implemented");
  Session &s = current();
  std::vector<uint32_t> x(points.size()), y(points.size()), idx;
  for (size_t i = 0; i < points.size(); ++i) {
    if (points[i].pos.x < 0 || points[i].pos.x > 4294967295.0 || points[i].pos.y < 0 || points[i].pos.y > 4294967295.0 || points[i].pos.x != floor(points[i].pos.x) ||
        points[i].pos.y != floor(points[i].pos.y))
      throw std::invalid_argument("init_cluster: coordinates must be genome positions (integers below 2^32)");
    x[i] = (uint32_t)points[i].pos.x; y[i] = (uint32_t)points[i].pos.y;
  }
  std::vector<int32_t> cl;
  int32_t roots = 0;
  cluster_points(s, 0, x, y, (double)distance_threshold, idx, cl, roots);
  main_cluster = cluster_struct();
  main_cluster.num_points = points.size();
  std::vector<char> merged(points.size(), 0);
  for (uint32_t i : idx) merged[i] = 1;
  auto centre = [&](node &nd) {
    double sx = 0, sy = 0;
    for (int j : nd.points) { sx += points[(size_t)j].pos.x; sy += points[(size_t)j].pos.y; }
    nd.centroid.x = sx / nd.num_points; nd.centroid.y = sy / nd.num_points;
  };
  for (size_t i = 0; i < points.size(); ++i) {            // leaves nobody merged with stay roots, at their own index order
    if (merged[i]) continue;
    node nd;
    nd.type = LEAF_NODE; nd.is_root = 1; nd.num_points = 1; nd.points.push_back((int)i); nd.label = points[i].label;
    centre(nd);
    main_cluster.nodes.push_back(nd);
  }
  for (size_t i = 0; i < idx.size();) {                   // merged roots in creation order = cluster id order
    size_t j = i;
    node nd;
    nd.type = MERGER; nd.is_root = 1;
    while (j < idx.size() && cl[j] == cl[i]) { nd.points.push_back((int)idx[j]); ++j; }
    nd.num_points = (int)nd.points.size();
    nd.height = 1;
    centre(nd);
    main_cluster.nodes.push_back(nd);
    i = j;
  }
  main_cluster.num_nodes = (int)main_cluster.nodes.size();
  main_cluster.num_root_clusters = main_cluster.num_nodes;
  for (size_t i = 0; i < idx.size(); ++i) points[idx[i]].cluster_id = cl[i];
  (void)roots;
}

int print_root_nodes(cluster_struct &main_cluster)
{
  int k = 0;
  for (int i = 0; i < main_cluster.num_nodes; ++i) k += main_cluster.nodes[(size_t)i].is_root ? 1 : 0;
  return k;
}

// ---------------------------------------------------------------------------------------------------------------
// src/util_bam.h:56-61
// ---------------------------------------------------------------------------------------------------------------
uint32_t combine_genome_chr_pos(bam_header_t *header, int chromID, int32_t position)
{
  uint32_t g = 0;
  for (int t = 0; t < chromID; ++t) g += header->target_len[t];
  return g + (uint32_t)position;
}

static std::string nib_bases(const std::string &chrom, long first0, long n, const std::string &dir)
{
  nib nb;
  nb.open(dir + "/hg19_" + chrom + ".nib");
  std::string out;
  char base = 0;                                   // a failed getBase leaves the previous base in place, like the reference's loop
  for (long i = first0; i < first0 + n; ++i) { nb.getBase(&base, (unsigned long)i); out += base; }
  return out;
}
// getBase takes 0-based positions: the right neighbour of 1-based p starts at 0-based p (src/util_bam.cc:78-96),
// the left neighbour covers 0-based [p-length-1, p-1) (src/util_bam.cc:105-122)
std::string get_right_neighbor_sequence_nib(std::string chrom, int32_t pos_1based, int length, std::string nib) { return nib_bases(chrom, pos_1based, length, nib); }
std::string get_left_neighbor_sequence_nib(std::string chrom, int32_t pos_1based, int length, std::string nib) { return nib_bases(chrom, (long)pos_1based - length - 1, length, nib); }
std::string get_sequence_nib(std::string chrom, int32_t start_1based, int32_t end_1based, std::string nib)
{
  return end_1based < start_1based ? std::string() : nib_bases(chrom, (long)start_1based - 1, (long)end_1based - start_1based + 1, nib);
}

std::string chromID2ChrName(int refID)
{
  if (refID == 23) return "chrY";
  if (refID == 22) return "chrX";
  if (refID >= 0 && refID < 22) return "chr" + std::to_string(refID + 1);
  return "";
}

// ---------------------------------------------------------------------------------------------------------------
// src/util_bed.h:22-31 (string helpers)
// ---------------------------------------------------------------------------------------------------------------
int find_longest_repeat_substring(const std::string &s)
{
  size_t best = 0;
  for (size_t i = 0; i < s.size();) {
    size_t j = i + 1;
    while (j < s.size() && s[j] == s[i]) ++j;
    if (j - i > best) best = j - i;
    i = j;
  }
  return (int)best;
}

std::vector<std::string> split_string(const std::string &s, const std::string &delim)
{
  std::vector<std::string> out;
  if (delim.empty()) { out.push_back(s); return out; }
  for (size_t a = 0;;) {
    size_t b = s.find(delim, a);
    std::string piece = s.substr(a, b == std::string::npos ? std::string::npos : b - a);
    if (!piece.empty()) out.push_back(piece);
    if (b == std::string::npos) break;
    a = b + delim.size();
  }
  return out;
}
