// include/compat/BreakID_stages.h -- the stage functions the reference's main() calls (src/BreakID.h:157-228,
// src/BreakID.cc:94-170), with the reference's signatures, implemented over the B200 C ABI (include/breakid_b200.h) by
// breakid_b200/host/compat_stages.cc (libbreakid_compat.so).
//
//   reference stage (src/BreakID.cc)              what runs
//   --------------------------------------------  ----------------------------------------------------------------
//   get_mean_insert_size            :1909-1950    BGZF inflate + BAM decode + K1 + the truncating sd pass, on the GPU
//   scan_discordant_pairs           :1363-1527    candidates -> mate join by read name -> "chrA_chrB" buckets, on the GPU
//   add_enspan_point_id             :1287-1293    host (labels)
//   remove_isolated_pairs           :1271-1285    bkid_op_remove_isolated (introsort replay + mask) on the GPU
//   find_cluster_pairs_enspan_ahc   :1304-1326    bkid_op_cluster mode 0 on the GPU
//   find_cluster_pairs_enspan_fast  :1046-1168    bkid_op_cluster mode 1 on the GPU
//   findClusterBreakPointInfoSaTag  :201-381      bkid_op_summarize + bkid_refine on the GPU, refGene annotation on the host
//   write_enspan_out / _params      :1170-1263    host
//
// The functions share one session per input BAM (the records stay resident on the device between the calls); the device is
// $BREAKID_GPU (default 0), refGene.txt is $BREAKID_REFGENE or $BREAKID_INSTALLDIR/ref_files/refGene.txt.
// Differences a caller can see: discordant_pair::qname holds the 128-bit read-name hash as 32 hex digits (names are not
// kept on the device), cluster_info::discordant_reads / split_reads stay empty, the order of the pairs INSIDE one cluster
// after find_cluster_pairs_* may differ from the reference's (nothing downstream depends on it), and `split_reads` is
// left empty.  Call files are byte-identical (tests/test_compat.py builds the reference's own main() against this header).
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "util_bam.h"
#include "util_cluster.h"

struct discordant_pair {
  std::string id, qname;
  long p1_flag = 0, p2_flag = 0;
  std::string p1_chr, p2_chr;
  uint32_t p1_pos = 0, p2_pos = 0;                 // 1-based
  long p1_mapq = 0, p2_mapq = 0;
  char p1_strand = '+', p2_strand = '+';
  int is_isolated = 0;
  int cluster = -1;
  std::string cluster_id;
  uint32_t p1_chr_pos = 0, p2_chr_pos = 0;         // genome-wide (combine_genome_chr_pos)
};

struct cluster_info {
  long id = 0;
  std::string p1_chr;
  uint64_t p1_mean_pos = 0;
  uint32_t p1_min_pos = 0, p1_max_pos = 0, p1_exact_pos = 0;
  std::string p2_chr;
  uint64_t p2_mean_pos = 0;
  uint32_t p2_min_pos = 0, p2_max_pos = 0;
  int32_t p2_exact_pos = 0;
  long n_split_read = 0, n_discordant_pair = 0;
  bool inv = false;
  std::string inv_type = ".", fusion_type = ".";
  std::string discordant_reads, split_reads;
  std::string p1_behalf_gene, p1_gene_part = "0", p1_genes;
  std::string p2_behalf_gene, p2_gene_part = "0", p2_genes;
  std::string p1_rpt, p2_rpt;
  std::string p1_strand, p2_strand;
  std::string p1_exon_info, p2_exon_info;
  std::string p1_part, p2_part;
  bool hotspot = false, cosmic = false;
  std::set<std::string> drp_type_set;
  std::string fusion_pair, p1_bp_exon, p2_bp_exon, up_gene, down_gene;
  bool sino_pair_match = false, cosmic_pair_match = false;
  double p1_bp_depth = 0, p2_bp_depth = 0, p1_coverage = 0, p2_coverage = 0;
  float p1_alle_freq = 0.0f, p2_alle_freq = 0.0f;
  bool is_rpt = false;
};

// orderings the reference sorts with (src/BreakID.h:170-188)
inline bool cmp_p1_enspan_pairs(discordant_pair a, discordant_pair b) { return a.p1_chr_pos < b.p1_chr_pos; }
inline bool cmp_p2_enspan_pairs(discordant_pair a, discordant_pair b) { return a.p2_chr_pos < b.p2_chr_pos; }
inline bool cmp_enspan_id(discordant_pair a, discordant_pair b) { return a.cluster < b.cluster; }
inline bool cmp_cluster(cluster_info a, cluster_info b) { return a.n_discordant_pair > b.n_discordant_pair; }

void get_mean_insert_size(std::string input_bam, std::vector<double> &insert);
void scan_discordant_pairs(const std::string &inp_file, const std::string &build, long qual, double w,
                           std::map<std::string, std::vector<discordant_pair>> &enspan_map, std::string nib_dir);
void add_enspan_point_id(std::vector<discordant_pair> &enspan_vec);
void remove_isolated_pairs(std::vector<discordant_pair> &enspans, double w);
int find_cluster_pairs_enspan_ahc(std::vector<discordant_pair> &enspan, double distance_threshold, int distance_type, int min_reads_per_cluster);
int find_cluster_pairs_enspan_fast(std::vector<discordant_pair> &enspan, double w, int min_reads);
void build_pair_array(std::vector<discordant_pair> &enspan, std::vector<point> &points);
void add_cluster_id_for_enspan_vec(cluster_struct &main_cluster, std::vector<discordant_pair> &enspan, int min_reads_per_cluster);
void findClusterBreakPointInfoSaTag(std::string bam_file, std::vector<discordant_pair> &enspan, double w, std::vector<cluster_info> &cluster_vec,
                                    std::vector<bam1_t *> &split_reads, std::string nib_dir);
void annotate_cluster_for_sa_tag(std::vector<cluster_info> &clusters, std::string nib_dir);   // gene / exon columns (refGene); the 41-mers and is_rpt come from the device record
std::string determine_fusion_type_from_drp(cluster_info &cluster);
void write_enspan_out(std::string out_file, std::vector<cluster_info> &cluster, bool filter);
void write_enspan_params(std::string inp_file, std::string out_file, std::string build, double w, long qual);

// not part of the reference: release the session (device memory, the mapped BAM) before the process ends
void breakid_compat_close();
