/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-old-data layouts shared by
 *   - oracle/oracle.cc     : this repo's CPU restatement of the BreakID hot path
 *                            (exports orc_* with C linkage -> oracle/liboracle.so), and
 *   - oracle/ref_shim.cc   : thin C wrappers around the REAL reference functions, compiled
 *                            from /root/reference/src where they lie (exports ref_* ->
 *                            oracle/_ref/libbreakid_ref.so).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load these libraries.  The product (breakid_b200/) never links or calls them.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One discordant pair as emitted by scan_discordant_pairs (reference src/BreakID.cc:1428-1480,
 * struct discordant_pair src/BreakID.h:39-58).  name_lo/name_hi = 128-bit hash of the read
 * name (see orc_name_hash); tids are header indices of p?_chr (-1 for "*"). */
typedef struct {
  uint64_t name_lo, name_hi;
  int32_t  p1_tid, p2_tid;
  uint32_t p1_pos, p2_pos;          /* 1-based positions (pos+1) */
  uint32_t p1_chr_pos, p2_chr_pos;  /* genome-wide, uint32 wrap (src/util_bam.cc:57-68) */
  uint16_t p1_flag, p2_flag;
  uint8_t  p1_mapq, p2_mapq;
  uint8_t  p1_strand, p2_strand;    /* '+' / '-' */
  int32_t  bucket;                  /* rank of "chrA_chrB" in std::map<string> order */
  int32_t  cluster;                 /* filled by the clustering stage */
  uint32_t orig;                    /* index in scan emission order (test bookkeeping) */
  uint32_t _pad;
} orc_pair;

/* One split-read evidence row (struct split_align_pair, src/BreakID.h:116-133). chr ids are the
 * t in 0..23 with chromID2ChrName(t) == name (src/util_bam.cc:128-142), -1 = "" (no such t),
 * -2-k = some other string (k = per-call id so that equal strings get equal ids). */
typedef struct {
  uint64_t name_lo, name_hi;
  int32_t  primary_chr, secondary_chr;
  uint32_t primary_start, secondary_start, primary_end, secondary_end;
  uint32_t primary_bp, secondary_bp;
  uint64_t primary_cigar_h, secondary_cigar_h;   /* FNV-1a of the cigar *string* bytes */
  uint16_t flag;
  uint8_t  secondary;
  uint8_t  _pad[5];
} orc_evidence;

/* Final per-cluster record (struct cluster_info, src/BreakID.h:60-113) -- the numeric part. */
typedef struct {
  int32_t  bucket;
  int32_t  id;
  int32_t  p1_tid, p2_tid;
  uint64_t p1_mean_pos, p2_mean_pos;
  uint32_t p1_min_pos, p1_max_pos, p2_min_pos, p2_max_pos;
  uint32_t p1_exact_pos;
  int32_t  p2_exact_pos;
  int64_t  n_split_read, n_discordant_pair;
  double   p1_bp_depth, p2_bp_depth;
  float    p1_alle_freq, p2_alle_freq;
  int32_t  fusion_type;      /* 0 Unknown 1 Translocation 2 Inversion 3 Duplication 4 Deletion */
  int32_t  is_rpt;
  char     p1_rpt[44], p2_rpt[44];   /* 41-mers, NUL padded */
} orc_cluster;

/* 128-bit read-name hash used everywhere a read name is compared for equality
 * (mate join src/BreakID.cc:1424, split-read name match src/BreakID.cc:605).  The product's
 * host decoder implements the same function (include/breakid_b200.h: bkid_name_hash). */
static inline void orc_name_hash(const char *s, uint64_t *lo, uint64_t *hi)
{
  uint64_t a = 0xcbf29ce484222325ULL, b = 0x9E3779B97F4A7C15ULL;
  for (; *s; ++s) {
    uint64_t c = (unsigned char)*s;
    a = (a ^ c) * 0x100000001b3ULL;
    b = (b ^ c) * 0xff51afd7ed558ccdULL;
    b ^= b >> 32;
  }
  *lo = a; *hi = b;
}
static inline uint64_t orc_str_hash(const char *s)
{
  uint64_t a = 0xcbf29ce484222325ULL;
  for (; *s; ++s) a = (a ^ (unsigned char)*s) * 0x100000001b3ULL;
  return a;
}

/* banded unit-cost edit distance (checker of bkid_op_banded_align; an extension, nothing in the reference to follow):
 * plain row-by-row DP over the band |j - i| <= w; -1 when |nr - nq| > w */
int orc_banded_edit(const unsigned char *q, int nq, const unsigned char *r, int nr, int w);

#ifdef __cplusplus
}
#endif
