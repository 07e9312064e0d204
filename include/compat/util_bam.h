// include/compat/util_bam.h -- the helpers of the reference's src/util_bam.h:56-61 that the hot path touches, without a
// samtools / htslib dependency: BAM decode happens on the device (bkid_push_bgzf), so the record-level helper
// read_bam_reduced_record(bam1_t*, ...) has no counterpart here.
#pragma once
#include <cstdint>
#include <string>

// the three fields of samtools' bam_header_t the helpers read; a translation unit that already has <sam.h> keeps its own
#if !defined(BAM_BAM_H) && !defined(BAM_H) && !defined(HTSLIB_SAM_H)
struct bam_header_t { int32_t n_targets; char **target_name; uint32_t *target_len; };
struct bam1_t;                                       // opaque: the reference's main() only holds pointers to it
#endif

// genome-wide 0-based coordinate = sum of the lengths of the targets before chromID + position (uint32 wrap-around kept)
uint32_t combine_genome_chr_pos(bam_header_t *header, int chromID, int32_t position);
// `length` bases right of / left of / between 1-based positions from <nib>/hg19_<chrom>.nib, upper case
std::string get_right_neighbor_sequence_nib(std::string chrom, int32_t pos_1based, int length, std::string nib);
std::string get_left_neighbor_sequence_nib(std::string chrom, int32_t pos_1based, int length, std::string nib);
std::string get_sequence_nib(std::string chrom, int32_t start_1based, int32_t end_1based, std::string nib);
std::string chromID2ChrName(int refID);              // 0..21 -> chr1..chr22, 22 -> chrX, 23 -> chrY, else ""
