// annotate.h -- refGene transcript lookup for the Gene / BreakPoint_Info columns of the call file.
// Host-side (tiny, string heavy); reproduces the reference's observable rules:
//   * rows whose transcript id contains "NR_" are skipped (src/RefSeqTranscript.cc:220,239-242)
//   * UTRs are stripped from exon coordinates (src/RefSeqTranscript.cc:92-139)
//   * containment test is inclusive on both ends (src/BreakID.cc:1555-1556)
//   * the LAST overlapping transcript with a CDS wins (src/RefSeqTranscript.cc:311-320)
//   * exon numbering of src/BreakID.cc:1753-1793
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct Transcript {
  std::string id, chrom, strand, gene;
  uint32_t tx_start = 0, tx_end = 0, cds_start = 0, cds_end = 0;
  int coding_exons = 0;
  long cdna_len = 0;
  std::vector<uint32_t> coding_parts;    // start0, end0, start1, end1, ...
};

struct SideAnnotation {
  std::string gene, exon_info, strand;   // "intergenic"/"."/"." when nothing overlaps
};

// returns false when the file cannot be opened
bool load_refgene(const std::string &path, std::vector<Transcript> &out);
SideAnnotation annotate_side(const std::vector<Transcript> &tx, const std::string &chrom, long pos);
