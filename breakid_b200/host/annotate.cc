#include "annotate.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

static std::vector<uint32_t> split_u32(const std::string &s)
{
  std::vector<uint32_t> r;
  std::string cur;
  for (char c : s) {
    if (c == ',') { if (!cur.empty()) r.push_back((uint32_t)atol(cur.c_str())); cur.clear(); }
    else cur += c;
  }
  if (!cur.empty()) r.push_back((uint32_t)atol(cur.c_str()));
  return r;
}

bool load_refgene(const std::string &path, std::vector<Transcript> &out)
{
  out.clear();
  std::ifstream in(path);
  if (!in.is_open()) return false;
  std::string line;
  while (std::getline(in, line, '\n')) {
    std::vector<std::string> f;
    {
      std::stringstream ss(line);
      std::string t;
      while (std::getline(ss, t, '\t')) f.push_back(t);
    }
    f.resize(16);
    if (f[1].find("NR_") != std::string::npos) continue;
    Transcript t;
    t.id = f[1]; t.chrom = f[2]; t.strand = f[3];
    t.tx_start = (uint32_t)atol(f[4].c_str()); t.tx_end = (uint32_t)atol(f[5].c_str());
    t.cds_start = (uint32_t)atol(f[6].c_str()); t.cds_end = (uint32_t)atol(f[7].c_str());
    uint32_t n_exon = (uint32_t)atol(f[8].c_str());
    std::vector<uint32_t> es = split_u32(f[9]), ee = split_u32(f[10]);
    t.gene = f[12];
    // coding parts of every exon that overlaps the CDS
    if (t.cds_start != t.cds_end) {
      for (uint32_t i = 0; i < n_exon && i < es.size() && i < ee.size(); ++i) {
        uint32_t s = es[i], e = ee[i];
        if (!(s < t.cds_end && e > t.cds_start)) continue;
        uint32_t cs, ce;
        if (s < t.cds_start && e > t.cds_start && e <= t.cds_end) { cs = t.cds_start; ce = e; }
        else if (s < t.cds_end && e > t.cds_end && s >= t.cds_start) { cs = s; ce = t.cds_end; }
        else if (e > t.cds_end && s < t.cds_start) { cs = t.cds_start; ce = t.cds_end; }
        else { cs = s; ce = e; }
        t.coding_parts.push_back(cs); t.coding_parts.push_back(ce);
        t.cdna_len += (long)ce - (long)cs;
      }
      t.coding_exons = (int)(t.coding_parts.size() / 2);
    }
    out.push_back(t);
  }
  return true;
}

SideAnnotation annotate_side(const std::vector<Transcript> &tx, const std::string &chrom, long pos)
{
  SideAnnotation a;
  if (pos == -1) { a.gene = "."; a.exon_info = "."; a.strand = "."; return a; }
  const Transcript *pick = nullptr;
  bool any = false;
  for (const Transcript &t : tx)
    if (chrom == t.chrom && pos >= (long)t.tx_start && pos <= (long)t.tx_end) {
      any = true;
      if (t.cdna_len > 0) pick = &t;                 // every CDS-bearing hit overwrites: the last one wins
    }
  if (!any) { a.gene = "intergenic"; a.exon_info = "."; a.strand = "."; return a; }
  if (!pick) {
    // the reference dereferences an empty transcript here (src/BreakID.cc:1757 underflow); out of its domain
    a.gene = ""; a.strand = ""; a.exon_info = ":0-0";
    return a;
  }
  int e0 = 0, e1 = 0;
  const std::vector<uint32_t> &p = pick->coding_parts;
  for (size_t i = 0; i + 1 < p.size(); ++i)
    if (pos >= (long)p[i] && pos <= (long)p[i + 1]) {
      int k = (int)(i / 2) + 1;
      if (pick->strand == "+") {
        if (i % 2 == 1) { e0 = k; e1 = k + 1; } else { e0 = k; e1 = k; }
      }
      if (pick->strand == "-") {
        if (i % 2 == 1) { e0 = pick->coding_exons + 1 - (k + 1); e1 = pick->coding_exons + 1 - k; }
        else { e0 = pick->coding_exons + 1 - (k + 1); e1 = pick->coding_exons + 1 - (k + 1); }
      }
      break;
    }
  a.gene = pick->gene; a.strand = pick->strand;
  a.exon_info = pick->id + ":" + std::to_string(e0) + "-" + std::to_string(e1);
  return a;
}
