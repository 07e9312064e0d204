#include "annotate.h"

#include <algorithm>
#include <cstdio>
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

static SideAnnotation describe(const std::vector<Transcript> &tx, long which, long pos)
{
  SideAnnotation a;
  if (which == -2) { a.gene = "intergenic"; a.exon_info = "."; a.strand = "."; return a; }
  if (which == -1) {
    // the reference dereferences an empty transcript here (src/BreakID.cc:1757 underflow); out of its domain
    a.gene = ""; a.strand = ""; a.exon_info = ":0-0";
    return a;
  }
  const Transcript *pick = &tx[(size_t)which];
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

SideAnnotation annotate_side(const std::vector<Transcript> &tx, const std::string &chrom, long pos)
{
  if (pos == -1) { SideAnnotation a; a.gene = "."; a.exon_info = "."; a.strand = "."; return a; }
  long pick = -2;
  for (size_t i = 0; i < tx.size(); ++i) {
    const Transcript &t = tx[i];
    if (chrom == t.chrom && pos >= (long)t.tx_start && pos <= (long)t.tx_end) {
      if (pick == -2) pick = -1;
      if (t.cdna_len > 0) pick = (long)i;            // every CDS-bearing hit overwrites: the last one wins
    }
  }
  return describe(tx, pick, pos);
}

void RefGeneIndex::build(const std::vector<Transcript> &tx)
{
  chroms.clear(); per_chrom.clear();
  for (size_t i = 0; i < tx.size(); ++i) {
    size_t c = 0;
    while (c < chroms.size() && chroms[c] != tx[i].chrom) ++c;
    if (c == chroms.size()) { chroms.push_back(tx[i].chrom); per_chrom.emplace_back(); }
    per_chrom[c].push_back(Entry{tx[i].tx_start, tx[i].tx_end, 0u, (uint32_t)i});
  }
  for (auto &v : per_chrom) {
    std::stable_sort(v.begin(), v.end(), [](const Entry &a, const Entry &b) { return a.start < b.start; });
    uint32_t m = 0;
    for (auto &e : v) { m = std::max(m, e.end); e.max_end_so_far = m; }
  }
}

long RefGeneIndex::lookup(const std::vector<Transcript> &tx, const std::string &chrom, long pos) const
{
  size_t c = 0;
  while (c < chroms.size() && chroms[c] != chrom) ++c;
  if (c == chroms.size() || pos < 0) return -2;
  const std::vector<Entry> &v = per_chrom[c];
  // entries with start <= pos form a prefix; walk it backwards while anything there can still reach pos
  size_t hi = (size_t)(std::upper_bound(v.begin(), v.end(), pos, [](long p, const Entry &e) { return p < (long)e.start; }) - v.begin());
  long pick = -2;
  for (size_t k = hi; k-- > 0;) {
    if ((long)v[k].max_end_so_far < pos) break;
    if ((long)v[k].end >= pos) {
      if (pick == -2) pick = -1;
      if (tx[v[k].order].cdna_len > 0 && (long)v[k].order > pick) pick = (long)v[k].order;      // last in file order wins
    }
  }
  return pick;
}

SideAnnotation annotate_side(const std::vector<Transcript> &tx, const RefGeneIndex &ix, const std::string &chrom, long pos)
{
  if (pos == -1) { SideAnnotation a; a.gene = "."; a.exon_info = "."; a.strand = "."; return a; }
  return describe(tx, ix.lookup(tx, chrom, pos), pos);
}

// test hook (libbreakid_host.so): both lookups on one query, "gene\tstrand\texon_info" each
extern "C" int bkid_host_annotate_both(const char *refgene, const char *chrom, long pos, char *linear, char *indexed, int cap)
{
  static std::string loaded;
  static std::vector<Transcript> tx;
  static RefGeneIndex ix;
  if (loaded != refgene) { if (!load_refgene(refgene, tx)) return -1; ix.build(tx); loaded = refgene; }
  SideAnnotation a = annotate_side(tx, chrom, pos), b = annotate_side(tx, ix, chrom, pos);
  snprintf(linear, cap, "%s\t%s\t%s", a.gene.c_str(), a.strand.c_str(), a.exon_info.c_str());
  snprintf(indexed, cap, "%s\t%s\t%s", b.gene.c_str(), b.strand.c_str(), b.exon_info.c_str());
  return 0;
}
