// bkid_align.cuh -- banded alignment operator (BASELINE.json north_star kernel 4; an EXTENSION: the reference has
// no sequence alignment).  Included by bkid_core.cu.
//
// One warp per (clipped read segment, reference window) pair.  Both strings are staged in shared memory; lane l owns
// diagonal k = l - w of the band |j - i| <= w and the warp sweeps the ANTI-DIAGONALS t = i + j: the cells of one
// anti-diagonal are independent, a lane is active on every second step (parity of t - k), and each lane needs a
// single register -- its own last value is D(i-1, j-1) (two steps ago, same diagonal) while D(i-1, j) and D(i, j-1)
// are the last values of the two neighbouring lanes, fetched with one shuffle each.  Integer unit-cost edit distance
// (match 0, mismatch / insertion / deletion 1, N never matches); cells outside the band are +inf.
// Oracle: oracle/oracle.cc orc_banded_edit (plain row-by-row DP).
#pragma once

constexpr int AL_WARPS = 8;
constexpr int AL_MAXLEN = 512;
constexpr int AL_INF = 1 << 20;

// the wavefront itself: strings already staged in shared memory; every lane returns the distance (or -1)
__device__ __forceinline__ int banded_edit_warp(const uint8_t *__restrict__ sq, int nq, const uint8_t *__restrict__ sr, int nr, int w)
{
  if (nr - nq > w || nq - nr > w) return -1;                              // the end cell lies outside the band
  const int lane = threadIdx.x & 31;
  const int k = lane - w;                                                 // this lane's diagonal: j = i + k
  const bool in_band = lane <= 2 * w;
  int val = AL_INF;
  for (int t = 0; t <= nq + nr; ++t) {
    int up = __shfl_down_sync(0xffffffffu, val, 1);                       // D(i-1, j): diagonal k+1, previous step
    int left = __shfl_up_sync(0xffffffffu, val, 1);                       // D(i, j-1): diagonal k-1, previous step
    if (lane == 31 || lane >= 2 * w) up = AL_INF;
    if (lane == 0) left = AL_INF;
    if (in_band && ((t - k) & 1) == 0) {
      int i = (t - k) >> 1, j = i + k;                                    // arithmetic shift: t - k may be negative
      int nv = AL_INF;
      if (i >= 0 && j >= 0 && i <= nq && j <= nr) {
        if (i == 0) nv = j;
        else if (j == 0) nv = i;
        else {
          uint8_t a = sq[i - 1], b = sr[j - 1];
          int sub = (a == b && a != 'N') ? 0 : 1;
          nv = min(val + sub, min(up, left) + 1);
        }
      }
      val = nv;
    }
  }
  return __shfl_sync(0xffffffffu, val, (nr - nq) + w);
}

__global__ void __launch_bounds__(AL_WARPS * 32)
op_banded_align(const uint8_t *__restrict__ q, const uint32_t *__restrict__ q_off, const uint8_t *__restrict__ r, const uint32_t *__restrict__ r_off, long long n, int w,
                int32_t *__restrict__ out)
{
  __shared__ uint8_t sq[AL_WARPS][AL_MAXLEN], sr[AL_WARPS][AL_MAXLEN];
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long p = (long long)blockIdx.x * AL_WARPS + wi; p < n; p += (long long)gridDim.x * AL_WARPS) {
    const int nq = (int)(q_off[p + 1] - q_off[p]), nr = (int)(r_off[p + 1] - r_off[p]);
    int res;
    if (nq > AL_MAXLEN || nr > AL_MAXLEN) res = -2;                       // beyond the staged window size
    else {
      __syncwarp();
      for (int i = lane; i < nq; i += 32) sq[wi][i] = q[q_off[p] + i];
      for (int i = lane; i < nr; i += 32) sr[wi][i] = r[r_off[p] + i];
      __syncwarp();
      res = banded_edit_warp(sq[wi], nq, sr[wi], nr, w);
    }
    if (lane == 0) out[p] = res;
  }
}

// ---- evidence validator (bkid_params.validate_align; default off, extension) ---------------------------------------
// A split alignment says: the soft-clipped bases of this record align at SA:(chr, pos, strand) with SA:cigar.  One warp
// per SA-tagged record re-checks that claim: the clipped bases (reverse-complemented when the two alignments are on
// different strands) against the reference window [sa_pos, sa_pos + SA matches) from the 4-bit nib genome, banded edit
// distance with band 12 (the complementary-cigar test already allows the two lengths to differ by 10).  A row whose
// clipped bases do not align within len/10 + 2 edits stops being evidence (ok = 0); rows that cannot be checked (OC tag,
// no nib for the target, strings beyond the staged window) are left as they are.
constexpr int AL_BAND = 12;
__device__ __forceinline__ uint8_t bam_base(const uint8_t *__restrict__ seq4, int i)
{
  int b = seq4[i >> 1];
  int x = (i & 1) ? (b & 0xf) : (b >> 4);                                 // "=ACMGRSVTWYHKDBN", high nibble first
  return x == 1 ? 'A' : x == 2 ? 'C' : x == 4 ? 'G' : x == 8 ? 'T' : 'N';
}
__device__ __forceinline__ uint8_t comp_base(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

__global__ void __launch_bounds__(AL_WARPS * 32)
k7_validate_rows(const uint32_t *__restrict__ sa_rec, long long n_sa, const uint16_t *__restrict__ flag, const uint32_t *__restrict__ cig_off, const uint32_t *__restrict__ cig_ops,
                 const uint32_t *__restrict__ sa_off, const uint8_t *__restrict__ sa_txt, const uint32_t *__restrict__ oc_off, const uint32_t *__restrict__ seq_off,
                 const uint8_t *__restrict__ seq4, const int32_t *__restrict__ seq_len, const uint64_t *__restrict__ canon, int nt, const uint8_t *const *__restrict__ nib,
                 const uint64_t *__restrict__ nib_len, EvRow *__restrict__ rows)
{
  __shared__ uint8_t sq[AL_WARPS][AL_MAXLEN], sr[AL_WARPS][AL_MAXLEN];
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long k = (long long)blockIdx.x * AL_WARPS + wi;
  if (k >= n_sa || !rows[k].ok) return;
  if (oc_off[k + 1] != oc_off[k]) return;                                 // the record's own cigar is not the one that was tested
  uint32_t i = sa_rec[k];
  Roller rec;
  roller_set_bam(rec, cig_ops + cig_off[k], cig_off[k + 1] - cig_off[k]);
  int lseq = seq_len[k];
  int q0, qn;
  if (rec.begin_clips != 0) { q0 = 0; qn = rec.begin_clips; }
  else if (rec.end_clips != 0) { qn = rec.end_clips; q0 = lseq - qn; }
  else return;
  if (qn <= 0 || q0 < 0 || q0 + qn > lseq || qn > AL_MAXLEN) return;
  // SA fields 0..3 of the first entry (same tokenisation as k7_evidence_rows)
  const uint8_t *sa = sa_txt + sa_off[k]; uint32_t sal = sa_off[k + 1] - sa_off[k];
  uint32_t fs[4], fe[4]; int nf = 0;
  {
    uint32_t p = 0;
    while (p < sal && nf < 4) {
      while (p < sal && sa[p] == ',') ++p;
      if (p >= sal) break;
      uint32_t e = p;
      while (e < sal && sa[e] != ',') ++e;
      fs[nf] = p; fe[nf] = e; ++nf;
      p = e;
    }
  }
  if (nf < 4) return;
  Roller sac;
  roller_set_text(sac, sa + fs[3], fe[3] - fs[3]);
  int rn = sac.matches;
  if (rn <= 0 || rn > AL_MAXLEN) return;
  long long sv = 0;
  for (uint32_t p = fs[1]; p < fe[1] && sa[p] >= '0' && sa[p] <= '9'; ++p) { sv = sv * 10 + (sa[p] - '0'); if (sv > 0x7fffffffll) sv = 0x7fffffffll; }
  uint64_t code = chr_code(sa + fs[0], fe[0] - fs[0]);
  int t = -1;
  for (int x = 0; x < nt; ++x) if (canon[x] == code) { t = x; break; }
  if (t < 0 || !nib || !nib[t]) return;
  bool own_minus = (flag[i] & F_REVERSE) != 0, sa_minus = (fe[2] > fs[2]) && sa[fs[2]] == '-';
  bool rc = own_minus != sa_minus;
  const uint8_t *s4 = seq4 + seq_off[k];
  __syncwarp();
  for (int j = lane; j < qn; j += 32) sq[wi][j] = rc ? comp_base(bam_base(s4, q0 + qn - 1 - j)) : bam_base(s4, q0 + j);
  for (int j = lane; j < rn; j += 32) sr[wi][j] = (uint8_t)nib_base(nib[t], nib_len[t], sv - 1 + j, 'N');
  __syncwarp();
  int d = banded_edit_warp(sq[wi], qn, sr[wi], rn, AL_BAND);
  if (lane == 0 && (d < 0 || d * 10 > qn + 20)) rows[k].ok = 0;
}
