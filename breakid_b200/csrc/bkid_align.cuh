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
    else if (nr - nq > w || nq - nr > w) res = -1;                        // the end cell lies outside the band
    else {
      __syncwarp();
      for (int i = lane; i < nq; i += 32) sq[wi][i] = q[q_off[p] + i];
      for (int i = lane; i < nr; i += 32) sr[wi][i] = r[r_off[p] + i];
      __syncwarp();
      const int k = lane - w;                                             // this lane's diagonal: j = i + k
      const bool in_band = lane <= 2 * w;
      int val = AL_INF;
      for (int t = 0; t <= nq + nr; ++t) {
        int up = __shfl_down_sync(0xffffffffu, val, 1);                   // D(i-1, j): diagonal k+1, previous step
        int left = __shfl_up_sync(0xffffffffu, val, 1);                   // D(i, j-1): diagonal k-1, previous step
        if (lane == 31 || lane >= 2 * w) up = AL_INF;
        if (lane == 0) left = AL_INF;
        if (in_band && ((t - k) & 1) == 0) {
          int i = (t - k) >> 1, j = i + k;                                // arithmetic shift: t - k may be negative
          int nv = AL_INF;
          if (i >= 0 && j >= 0 && i <= nq && j <= nr) {
            if (i == 0) nv = j;
            else if (j == 0) nv = i;
            else {
              uint8_t a = sq[wi][i - 1], b = sr[wi][j - 1];
              int sub = (a == b && a != 'N') ? 0 : 1;
              nv = min(val + sub, min(up, left) + 1);
            }
          }
          val = nv;
        }
      }
      res = __shfl_sync(0xffffffffu, val, (nr - nq) + w);
    }
    if (lane == 0) out[p] = res;
  }
}
