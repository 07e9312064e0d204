// prims.cuh -- device-wide building blocks written for this pipeline (sm_100a):
//   * exclusive scan (uint32 -> uint32 / uint64 totals)
//   * stable LSD radix sort on (uint64 key, uint32 value) with a selectable bit range
//   * small helpers (launch counting, error checks)
// Everything is HBM-bound integer work: 128-bit loads where alignment allows, grids sized from the
// SM count, no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define BK_CUDA(x)                                                          \
  do {                                                                      \
    cudaError_t e__ = (x);                                                  \
    if (e__ != cudaSuccess) { bk_set_cuda_error(e__, __FILE__, __LINE__); return BKID_ERR_CUDA; } \
  } while (0)

void bk_set_cuda_error(cudaError_t e, const char *file, int line);
extern thread_local long long g_bk_launches;
#define BK_LAUNCH(kernel, grid, block, smem, stream, ...)                   \
  do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); ++g_bk_launches; } while (0)

namespace bk {

static inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// warp / block helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v)
{
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane_id() >= (unsigned)o) v += u;
  }
  return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide exclusive scan; blockDim.x must be a multiple of 32 and <= 1024.  `total` gets the
// block sum.  sh must hold 33 elements of T.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *sh, T &total)
{
  unsigned w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
  T inc = warp_incl_scan(v);
  if (l == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = l < nw ? sh[l] : T(0);
    T si = warp_incl_scan(s);
    sh[l] = si - s;
    if (l == 31) sh[32] = si;
  }
  __syncthreads();
  T r = inc - v + sh[w];
  total = sh[32];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan of uint32 counts -> uint64-safe offsets (stored as uint32 when the
// total fits; callers here never exceed 2^32 items).
// 3 kernels: tile sums, scan of tile sums (one block), tile scan with carried offset.
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename InT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const InT *__restrict__ in, long long n, unsigned long long *__restrict__ tile_sum)
{
  __shared__ unsigned long long sh[33];
  long long base = (long long)blockIdx.x * SCAN_TILE;
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += (unsigned long long)in[i];
  }
  unsigned long long tot;
  block_excl_scan<unsigned long long>(s, sh, tot);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_tile_offsets(unsigned long long *__restrict__ tile_sum, int ntiles, unsigned long long *__restrict__ total_out)
{
  __shared__ unsigned long long sh[33];
  unsigned long long carry = 0;
  for (int b = 0; b < ntiles; b += 1024) {
    int i = b + threadIdx.x;
    unsigned long long v = i < ntiles ? tile_sum[i] : 0ull, tot;
    unsigned long long ex = block_excl_scan<unsigned long long>(v, sh, tot);
    if (i < ntiles) tile_sum[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename InT, typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const InT *in, long long n, const unsigned long long *__restrict__ tile_off, OutT *out)
{
  // each thread owns SCAN_ITEMS consecutive elements so the scan order is the array order
  __shared__ unsigned long long sh[33];
  long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { long long i = base + k; v[k] = i < n ? (unsigned long long)in[i] : 0ull; s += v[k]; }
  unsigned long long tot;
  unsigned long long ex = block_excl_scan<unsigned long long>(s, sh, tot) + tile_off[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { long long i = base + k; if (i < n) out[i] = (OutT)ex; ex += v[k]; }
}

// exclusive scan; tmp must hold div_up(n, SCAN_TILE)+1 uint64.  total_out (device) optional.
template <typename InT, typename OutT>
static inline void exclusive_scan(const InT *in, OutT *out, long long n, unsigned long long *tmp, unsigned long long *total_out, cudaStream_t st)
{
  if (n <= 0) { if (total_out) cudaMemsetAsync(total_out, 0, 8, st); return; }
  int ntiles = div_up(n, SCAN_TILE);
  BK_LAUNCH((scan_tile_sums<InT>), ntiles, SCAN_THREADS, 0, st, in, n, tmp);
  BK_LAUNCH(scan_tile_offsets, 1, 1024, 0, st, tmp, ntiles, total_out);
  BK_LAUNCH((scan_tiles<InT, OutT>), ntiles, SCAN_THREADS, 0, st, in, n, tmp, out);
}
static inline size_t scan_tmp_elems(long long n) { return (size_t)div_up(n > 0 ? n : 1, SCAN_TILE) + 1; }

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, (uint64 key, uint32 value).
// Per pass: histogram per chunk (RS_CHUNK elements per block), column scan, ranked scatter.
// Ranking inside a block: elements are taken in rounds of blockDim; within a round warp w lane l
// holds element w*32+l; __match_any_sync groups equal digits inside a warp; per-(round,warp) digit
// counts are prefix-summed per digit by one thread per digit.  This is the classic stable
// multi-split; it keeps the pass at one read of keys for the histogram and one read+write of
// keys+values for the scatter.
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;                       // elements per thread per block
constexpr int RS_CHUNK = RS_THREADS * RS_ROUNDS;   // 2048 elements per block
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const uint64_t *__restrict__ keys, long long n, int shift, uint32_t *__restrict__ hist /*[256][nblocks]*/, int nblocks)
{
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  long long base = (long long)blockIdx.x * RS_CHUNK;
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    long long i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, long long n, int shift,
                                                         const uint32_t *__restrict__ hist_scanned, int nblocks,
                                                         uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
  __shared__ uint16_t cnt[RS_ROUNDS * RS_WARPS][256];   // 32 KB
  __shared__ uint32_t goff[256];
  for (int i = threadIdx.x; i < RS_ROUNDS * RS_WARPS * 256 / 2; i += RS_THREADS) ((uint32_t *)cnt)[i] = 0;
  goff[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  long long base = (long long)blockIdx.x * RS_CHUNK;
  unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  uint64_t k[RS_ROUNDS];
  uint32_t v[RS_ROUNDS];
  uint16_t rank_in_warp[RS_ROUNDS];
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    long long i = base + r * RS_THREADS + threadIdx.x;
    bool ok = i < n;
    k[r] = ok ? keys[i] : ~0ull;
    v[r] = ok ? vals[i] : 0u;
    unsigned d = ok ? (unsigned)((k[r] >> shift) & 255u) : 256u;
    unsigned m = __match_any_sync(0xffffffffu, d);
    rank_in_warp[r] = (uint16_t)__popc(m & ((1u << l) - 1u));
    if (ok && rank_in_warp[r] == 0) cnt[r * RS_WARPS + w][d] = (uint16_t)__popc(m);
  }
  __syncthreads();
  {  // exclusive prefix over the (round, warp) slots for digit = threadIdx.x
    unsigned d = threadIdx.x, run = 0;
#pragma unroll 4
    for (int s = 0; s < RS_ROUNDS * RS_WARPS; ++s) { unsigned c = cnt[s][d]; cnt[s][d] = (uint16_t)run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ROUNDS; ++r) {
    long long i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) {
      unsigned d = (unsigned)((k[r] >> shift) & 255u);
      uint32_t dst = goff[d] + cnt[r * RS_WARPS + w][d] + rank_in_warp[r];
      keys_out[dst] = k[r];
      vals_out[dst] = v[r];
    }
  }
}

struct RadixTmp {
  uint64_t *keys_alt; uint32_t *vals_alt; uint32_t *hist; unsigned long long *scan_tmp;
};
static inline size_t radix_hist_elems(long long n) { return (size_t)256 * (size_t)div_up(n > 0 ? n : 1, RS_CHUNK); }

// sorts by bits [lo_bit, hi_bit) of the key; result is left in (keys, vals) (copies back if the
// number of passes is odd).  Stable.
static inline int radix_sort_pairs(uint64_t *keys, uint32_t *vals, long long n, int lo_bit, int hi_bit, const RadixTmp &t, cudaStream_t st)
{
  if (n <= 1) return 0;
  int nblocks = div_up(n, RS_CHUNK);
  uint64_t *ki = keys, *ko = t.keys_alt;
  uint32_t *vi = vals, *vo = t.vals_alt;
  int passes = 0;
  for (int shift = lo_bit; shift < hi_bit; shift += 8) {
    BK_LAUNCH(rs_histogram, nblocks, RS_THREADS, 0, st, ki, n, shift, t.hist, nblocks);
    exclusive_scan<uint32_t, uint32_t>(t.hist, t.hist, (long long)256 * nblocks, t.scan_tmp, nullptr, st);
    BK_LAUNCH(rs_scatter, nblocks, RS_THREADS, 0, st, ki, vi, n, shift, t.hist, nblocks, ko, vo);
    uint64_t *tk = ki; ki = ko; ko = tk;
    uint32_t *tv = vi; vi = vo; vo = tv;
    ++passes;
  }
  if (passes & 1) {
    cudaMemcpyAsync(keys, ki, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(vals, vi, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
  }
  return passes;
}

}  // namespace bk
