// prims.cuh -- device-wide building blocks written for this pipeline (sm_100a):
//   * exclusive scan (uint32 -> uint32 / uint64 totals)
//   * stable LSD radix sort on (uint64 key, uint32 value) with a selectable bit range
//   * small helpers (launch counting, error checks)
// Everything is HBM-bound integer work: 128-bit loads where alignment allows, grids sized from the
// SM count, no tensor cores.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#define BK_CUDA(x)                                                          \
  do {                                                                      \
    cudaError_t e__ = (x);                                                  \
    if (e__ != cudaSuccess) { bk_set_cuda_error(e__, __FILE__, __LINE__); return BKID_ERR_CUDA; } \
  } while (0)

void bk_set_cuda_error(cudaError_t e, const char *file, int line);
extern thread_local long long g_bk_launches;
// per-kernel device times (bkid_profile_kernels): when switched on, every launch is bracketed by two CUDA events on
// the stream it is launched on; off (the default) costs one predictable branch
extern bool g_bk_prof_on;
void bk_prof_begin(const char *name, cudaStream_t st);
void bk_prof_end(cudaStream_t st);
#define BK_LAUNCH(kernel, grid, block, smem, stream, ...)                   \
  do {                                                                      \
    if (g_bk_prof_on) bk_prof_begin(#kernel, (stream));                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    if (g_bk_prof_on) bk_prof_end((stream));                                \
    ++g_bk_launches;                                                        \
  } while (0)

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
// stable LSD radix sort on (uint64 key, uint32 value), 8-bit digits, ONESWEEP form:
//   os_hist       one pass over the keys builds the digit histograms of ALL passes (shared-memory atomics, one
//                 global atomicAdd per bin and CTA); os_hist_scan turns them into exclusive digit offsets
//   os_pass       one kernel per digit: a CTA takes the next tile from an atomic ticket, ranks its 4096 keys
//                 (match-any inside a warp, per-warp digit counters, no atomics), publishes its digit counts and
//                 resolves "keys with this digit in earlier tiles" by decoupled look-back over the tile status
//                 words (flag + count in ONE 32-bit word, so no fence protocol), reorders the tile in shared
//                 memory and writes every digit run as one contiguous, coalesced range.
// A 64-bit sort is 2 + 8 launches and reads the keys once for the histograms plus once per pass, instead of the
// 40 launches (histogram + 3-kernel scan + uncoalesced scatter per pass) of the round-1 LSD sort.
// The element count may live on the device (n_dev): grids are sized from a host upper bound, late tiles exit.
// Elements per sort < 2^30 (tile status words carry 30-bit counts); everything sorted here is candidate-sized.
// Measured and not kept (round 2, B200, 44 passes of the resident step = 0.88 ms with the look-back below): a
// warp-cooperative 32-tile window per digit (2.1 ms: 32 serial digits per warp cost every small sort eight round trips)
// and an 8-tile window per thread (2.3 ms: the spinning first wave re-reads eight status words per thread and starves L2).
// ---------------------------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_ITEMS = 16;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;          // 4096 keys per tile
constexpr int OS_MAX_PASSES = 8;
constexpr unsigned OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VAL_MASK = (1u << 30) - 1u;
constexpr size_t OS_SMEM = (size_t)OS_TILE * 12 + (size_t)OS_WARPS * 256 * 4 + 2 * 256 * 4 + 40 * 4;

__global__ void __launch_bounds__(256) os_hist(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ n_dev, uint32_t n_host, int lo_bit, int npass,
                                               uint32_t *__restrict__ ghist /* [npass][256] */)
{
  __shared__ uint32_t h[OS_MAX_PASSES][256];
  for (int p = 0; p < npass; ++p) h[p][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = n_dev ? *n_dev : n_host;
  for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < n; i += gridDim.x * 256u) {
    uint64_t k = keys[i] >> lo_bit;
    for (int p = 0; p < npass; ++p) { atomicAdd(&h[p][(uint32_t)k & 255u], 1u); k >>= 8; }
  }
  __syncthreads();
  for (int p = 0; p < npass; ++p) { uint32_t v = h[p][threadIdx.x]; if (v) atomicAdd(&ghist[p * 256 + threadIdx.x], v); }
}

__global__ void __launch_bounds__(256) os_hist_scan(uint32_t *__restrict__ ghist, int npass)
{
  __shared__ uint32_t sh[33];
  for (int p = 0; p < npass; ++p) {
    uint32_t v = ghist[p * 256 + threadIdx.x], tot;
    uint32_t ex = block_excl_scan<uint32_t>(v, sh, tot);
    ghist[p * 256 + threadIdx.x] = ex;
  }
}

__global__ void __launch_bounds__(OS_THREADS) os_pass(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout, uint32_t *__restrict__ vout,
                                                      const uint32_t *__restrict__ n_dev, uint32_t n_host, int shift, const uint32_t *__restrict__ ghist_excl,
                                                      uint32_t *__restrict__ status /* [ntiles][256], zeroed */, uint32_t *__restrict__ ticket)
{
  extern __shared__ __align__(16) unsigned char os_smem[];
  uint64_t *skey = reinterpret_cast<uint64_t *>(os_smem);
  uint32_t *sval = reinterpret_cast<uint32_t *>(os_smem + (size_t)OS_TILE * 8);
  uint32_t *whist = sval + OS_TILE;                 // [OS_WARPS][256]
  uint32_t *toff = whist + OS_WARPS * 256;          // [256] first slot of digit d inside the reordered tile
  uint32_t *gbase = toff + 256;                     // [256] global slot of the first key of digit d of this tile
  uint32_t *misc = gbase + 256;                     // [0] tile id, [1..33] scan scratch
  if (threadIdx.x == 0) misc[0] = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < OS_WARPS * 256; i += OS_THREADS) whist[i] = 0;
  __syncthreads();
  const uint32_t tile = misc[0];
  const uint32_t n = n_dev ? *n_dev : n_host;
  const uint32_t base = tile * (uint32_t)OS_TILE;
  if (base >= n) return;                             // tickets are handed out in order: no earlier tile waits on this one
  const uint32_t cnt = min((uint32_t)OS_TILE, n - base);
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31, ltmask = (1u << l) - 1u;
  uint64_t k[OS_ITEMS];
  uint32_t v[OS_ITEMS];
  uint16_t rk[OS_ITEMS];
  // warp w owns the contiguous slice [w * 32 * ITEMS, ...): element order inside the tile = (warp, round, lane)
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    uint32_t i = w * (32u * OS_ITEMS) + (uint32_t)r * 32u + l;
    bool ok = i < cnt;
    k[r] = ok ? kin[base + i] : ~0ull;
    v[r] = ok ? vin[base + i] : 0u;
  }
  uint32_t *wh = whist + w * 256;
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    uint32_t i = w * (32u * OS_ITEMS) + (uint32_t)r * 32u + l;
    bool ok = i < cnt;
    unsigned d = ok ? (unsigned)(k[r] >> shift) & 255u : 256u;
    unsigned m = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(m) - 1;
    uint32_t old = 0;
    if ((int)l == leader && ok) { old = wh[d]; wh[d] = old + __popc(m); }
    old = __shfl_sync(0xffffffffu, old, leader);
    rk[r] = (uint16_t)(old + __popc(m & ltmask));
    __syncwarp();
  }
  __syncthreads();
  // digit d = threadIdx.x: exclusive prefix over the warps, tile total
  uint32_t c = 0;
  {
    const unsigned d = threadIdx.x;
#pragma unroll
    for (int ww = 0; ww < OS_WARPS; ++ww) { uint32_t t = whist[ww * 256 + d]; whist[ww * 256 + d] = c; c += t; }
    volatile uint32_t *st = status + (size_t)tile * 256 + d;
    if (tile == 0) *st = OS_FLAG_PREFIX | c;
    else {
      *st = OS_FLAG_AGG | c;
      uint32_t excl = 0;
      for (int t = (int)tile - 1; t >= 0; --t) {
        volatile uint32_t *q = status + (size_t)t * 256 + d;
        uint32_t sv;
        do { sv = *q; } while ((sv >> 30) == 0u);
        excl += sv & OS_VAL_MASK;
        if ((sv >> 30) == 2u) break;
      }
      *st = OS_FLAG_PREFIX | (excl + c);
      gbase[d] = ghist_excl[d] + excl;
    }
    if (tile == 0) gbase[d] = ghist_excl[d];
    uint32_t tot;
    toff[d] = block_excl_scan<uint32_t>(c, misc + 1, tot);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    uint32_t i = w * (32u * OS_ITEMS) + (uint32_t)r * 32u + l;
    if (i < cnt) {
      unsigned d = (unsigned)(k[r] >> shift) & 255u;
      uint32_t pos = toff[d] + wh[d] + rk[r];
      skey[pos] = k[r]; sval[pos] = v[r];
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < cnt; i += OS_THREADS) {
    uint64_t kk = skey[i];
    unsigned d = (unsigned)(kk >> shift) & 255u;
    uint32_t dst = gbase[d] + (i - toff[d]);
    kout[dst] = kk; vout[dst] = sval[i];
  }
}

struct RadixTmp {
  uint64_t *keys_alt; uint32_t *vals_alt; uint32_t *hist; unsigned long long *scan_tmp;
};
// work space of one sort in 32-bit words: digit histograms of all passes, the tickets, the tile status words of all passes
static inline size_t radix_hist_elems(long long n)
{
  size_t tiles = (size_t)div_up(n > 0 ? n : 1, OS_TILE);
  return (size_t)OS_MAX_PASSES * 256 + 64 + (size_t)OS_MAX_PASSES * tiles * 256;
}

// sorts by bits [lo_bit, hi_bit) of the key; result is left in (keys, vals) (copies back if the number of passes is
// odd).  Stable.  n = host upper bound of the element count; n_dev (optional) = exact count on the device.
static inline int radix_sort_pairs(uint64_t *keys, uint32_t *vals, long long n, int lo_bit, int hi_bit, const RadixTmp &t, cudaStream_t st, const uint32_t *n_dev = nullptr)
{
  if (n <= 1 || hi_bit <= lo_bit) return 0;
  if (n >= (1ll << 30)) return -1;
  static std::atomic<unsigned long long> attr_done{0};   // one bit per device: the attribute belongs to the device's copy of the function
  int dev = 0; cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1)) {   // ranks-as-threads call this concurrently, one device each
    cudaFuncSetAttribute(os_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OS_SMEM);
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  int npass = (hi_bit - lo_bit + 7) / 8;
  if (npass > OS_MAX_PASSES) npass = OS_MAX_PASSES;
  int ntiles = div_up(n, OS_TILE);
  uint32_t *ghist = t.hist, *tickets = t.hist + OS_MAX_PASSES * 256, *status = t.hist + OS_MAX_PASSES * 256 + 64;
  cudaMemsetAsync(t.hist, 0, ((size_t)OS_MAX_PASSES * 256 + 64 + (size_t)npass * ntiles * 256) * 4, st);
  int hgrid = ntiles < 148 * 8 ? ntiles : 148 * 8;
  BK_LAUNCH(os_hist, hgrid, 256, 0, st, keys, n_dev, (uint32_t)n, lo_bit, npass, ghist);
  BK_LAUNCH(os_hist_scan, 1, 256, 0, st, ghist, npass);
  uint64_t *ki = keys, *ko = t.keys_alt;
  uint32_t *vi = vals, *vo = t.vals_alt;
  for (int p = 0; p < npass; ++p) {
    BK_LAUNCH(os_pass, ntiles, OS_THREADS, OS_SMEM, st, ki, vi, ko, vo, n_dev, (uint32_t)n, lo_bit + 8 * p, ghist + p * 256, status + (size_t)p * ntiles * 256, tickets + p);
    uint64_t *tk = ki; ki = ko; ko = tk;
    uint32_t *tv = vi; vi = vo; vo = tv;
  }
  if (npass & 1) {
    cudaMemcpyAsync(keys, ki, (size_t)n * 8, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(vals, vi, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
  }
  return npass;
}

}  // namespace bk
