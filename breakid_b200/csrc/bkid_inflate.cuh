// bkid_inflate.cuh -- DEFLATE (RFC 1951) decoder for BGZF blocks: ONE LANE PER BLOCK.
//
// Role in the reference: htslib's inflate_block / bgzf_read_block (thirdparty/samtools/samtools-1.3.1/
// htslib-1.3.1/bgzf.c:388-419,545-600), i.e. zlib's inflate() called once per <= 64 KiB BGZF block.  BGZF blocks
// are independent deflate streams, so a whole BAM decompresses block-parallel.
//
// Execution model.  Every lane of a warp owns its own BGZF block and runs an independent scalar decoder: private
// Huffman tables in shared memory (2 KB per lane), a 64-bit bit buffer fed by aligned 32-bit loads issued one word
// ahead, and an 8-byte PENDING WORD through which all output goes, so that the decoder issues one aligned 64-bit
// store per 8 output bytes instead of byte stores (with 32 lanes on 32 different blocks every memory instruction
// touches 32 different lines: the instruction count per byte, not coalescing, is what has to be small).  One
// lane_step() decodes up to LITMAX literals, or one match token, and moves up to 8 match bytes: the source of a
// match is read with two aligned 64-bit loads (or taken from the pending word when it has not been stored yet), an
// overlapping match (distance < 8) is extended in registers (the output is periodic in `distance`).
// A warp therefore advances 32 blocks per issued instruction.  The round-1 kernel ("group-uniform": 8 lanes per
// block, all lanes decoding the same bits) needed 15.5 warp-instructions per output byte and 20 ms per 64 KiB block.
//
// The same source compiles as plain C++ so the decoder logic is unit-tested on the CPU against zlib-compressed
// streams (tests/test_inflate_host.py) before it ever runs on the GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef __CUDACC__
#define BKI_FN __host__ __device__ __forceinline__
#else
#define BKI_FN inline
#endif

namespace bki {

constexpr int LIT_BITS = 8;
constexpr int DIST_BITS = 6;

// per-lane decode tables (shared memory on the device): 1392 bytes, 43.5 KB per warp -> 5 warps per SM.  The number of
// blocks in flight per SM is what this decoder's throughput scales with (it is latency-bound), hence the small root
// tables: 96.4 % of the literal/length codes of a BAM stream are <= 8 bits (measured on the bench sample).
struct Tab {
  uint16_t lit_fast[1 << LIT_BITS];     // (symbol << 4) | code length; 0 = code longer than LIT_BITS (or unused).
                                        // Doubles as the code-length scratch while a block header is read (320 of 512 bytes)
  uint16_t lit_sym[288];                // symbols ordered by (length, symbol)
  uint16_t lit_lim[16], lit_off[16];    // canonical code, left-justified to 15 bits: lim[l] = end of the codes of length l
  uint16_t dist_fast[1 << DIST_BITS];
  uint16_t dist_lim[16], dist_off[16];
  uint8_t dist_sym[32];
  uint8_t pad[16];
};
static_assert(sizeof(Tab) == 1392 && sizeof(Tab) % 16 == 0, "one lane's tables");

enum Err { OK = 0, ERR_BTYPE = 1, ERR_STORED = 2, ERR_CODELEN = 3, ERR_OVERSUB = 4, ERR_SYMBOL = 5, ERR_DIST = 6, ERR_OUTPUT = 7, ERR_INPUT = 8, ERR_SIZE = 9 };
enum Phase { PH_HEADER = 0, PH_TOKENS = 1, PH_STORED = 2, PH_DONE = 3 };

BKI_FN uint32_t ld32a(const uint8_t *p)
{
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const uint32_t *>(p);
#else
  uint32_t v; memcpy(&v, p, 4); return v;
#endif
}
BKI_FN uint64_t ld64a(const uint8_t *p)
{
#ifdef __CUDA_ARCH__
  return *reinterpret_cast<const uint64_t *>(p);
#else
  uint64_t v; memcpy(&v, p, 8); return v;
#endif
}
BKI_FN void st64a(uint8_t *p, uint64_t v)
{
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint64_t *>(p) = v;
#else
  memcpy(p, &v, 8);
#endif
}
BKI_FN uint32_t rev15(uint32_t v)
{
#ifdef __CUDA_ARCH__
  return __brev(v) >> 17;
#else
  uint32_t r = 0;
  for (int i = 0; i < 15; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
#endif
}
BKI_FN uint32_t rev_bits(uint32_t v, int n)
{
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

// One lane's stream state.  Output positions are offsets from `ob`, the 8-byte aligned address at or below the
// block's first output byte: R0 = first byte, R = next byte, Rend = one past the last.
struct Lane {
  const uint8_t *in; uint32_t in_len, ipos;      // ipos: offset of the next (4-byte aligned) input word
  uint64_t bits; uint32_t nbits, nextw, nextw2;   // bit buffer, its fill, the two words loaded ahead (each load is consumed two refills later)
  uint64_t m0, m1;                                // match source words of the next chunk, loaded one step ahead (see lane_step)
  uint8_t *ob; uint32_t R, R0, Rend;
  uint64_t pend; uint32_t head_lo;                // pending output word; first valid byte of the block's FIRST word (foreign bytes below)
  uint32_t copy_rem, copy_dist, stored_rem;
  int phase, last;
};

// The input may be read up to 24 bytes past `in_len` and 3 bytes before `in` (aligned word loads); the caller's
// buffer provides that slack (BGZF: the next block's header / the staging buffer's padding).
BKI_FN void lane_init(Lane &L, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
  uint32_t a = (uint32_t)((uintptr_t)in & 3u);
  L.in = in; L.in_len = in_len;
  L.bits = (uint64_t)(ld32a(in - a) >> (8u * a)); L.nbits = 32u - 8u * a; L.ipos = 4u - a;
  L.nextw = ld32a(in + L.ipos); L.nextw2 = ld32a(in + L.ipos + 4);
  L.m0 = 0; L.m1 = 0;
  uint32_t a0 = (uint32_t)((uintptr_t)out & 7u);
  L.ob = out - a0; L.R0 = a0; L.R = a0; L.Rend = a0 + out_len;
  L.pend = 0; L.head_lo = a0;
  L.copy_rem = 0; L.copy_dist = 0; L.stored_rem = 0;
  L.phase = PH_HEADER; L.last = 0;
}

// > 32 valid bits afterwards
BKI_FN void refill(Lane &L)
{
  if (L.nbits <= 32u) {
    L.bits |= (uint64_t)L.nextw << L.nbits;
    L.nbits += 32u; L.ipos += 4u;
    L.nextw = L.nextw2;
    L.nextw2 = ld32a(L.in + L.ipos + 4u);
  }
}
BKI_FN void drop(Lane &L, uint32_t n) { L.bits >>= n; L.nbits -= n; }
BKI_FN uint32_t take(Lane &L, uint32_t n) { uint32_t v = (uint32_t)L.bits & ((1u << n) - 1u); drop(L, n); return v; }

// canonical Huffman tables from code lengths.  Returns 0, or ERR_OVERSUB for an over-subscribed set; incomplete
// sets are accepted (zlib accepts them for a single distance code; the unused code space decodes to an error).
template <typename SYM>
BKI_FN int build_table(const uint8_t *lens, int n, uint16_t *fast, int fast_bits, uint16_t *lim, uint16_t *off, SYM *sym)
{
  uint16_t count[16], offs[16];
  for (int l = 0; l < 16; ++l) count[l] = 0;
  for (int s = 0; s < n; ++s) count[lens[s]]++;
  count[0] = 0;
  int left = 1;
  for (int l = 1; l < 16; ++l) { left <<= 1; left -= count[l]; if (left < 0) return ERR_OVERSUB; }
  offs[0] = 0; offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
  uint32_t code = 0;
  lim[0] = 0; off[0] = 0;
  for (int l = 1; l < 16; ++l) {
    code = (code + count[l - 1]) << 1;                       // first code of length l
    lim[l] = (uint16_t)((code + count[l]) << (15 - l));      // <= 32768
    off[l] = offs[l];
  }
  for (int s = 0; s < n; ++s) if (lens[s]) sym[offs[lens[s]]++] = (SYM)s;
  for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
  // fast table: codes no longer than fast_bits, indexed by the next bits of the stream (LSB first)
  code = 0;
  for (int l = 1; l <= fast_bits; ++l) {
    code = (code + count[l - 1]) << 1;
    for (int k = 0; k < count[l]; ++k) {
      uint32_t r = rev_bits(code + (uint32_t)k, l);
      uint16_t e = (uint16_t)(((uint32_t)sym[off[l] + k] << 4) | (uint32_t)l);
      for (uint32_t j = r; j < (1u << fast_bits); j += (1u << l)) fast[j] = e;
    }
  }
  return 0;
}

// codes longer than the fast table: compare the next 15 bits (in code order) against the per-length limits.  The
// limits of all candidate lengths are loaded first (independent loads), then one compare chain finds the length.
template <typename SYM, int FROM>
BKI_FN int slow_sym(uint64_t bits, const uint16_t *lim, const uint16_t *off, const SYM *sym, uint32_t &len)
{
  uint32_t v = rev15((uint32_t)bits & 0x7fffu);
  uint32_t lm[17 - FROM];                                   // lim[FROM-1 .. 15]
#pragma unroll
  for (int k = 0; k < 17 - FROM; ++k) lm[k] = lim[FROM - 1 + k];
  int l = 16; uint32_t lo = 0;
#pragma unroll
  for (int k = 16 - FROM; k >= 1; --k) if (v < lm[k]) { l = FROM - 1 + k; lo = lm[k - 1]; }
  if (l > 15) return -1;
  len = (uint32_t)l;
  return (int)sym[off[l] + ((v - lo) >> (15 - l))];
}

// One deflate block header.  A stored block switches the lane to PH_STORED (the raw bytes then flow through the
// bit buffer); a Huffman block gets its tables built and the lane moves to PH_TOKENS.
BKI_FN int lane_header(Lane &L, Tab &T)
{
  refill(L);
  L.last = (int)take(L, 1);
  uint32_t type = take(L, 2);
  if (type == 0) {
    drop(L, L.nbits & 7u);                                  // to the byte boundary (ipos * 8 is a multiple of 8)
    refill(L);
    uint32_t len = take(L, 16), nlen = take(L, 16);
    if ((len ^ 0xffffu) != nlen) return ERR_STORED;
    uint32_t at = L.ipos - (L.nbits >> 3);                  // byte offset of the first raw byte
    if (at + len > L.in_len) return ERR_INPUT;
    if (L.R + len > L.Rend) return ERR_OUTPUT;
    L.stored_rem = len;
    L.phase = len ? PH_STORED : (L.last ? PH_DONE : PH_HEADER);
    return OK;
  }
  if (type == 3) return ERR_BTYPE;
  uint8_t *lens = reinterpret_cast<uint8_t *>(T.lit_fast);    // 320 bytes of scratch: build_table reads the lengths before it fills `fast`
  int nlen, ndist;
  if (type == 1) {
    nlen = 288; ndist = 30;
    for (int k = 0; k < 144; ++k) lens[k] = 8;
    for (int k = 144; k < 256; ++k) lens[k] = 9;
    for (int k = 256; k < 280; ++k) lens[k] = 7;
    for (int k = 280; k < 288; ++k) lens[k] = 8;
    for (int k = 0; k < 30; ++k) lens[288 + k] = 5;
  } else {
    nlen = (int)take(L, 5) + 257; ndist = (int)take(L, 5) + 1;
    int ncode = (int)take(L, 4) + 4;
    if (nlen > 286 || ndist > 30) return ERR_CODELEN;
    // code-length code: 19 symbols of <= 7 bits, kept in registers and decoded bit-serially
    const char *order = "\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f";
    uint64_t cl = 0;                                      // 19 x 3 bits
    for (int i = 0; i < ncode; ++i) { refill(L); cl |= (uint64_t)take(L, 3) << (3 * (int)order[i]); }
    int cl_count[8], cl_offs[8];
    for (int l = 0; l < 8; ++l) cl_count[l] = 0;
    for (int k = 0; k < 19; ++k) cl_count[(cl >> (3 * k)) & 7]++;
    {
      int left = 1;
      for (int l = 1; l < 8; ++l) { left <<= 1; left -= cl_count[l]; if (left < 0) return ERR_OVERSUB; }
    }
    cl_offs[1] = 0;
    for (int l = 1; l < 7; ++l) cl_offs[l + 1] = cl_offs[l] + cl_count[l];
    uint64_t cl_sym_lo = 0, cl_sym_hi = 0;                // sorted symbols, 5 bits each: 12 in lo, 7 in hi
    for (int k = 0; k < 19; ++k) {
      int l = (int)((cl >> (3 * k)) & 7);
      if (!l) continue;
      int q = cl_offs[l]++;
      if (q < 12) cl_sym_lo |= (uint64_t)k << (5 * q); else cl_sym_hi |= (uint64_t)k << (5 * (q - 12));
    }
    int total = nlen + ndist, i = 0, prev = 0;
    while (i < total) {
      refill(L);
      if (L.ipos > L.in_len + 16u) return ERR_INPUT;
      int code = 0, first = 0, index = 0, sym = -1;
      uint64_t b = L.bits;
      for (int l = 1; l <= 7; ++l) {
        code |= (int)(b & 1u); b >>= 1;
        int c = cl_count[l];
        if (code - c < first) {
          int q = index + (code - first);
          sym = (int)((q < 12 ? (cl_sym_lo >> (5 * q)) : (cl_sym_hi >> (5 * (q - 12)))) & 31);
          drop(L, (uint32_t)l);
          break;
        }
        index += c; first += c; first <<= 1; code <<= 1;
      }
      if (sym < 0) return ERR_CODELEN;
      if (sym < 16) { lens[i] = (uint8_t)sym; prev = sym; ++i; }
      else {
        int rep, val = 0;
        if (sym == 16) { if (i == 0) return ERR_CODELEN; val = prev; rep = 3 + (int)take(L, 2); }
        else if (sym == 17) rep = 3 + (int)take(L, 3);
        else rep = 11 + (int)take(L, 7);
        if (i + rep > total) return ERR_CODELEN;
        for (int k = 0; k < rep; ++k) lens[i + k] = (uint8_t)val;
        i += rep; prev = val;
      }
    }
  }
  // distances first: the literal/length build ends by overwriting the scratch with its fast table
  int e2 = build_table<uint8_t>(lens + nlen, ndist, T.dist_fast, DIST_BITS, T.dist_lim, T.dist_off, T.dist_sym);
  int e1 = build_table<uint16_t>(lens, nlen, T.lit_fast, LIT_BITS, T.lit_lim, T.lit_off, T.lit_sym);
  if (e1 | e2) return ERR_OVERSUB;
  L.phase = PH_TOKENS;
  return OK;
}

BKI_FN void flush_word(Lane &L, uint32_t W)
{
  if (L.head_lo) {                                           // the block's first word: the bytes below head_lo belong to the previous block
    for (uint32_t b = L.head_lo; b < 8u; ++b) L.ob[W + b] = (uint8_t)(L.pend >> (8u * b));
    L.head_lo = 0;
  } else st64a(L.ob + W, L.pend);
}

// append the n (1..8) low bytes of v (higher bytes zero)
BKI_FN void emit(Lane &L, uint64_t v, uint32_t n)
{
  uint32_t k = L.R & 7u;
  L.pend |= v << (8u * k);
  if (k + n >= 8u) {
    flush_word(L, L.R & ~7u);
    L.pend = k ? v >> (64u - 8u * k) : 0ull;
  }
  L.R += n;
}

// match source of the next chunk (up to 8 bytes at S = R - dist, S < R): the two aligned words around S that are already in
// memory are LOADED here and used one step later, so that their latency (these are bytes written moments ago: an L2
// round trip) overlaps the decode work the warp does for its other lanes in between.  Words at or above the pending word
// are not in memory yet: lane_copy_chunk takes them from `pend`, which the lane does not touch between the two calls.
BKI_FN void lane_copy_issue(Lane &L)
{
  uint32_t S = L.R - L.copy_dist, Ws = S & ~7u, Wb = L.R & ~7u;
  if (Ws < Wb) L.m0 = ld64a(L.ob + Ws);
  if ((S & 7u) && Ws + 8u < Wb) L.m1 = ld64a(L.ob + Ws + 8u);
}
BKI_FN void lane_copy_chunk(Lane &L)
{
  uint32_t S = L.R - L.copy_dist, Ws = S & ~7u, s = S & 7u, Wb = L.R & ~7u;
  uint64_t m0 = (Ws == Wb) ? L.pend : L.m0;
  uint64_t v = m0;
  if (s) {
    uint32_t W1 = Ws + 8u;
    uint64_t m1 = (W1 == Wb) ? L.pend : (W1 < Wb ? L.m1 : 0ull);
    v = (m0 >> (8u * s)) | (m1 << (64u - 8u * s));
  }
  if (L.copy_dist < 8u) {                                // overlapping match: the output is periodic in `dist`
    uint32_t sh = 8u * L.copy_dist;
    v &= (1ull << sh) - 1ull;
    v |= v << sh;
    if (2u * sh < 64u) v |= v << (2u * sh);
    if (4u * sh < 64u) v |= v << (4u * sh);
  }
  uint32_t n = L.copy_rem < 8u ? L.copy_rem : 8u;
  if (n < 8u) v &= (1ull << (8u * n)) - 1ull;
  L.copy_rem -= n;
  emit(L, v, n);
  if (L.copy_rem) lane_copy_issue(L);
}

// One step of a lane in PH_TOKENS / PH_STORED: first up to 8 bytes of a match in progress (its source words were loaded a
// step ago), then -- unless the match goes on -- one token: up to LITMAX literals, or a match whose first source words
// are requested right away.  The match branch comes BEFORE the literal branch so that, in a warp, the literal work of
// the other lanes runs while those loads are in flight.
template <int LITMAX>
BKI_FN int lane_step(Lane &L, const Tab &T)
{
  if (L.phase == PH_STORED) {
    refill(L);
    uint32_t n = L.stored_rem < 4u ? L.stored_rem : 4u;
    uint64_t v = L.bits & ((1ull << (8u * n)) - 1ull);
    drop(L, 8u * n);
    L.stored_rem -= n;
    if (L.stored_rem == 0) L.phase = L.last ? PH_DONE : PH_HEADER;
    emit(L, v, n);
    return OK;
  }
  if (L.copy_rem) lane_copy_chunk(L);
  if (L.copy_rem) return OK;
  refill(L);
  if (L.ipos > L.in_len + 16u) return ERR_INPUT;           // a corrupt stream cannot run away from its payload
  uint32_t e = T.lit_fast[(uint32_t)L.bits & ((1u << LIT_BITS) - 1u)], len;
  int sym;
  if (e) { len = e & 15u; sym = (int)(e >> 4); }
  else { sym = slow_sym<uint16_t, LIT_BITS + 1>(L.bits, T.lit_lim, T.lit_off, T.lit_sym, len); if (sym < 0) return ERR_SYMBOL; }
  drop(L, len);
  if (sym > 256) {
    if (sym > 285) return ERR_SYMBOL;
    uint32_t mlen;
    if (sym < 265) mlen = (uint32_t)sym - 254u;
    else if (sym == 285) mlen = 258u;
    else {
      uint32_t eb = (uint32_t)(sym - 261) >> 2;
      mlen = ((4u + ((uint32_t)(sym - 261) & 3u)) << eb) + 3u + take(L, eb);
    }
    refill(L);
    uint32_t ed = T.dist_fast[(uint32_t)L.bits & ((1u << DIST_BITS) - 1u)], dl;
    int ds;
    if (ed) { dl = ed & 15u; ds = (int)(ed >> 4); }
    else { ds = slow_sym<uint8_t, DIST_BITS + 1>(L.bits, T.dist_lim, T.dist_off, T.dist_sym, dl); if (ds < 0) return ERR_DIST; }
    drop(L, dl);
    if (ds > 29) return ERR_DIST;
    uint32_t dist;
    if (ds < 4) dist = (uint32_t)ds + 1u;
    else {
      uint32_t eb = ((uint32_t)ds >> 1) - 1u;
      dist = ((2u + ((uint32_t)ds & 1u)) << eb) + 1u + take(L, eb);
    }
    if (dist > L.R - L.R0) return ERR_DIST;
    if (L.R + mlen > L.Rend) return ERR_OUTPUT;
    L.copy_rem = mlen; L.copy_dist = dist;
    lane_copy_issue(L);
    return OK;
  }
  if (sym == 256) { L.phase = L.last ? PH_DONE : PH_HEADER; return OK; }
  uint64_t v = (uint64_t)sym; uint32_t n = 1;
#pragma unroll
  for (int j = 1; j < LITMAX; ++j) {
    refill(L);
    uint32_t e2 = T.lit_fast[(uint32_t)L.bits & ((1u << LIT_BITS) - 1u)];
    if (e2 == 0 || e2 >= (256u << 4)) break;
    drop(L, e2 & 15u);
    v |= (uint64_t)(e2 >> 4) << (8u * n);
    ++n;
  }
  if (L.R + n > L.Rend) return ERR_OUTPUT;
  emit(L, v, n);
  return OK;
}

// end of the stream: store the bytes still pending, check the sizes
BKI_FN int lane_finish(Lane &L)
{
  uint32_t k = L.R & 7u;
  if (k) {
    uint32_t W = L.R & ~7u;
    for (uint32_t b = L.head_lo; b < k; ++b) L.ob[W + b] = (uint8_t)(L.pend >> (8u * b));
  }
  if ((uint64_t)L.ipos * 8u - L.nbits > (uint64_t)L.in_len * 8u) return ERR_INPUT;   // consumed bits beyond the payload
  return L.R == L.Rend ? OK : ERR_SIZE;
}

// Inflate one raw deflate stream of `in_len` bytes into exactly `out_len` bytes (one lane, run to completion).
template <int LITMAX>
BKI_FN int inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, Tab &T)
{
  Lane L;
  lane_init(L, in, in_len, out, out_len);
  while (L.phase != PH_DONE) {
    int rc = L.phase == PH_HEADER ? lane_header(L, T) : lane_step<LITMAX>(L, T);
    if (rc) return rc;
  }
  return lane_finish(L);
}

// ---- CRC-32 (gzip / zlib polynomial, reflected) ------------------------------------------------------------------
// htslib checks every BGZF block against the CRC32 stored after the deflate payload (bgzf.c:338,404-416).  One warp
// checks one block: every lane runs the byte-table recurrence over its own contiguous slice, advances its partial
// over the bytes that follow (multiplication by x^(8n) mod P, as zlib's crc32_combine does) and the warp XORs the
// partials together -- the CRC state is linear in (state, data), so
//   crc(S0 || S1 || ...) = final_xor ^ XOR_i shift(c_i, bytes after S_i),  c_0 started from ~0, c_i>0 from 0.
constexpr uint32_t CRC_POLY = 0xedb88320u;

BKI_FN uint32_t crc_table_entry(uint32_t i)
{
  uint32_t c = i;
  for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
  return c;
}
// raw table recurrence over [p, p+n) from state `c` (no pre / post inversion)
BKI_FN uint32_t crc_run(const uint32_t *tab, uint32_t c, const uint8_t *p, uint32_t n)
{
  for (uint32_t i = 0; i < n; ++i) c = tab[(c ^ p[i]) & 0xffu] ^ (c >> 8);
  return c;
}
// a(x) * b(x) mod P in the reflected representation (zlib crc32.c multmodp)
BKI_FN uint32_t crc_mul(uint32_t a, uint32_t b)
{
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
  }
  return p;
}
// state after `nbytes` more zero bytes: c * x^(8 nbytes) mod P by square-and-multiply
BKI_FN uint32_t crc_shift(uint32_t c, uint32_t nbytes)
{
  if (nbytes == 0 || c == 0) return c;
  uint32_t sq = crc_mul(1u << 30, 1u << 30);          // x^2
  sq = crc_mul(sq, sq); sq = crc_mul(sq, sq);          // x^8
  uint32_t p = 1u << 31;                                // x^0
  for (uint32_t n = nbytes; n; n >>= 1) {
    if (n & 1u) p = crc_mul(sq, p);
    sq = crc_mul(sq, sq);
  }
  return crc_mul(p, c);
}

}  // namespace bki
