// bkid_inflate.cuh -- DEFLATE (RFC 1951) decoder for BGZF blocks, written for one warp per block.
//
// Role in the reference: htslib's inflate_block / bgzf_read_block (thirdparty/samtools/samtools-1.3.1/
// htslib-1.3.1/bgzf.c:388-419,545-600), i.e. zlib's inflate() called once per <= 64 KiB BGZF block.  BGZF blocks
// are independent deflate streams, so a whole BAM decompresses block-parallel: one warp owns one block.
//
// Execution model ("group-uniform decode"): a GROUP of GS lanes (GS = 8: four groups per warp, each on its own
// BGZF block) runs the bit reader and the Huffman decode in lock step on the same data (shared-memory tables and
// uniform global loads broadcast, no shuffles) and splits the only data-parallel part, the LZ77 match copy,
// between its lanes.  An overlapping match (distance < length) is periodic in its first `distance` bytes, so
// every output byte of a match can be fetched independently: out[op+k] = out[op - dist + k % dist].
// The decoder is a resumable state machine (header_step / token_steps): the kernel drives all groups of a warp
// through the same phase in the same instruction stream, so one issued instruction serves up to four blocks --
// the first version (one block per warp) was instruction-issue bound at 31 warp-instructions per output byte.
//
// The same source compiles as plain C++ (one "lane") so the decoder logic is unit-tested on the CPU against
// zlib-compressed streams (tests/test_inflate_host.py) before it ever runs on the GPU.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define BKI_FN __host__ __device__ __forceinline__
#else
#define BKI_FN inline
#endif

// lane geometry of a group of GS lanes (GS = 1 on the host)
#if defined(__CUDA_ARCH__)
#define BKI_GSIZE(GS) ((unsigned)(GS))
#define BKI_GLANE(GS) (threadIdx.x & (unsigned)((GS) - 1))
#define BKI_GSYNC(GS) __syncwarp((GS) == 32 ? 0xffffffffu : ((((GS) == 32 ? 0u : (1u << ((GS) & 31))) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)((GS) - 1))))
#else
#define BKI_GSIZE(GS) 1u
#define BKI_GLANE(GS) 0u
#define BKI_GSYNC(GS) ((void)0)
#endif

namespace bki {

constexpr int FAST_LIT_BITS = 9;
constexpr int FAST_DIST_BITS = 7;

// per-group decode tables (shared memory on the device): 2.3 KB
struct Tables {
  uint16_t lit_fast[1 << FAST_LIT_BITS];     // (symbol << 4) | code length, 0 = longer than FAST_LIT_BITS
  uint16_t dist_fast[1 << FAST_DIST_BITS];
  uint16_t lit_count[16], dist_count[16];    // canonical code: number of codes of each length
  uint16_t lit_sym[288], dist_sym[32];       // symbols ordered by (length, symbol)
  uint8_t lens[320];                         // code lengths: literal/length then distance
};

enum Err { OK = 0, ERR_BTYPE = 1, ERR_STORED = 2, ERR_CODELEN = 3, ERR_OVERSUB = 4, ERR_SYMBOL = 5, ERR_DIST = 6, ERR_OUTPUT = 7, ERR_INPUT = 8, ERR_SIZE = 9 };

struct BitReader {
  const uint8_t *in; uint32_t len, pos; uint64_t buf; int cnt; int over;
  uint32_t nextw; int have_next;          // the aligned 32-bit word after `pos`, loaded one refill ahead
};

BKI_FN void br_init(BitReader &b, const uint8_t *in, uint32_t len) { b.in = in; b.len = len; b.pos = 0; b.buf = 0; b.cnt = 0; b.over = 0; b.nextw = 0; b.have_next = 0; }
BKI_FN void br_refill(BitReader &b)
{
  // keep > 32 valid bits.  Aligned 32-bit loads where the payload allows, issued one refill AHEAD of their use so
  // that the load latency hides behind the tokens decoded in between; bytes at a misaligned start and at the tail.
  // Past the end zeros are shifted in and counted (reported as ERR_INPUT).
  while (b.cnt <= 32) {
    if (b.have_next) {
      b.buf |= (uint64_t)b.nextw << b.cnt;
      b.cnt += 32; b.pos += 4;
      b.have_next = 0;
      if (b.pos + 4 <= b.len) { b.nextw = *reinterpret_cast<const uint32_t *>(b.in + b.pos); b.have_next = 1; }
      continue;
    }
    const uint8_t *p = b.in + b.pos;
    if ((((uintptr_t)p) & 3u) == 0 && b.pos + 4 <= b.len) {
      b.nextw = *reinterpret_cast<const uint32_t *>(p); b.have_next = 1;       // prime the pipeline
      continue;
    }
    uint64_t v = 0;
    if (b.pos < b.len) v = *p; else b.over++;
    b.pos++;
    b.buf |= v << b.cnt;
    b.cnt += 8;
  }
}
BKI_FN uint32_t br_peek(const BitReader &b, int n) { return (uint32_t)(b.buf & ((1ull << n) - 1ull)); }
BKI_FN void br_drop(BitReader &b, int n) { b.buf >>= n; b.cnt -= n; }
BKI_FN uint32_t br_bits(BitReader &b, int n)
{
  if (b.cnt < n) br_refill(b);
  uint32_t v = br_peek(b, n);
  br_drop(b, n);
  return v;
}

BKI_FN uint32_t rev_bits(uint32_t v, int n)
{
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

// canonical Huffman tables from code lengths (lane 0 only; callers synchronise).  Returns 0, or ERR_OVERSUB for
// an over-subscribed set.  Incomplete sets are accepted like zlib does for a single distance code.
BKI_FN int build(const uint8_t *lens, int n, uint16_t *count, uint16_t *sym, uint16_t *fast, int fast_bits)
{
  for (int l = 0; l < 16; ++l) count[l] = 0;
  for (int s = 0; s < n; ++s) count[lens[s]]++;
  int left = 1;
  for (int l = 1; l < 16; ++l) { left <<= 1; left -= count[l]; if (left < 0) return ERR_OVERSUB; }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
  for (int s = 0; s < n; ++s) if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
  for (int i = 0; i < (1 << fast_bits); ++i) fast[i] = 0;
  // fast table: codes no longer than fast_bits, indexed by the next bits of the stream (LSB first)
  uint32_t code = 0; int idx = 0;
  for (int l = 1; l <= fast_bits; ++l) {
    for (int k = 0; k < count[l]; ++k, ++idx, ++code) {
      uint32_t r = rev_bits(code, l);
      uint16_t e = (uint16_t)((sym[idx] << 4) | l);
      for (uint32_t j = r; j < (1u << fast_bits); j += (1u << l)) fast[j] = e;
    }
    code <<= 1;
  }
  return 0;
}

// one symbol: fast table, else bit-serial canonical decode (codes longer than the fast table are rare)
BKI_FN int decode_sym(BitReader &b, const uint16_t *fast, int fast_bits, const uint16_t *count, const uint16_t *sym)
{
  if (b.cnt < 16) br_refill(b);
  uint32_t e = fast[br_peek(b, fast_bits)];
  if (e) { br_drop(b, (int)(e & 15u)); return (int)(e >> 4); }
  int code = 0, first = 0, index = 0;
  uint64_t bits = b.buf;
  for (int l = 1; l <= 15; ++l) {
    code |= (int)(bits & 1u); bits >>= 1;
    int c = count[l];
    if (code - c < first) { br_drop(b, l); return sym[index + (code - first)]; }
    index += c; first += c; first <<= 1; code <<= 1;
  }
  return -1;
}

enum Phase { PH_HEADER = 0, PH_TOKENS = 1, PH_DONE = 2 };

struct Stream {
  BitReader b;
  uint8_t *out; uint32_t op, out_len;
  int last, phase;
};

BKI_FN void stream_init(Stream &s, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
  br_init(s.b, in, in_len);
  s.out = out; s.op = 0; s.out_len = out_len; s.last = 0; s.phase = PH_HEADER;
}

// one deflate block header: a stored block is copied whole; a Huffman block gets its tables built (lane 0 of the
// group) and the stream moves to PH_TOKENS.  All lanes of the group call this with identical state.
template <int GS>
BKI_FN int header_step(Stream &s, Tables &T)
{
  const unsigned lane = BKI_GLANE(GS);
  BitReader &b = s.b;
  s.last = (int)br_bits(b, 1);
  int type = (int)br_bits(b, 2);
  if (type == 0) {
    // stored: skip to the byte boundary, LEN / NLEN, raw bytes
    br_drop(b, b.cnt & 7);
    uint32_t len = br_bits(b, 16), nlen = br_bits(b, 16);
    if ((len ^ 0xffffu) != nlen) return ERR_STORED;
    uint32_t src = b.pos - (uint32_t)(b.cnt >> 3);        // bytes still in the bit buffer are at in[pos - cnt/8 ...]
    if (src + len > b.len) return ERR_INPUT;
    if (s.op + len > s.out_len) return ERR_OUTPUT;
    for (uint32_t k = lane; k < len; k += BKI_GSIZE(GS)) s.out[s.op + k] = b.in[src + k];
    s.op += len;
    const uint8_t *in = b.in; uint32_t in_len = b.len;
    br_init(b, in, in_len); b.pos = src + len;
    s.phase = s.last ? PH_DONE : PH_HEADER;
    return OK;
  }
  if (type == 3) return ERR_BTYPE;
  int err = 0;
  BKI_GSYNC(GS);                                          // previous block's table reads are done
  if (type == 1) {
    if (lane == 0) {
      for (int k = 0; k < 144; ++k) T.lens[k] = 8;
      for (int k = 144; k < 256; ++k) T.lens[k] = 9;
      for (int k = 256; k < 280; ++k) T.lens[k] = 7;
      for (int k = 280; k < 288; ++k) T.lens[k] = 8;
      for (int k = 0; k < 30; ++k) T.lens[288 + k] = 5;
      build(T.lens, 288, T.lit_count, T.lit_sym, T.lit_fast, FAST_LIT_BITS);
      build(T.lens + 288, 30, T.dist_count, T.dist_sym, T.dist_fast, FAST_DIST_BITS);
    }
  } else {
    int nlen = (int)br_bits(b, 5) + 257, ndist = (int)br_bits(b, 5) + 1, ncode = (int)br_bits(b, 4) + 4;
    if (nlen > 286 || ndist > 30) return ERR_CODELEN;
    // code-length code: 19 symbols of <= 7 bits, decoded bit-serially from per-lane (uniform) registers
    const char *order = "\x10\x11\x12\x00\x08\x07\x09\x06\x0a\x05\x0b\x04\x0c\x03\x0d\x02\x0e\x01\x0f";
    uint64_t cl = 0;                                      // 19 x 3 bits
    for (int i = 0; i < ncode; ++i) cl |= (uint64_t)br_bits(b, 3) << (3 * (int)order[i]);
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
      if (b.cnt < 16) br_refill(b);
      int code = 0, first = 0, index = 0, sym = -1;
      uint64_t bits = b.buf;
      for (int l = 1; l <= 7; ++l) {
        code |= (int)(bits & 1u); bits >>= 1;
        int c = cl_count[l];
        if (code - c < first) {
          int q = index + (code - first);
          sym = (int)((q < 12 ? (cl_sym_lo >> (5 * q)) : (cl_sym_hi >> (5 * (q - 12)))) & 31);
          br_drop(b, l);
          break;
        }
        index += c; first += c; first <<= 1; code <<= 1;
      }
      if (sym < 0) return ERR_CODELEN;
      if (sym < 16) { if (lane == 0) T.lens[i] = (uint8_t)sym; prev = sym; ++i; }
      else {
        int rep, val = 0;
        if (sym == 16) { if (i == 0) return ERR_CODELEN; val = prev; rep = 3 + (int)br_bits(b, 2); }
        else if (sym == 17) rep = 3 + (int)br_bits(b, 3);
        else rep = 11 + (int)br_bits(b, 7);
        if (i + rep > total) return ERR_CODELEN;
        if (lane == 0) for (int k = 0; k < rep; ++k) T.lens[i + k] = (uint8_t)val;
        i += rep; prev = val;
      }
    }
    if (lane == 0) {
      // the distance lengths follow the literal/length lengths directly: move them to their own slot
      uint8_t tmp[32];
      for (int k = 0; k < ndist; ++k) tmp[k] = T.lens[nlen + k];
      for (int k = nlen; k < 288; ++k) T.lens[k] = 0;
      for (int k = 0; k < 32; ++k) T.lens[288 + k] = k < ndist ? tmp[k] : 0;
      int e1 = build(T.lens, 288, T.lit_count, T.lit_sym, T.lit_fast, FAST_LIT_BITS);
      int e2 = build(T.lens + 288, 32, T.dist_count, T.dist_sym, T.dist_fast, FAST_DIST_BITS);
      T.lens[0] = (uint8_t)(e1 | e2);                     // status for the other lanes (lens[] is scratch from here on)
    }
    BKI_GSYNC(GS);
    err = T.lens[0] ? ERR_OVERSUB : 0;
  }
  BKI_GSYNC(GS);
  if (err) return err;
  s.phase = PH_TOKENS;
  return OK;
}

// up to `max_tokens` literal / match tokens of the current Huffman block; at the end-of-block symbol the stream moves
// on to PH_HEADER or PH_DONE
template <int GS>
BKI_FN int token_steps(Stream &s, const Tables &T, int max_tokens)
{
  const unsigned lane = BKI_GLANE(GS);
  BitReader &b = s.b;
  for (int n = 0; n < max_tokens; ++n) {
    int sym = decode_sym(b, T.lit_fast, FAST_LIT_BITS, T.lit_count, T.lit_sym);
    if (sym < 0) return ERR_SYMBOL;
    if (sym < 256) {
      if (s.op >= s.out_len) return ERR_OUTPUT;
      if (lane == 0) s.out[s.op] = (uint8_t)sym;
      ++s.op;
      continue;
    }
    if (sym == 256) { s.phase = s.last ? PH_DONE : PH_HEADER; return OK; }
    if (sym > 285) return ERR_SYMBOL;
    uint32_t len;
    if (sym < 265) len = (uint32_t)sym - 254u;
    else if (sym == 285) len = 258u;
    else {
      int e = (sym - 261) >> 2;
      len = ((4u + (uint32_t)((sym - 261) & 3)) << e) + 3u + br_bits(b, e);
    }
    int ds = decode_sym(b, T.dist_fast, FAST_DIST_BITS, T.dist_count, T.dist_sym);
    if (ds < 0 || ds > 29) return ERR_DIST;
    uint32_t dist;
    if (ds < 4) dist = (uint32_t)ds + 1u;
    else {
      int e = (ds >> 1) - 1;
      dist = ((2u + (uint32_t)(ds & 1)) << e) + 1u + br_bits(b, e);
    }
    if (dist > s.op) return ERR_DIST;
    if (s.op + len > s.out_len) return ERR_OUTPUT;
    BKI_GSYNC(GS);                                        // bytes written by other lanes of the group are visible before the copy reads them
    const uint8_t *src = s.out + (s.op - dist);
    uint8_t *dst = s.out + s.op;
    if (dist >= len) { for (uint32_t k = lane; k < len; k += BKI_GSIZE(GS)) dst[k] = src[k]; }
    else { for (uint32_t k = lane; k < len; k += BKI_GSIZE(GS)) dst[k] = src[k % dist]; }
    s.op += len;
  }
  return OK;
}

BKI_FN int stream_finish(const Stream &s)
{
  if (s.b.over > 8) return ERR_INPUT;                     // consumed bits beyond the payload (refill looks <= 5 bytes ahead)
  return s.op == s.out_len ? OK : ERR_SIZE;
}

// Inflate one raw deflate stream of `in_len` bytes into exactly `out_len` bytes (one group, run to completion).
template <int GS>
BKI_FN int inflate_raw(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, Tables &T)
{
  Stream s;
  stream_init(s, in, in_len, out, out_len);
  while (s.phase != PH_DONE) {
    int rc = s.phase == PH_HEADER ? header_step<GS>(s, T) : token_steps<GS>(s, T, 1 << 30);
    if (rc) return rc;
  }
  BKI_GSYNC(GS);
  return stream_finish(s);
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
