// CPU build of the lane-per-block DEFLATE decoder (breakid_b200/csrc/bkid_inflate.cuh compiles as plain C++): lets
// the CPU test-suite check the decoder logic against zlib-compressed streams without a GPU.
#include "../../breakid_b200/csrc/bkid_inflate.cuh"
#include <vector>

// The device decoder reads its input with aligned word loads (up to 3 bytes before `in`, 16 past its end) and writes
// its output through aligned 8-byte words: the harness provides that slack and lets the caller choose both
// misalignments; the bytes around the output window are checked to be untouched (they belong to other blocks).
static int run(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, int litmax, int in_mis, int out_mis)
{
  static bki::Tab T;
  std::vector<uint8_t> ibuf((size_t)in_len + 64, 0xA5);
  uint8_t *ip = ibuf.data() + 8;
  ip += (4 - ((uintptr_t)ip & 3)) & 3;
  ip += in_mis & 3;
  if (in_len) memcpy(ip, in, in_len);
  std::vector<uint8_t> obuf((size_t)out_len + 64, 0x5A);
  uint8_t *op = obuf.data() + 16;
  op += (8 - ((uintptr_t)op & 7)) & 7;
  op += out_mis & 7;
  int rc = litmax == 1 ? bki::inflate_raw<1>(ip, in_len, op, out_len, T) : bki::inflate_raw<4>(ip, in_len, op, out_len, T);
  for (uint8_t *q = obuf.data(); q < op; ++q) if (*q != 0x5A) return 100;              // wrote below its window
  for (uint8_t *q = op + out_len; q < obuf.data() + obuf.size(); ++q) if (*q != 0x5A) return 101;   // or above it
  if (out_len) memcpy(out, op, out_len);
  return rc;
}

extern "C" int bki_host_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
  return run(in, in_len, out, out_len, 4, 0, 0);
}

// other step shapes / alignments of the same stream: `litmax` literals per step, input and output misaligned by
// in_mis (0..3) and out_mis (0..7) bytes
extern "C" int bki_host_inflate_var(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, int litmax, int in_mis, int out_mis)
{
  return run(in, in_len, out, out_len, litmax, in_mis, out_mis);
}

// CRC-32 the way the device computes it: `nslices` contiguous slices, partials advanced and XORed together
extern "C" uint32_t bki_host_crc32_sliced(const uint8_t *p, uint32_t n, int nslices)
{
  uint32_t tab[256];
  for (uint32_t i = 0; i < 256; ++i) tab[i] = bki::crc_table_entry(i);
  uint32_t per = (n + nslices - 1) / nslices, acc = 0;
  for (int l = 0; l < nslices; ++l) {
    uint32_t lo = (uint32_t)l * per < n ? (uint32_t)l * per : n, hi = lo + per < n ? lo + per : n;
    uint32_t c = bki::crc_run(tab, l == 0 ? 0xffffffffu : 0u, p + lo, hi - lo);
    acc ^= bki::crc_shift(c, n - hi);
  }
  return acc ^ 0xffffffffu;
}
