// CPU build of the warp-uniform DEFLATE decoder (breakid_b200/csrc/bkid_inflate.cuh compiles as plain C++ with one
// "lane"): lets the CPU test-suite check the decoder logic against zlib-compressed streams without a GPU.
#include "../../breakid_b200/csrc/bkid_inflate.cuh"

extern "C" int bki_host_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
  static bki::Tables T;
  return bki::inflate_raw<1>(in, in_len, out, out_len, T);
}

// the resumable form the GPU kernel drives: a few tokens per call, state carried in the Stream
extern "C" int bki_host_inflate_stepped(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len, int tokens_per_step)
{
  static bki::Tables T;
  bki::Stream s;
  bki::stream_init(s, in, in_len, out, out_len);
  while (s.phase != bki::PH_DONE) {
    int rc = s.phase == bki::PH_HEADER ? bki::header_step<1>(s, T) : bki::token_steps<1>(s, T, tokens_per_step);
    if (rc) return rc;
  }
  return bki::stream_finish(s);
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
