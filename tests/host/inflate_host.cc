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
