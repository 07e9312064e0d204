// CPU build of the warp-uniform DEFLATE decoder (breakid_b200/csrc/bkid_inflate.cuh compiles as plain C++ with one
// "lane"): lets the CPU test-suite check the decoder logic against zlib-compressed streams without a GPU.
#include "../../breakid_b200/csrc/bkid_inflate.cuh"

extern "C" int bki_host_inflate(const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len)
{
  static bki::Tables T;
  return bki::inflate_raw(in, in_len, out, out_len, T);
}
