// Host build of the device inflate core (csrc/swb_inflate.cuh) for the CPU unit test: one "lane".
// Test infrastructure only -- the product runs this code on the GPU (swb_fastq_gpu.cu).
#include "../../mini_parallel_b200/csrc/swb_inflate.cuh"
extern "C" int swi_inflate_host(const uint8_t* in, uint64_t in_len, uint8_t* out, uint32_t out_cap, uint32_t* produced)
{
  static thread_local swi::Tables T;
  swi::Lanes L{0, 1};
  return swi::inflate_member(in, in_len, out, out_cap, produced, T, L);
}
