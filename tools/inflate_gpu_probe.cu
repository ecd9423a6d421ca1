// Debug probe: inflate raw-deflate members from a file on the GPU with the production kernel code and compare with the
// expected bytes.  File format: u32 n_members, then per member u32 in_len, u32 out_len, payload, expected bytes.
#include "../mini_parallel_b200/csrc/swb_inflate.cuh"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__global__ void probe(const uint8_t* in, const uint32_t* meta, uint32_t n, uint8_t* out, int* status, uint32_t* produced, int solo)
{
  __shared__ swi::Tables T[4];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t b = blockIdx.x * 4 + warp;
  if (b >= n) return;
  swi::Lanes L{solo ? 0 : (int)lane, solo ? 1 : 32};
  uint32_t p = 0;
  const int st = swi::inflate_member(in + meta[4 * b], meta[4 * b + 1], out + meta[4 * b + 2], meta[4 * b + 3], &p, T[warp], L);
  if (lane == 0) { status[b] = st; produced[b] = p; }
}
int main(int argc, char** argv)
{
  FILE* f = fopen(argv[1], "rb"); if (!f) return 2;
  const int solo = argc > 2;
  uint32_t n; fread(&n, 4, 1, f);
  std::vector<uint8_t> in, exp; std::vector<uint32_t> meta;
  for (uint32_t k = 0; k < n; ++k) {
    uint32_t a, b; fread(&a, 4, 1, f); fread(&b, 4, 1, f);
    meta.push_back((uint32_t)in.size()); meta.push_back(a); meta.push_back((uint32_t)exp.size()); meta.push_back(b);
    in.resize(in.size() + a); fread(in.data() + in.size() - a, 1, a, f);
    exp.resize(exp.size() + b); fread(exp.data() + exp.size() - b, 1, b, f);
  }
  uint8_t *din, *dout; uint32_t *dmeta, *dprod; int* dst;
  cudaMalloc(&din, in.size() + 64); cudaMalloc(&dout, exp.size() + 64); cudaMalloc(&dmeta, meta.size() * 4); cudaMalloc(&dprod, n * 4); cudaMalloc(&dst, n * 4);
  cudaMemcpy(din, in.data(), in.size(), cudaMemcpyHostToDevice); cudaMemcpy(dmeta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, exp.size() + 64);
  probe<<<(n + 3) / 4, 128>>>(din, dmeta, n, dout, dst, dprod, solo);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<uint8_t> got(exp.size()); std::vector<int> st(n); std::vector<uint32_t> pr(n);
  cudaMemcpy(got.data(), dout, exp.size(), cudaMemcpyDeviceToHost); cudaMemcpy(st.data(), dst, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(pr.data(), dprod, n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const bool same = st[k] == 0 && pr[k] == meta[4 * k + 3] && !memcmp(got.data() + meta[4 * k + 2], exp.data() + meta[4 * k + 2], meta[4 * k + 3]);
    if (!same) { if (bad < 8) printf("member %u: status %d produced %u of %u\n", k, st[k], pr[k], meta[4 * k + 3]); ++bad; }
  }
  printf("%u members, %d bad (%s)\n", n, bad, solo ? "solo" : "warp");
  return bad != 0;
}
