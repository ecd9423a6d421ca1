// cell_loop_bench.cu -- what the ALU pipe sustains on the REAL operand pattern of the cell update.
// issue_rate_bench.cu measures DPX instructions on chains that reuse two operand registers; the kernels update
// K slots whose three operands are all distinct registers.  This runs the inner loop of sw_stream_kernel
// (K = 10 slots, A/B ping-pong, trackers) with the substitution words in registers -- no LDS, no SHFL, no IMAD --
// so the only thing that can hold the ALU pipe below 64 thread-instr/clk/SM is operand delivery.
// Variants: 0 = the three DPX ops per cell; 1 = + one IMAD.IADD per cell (FMA pipe);
//           2 = + IMAD.IADD and a conflict-free LDS per cell (the full lookup).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
constexpr int K = 10;

template <int VAR>
__global__ void __launch_bounds__(128, 4) loop_kernel(uint32_t* out, unsigned long long* cycles, int iters, uint32_t seed)
{
  __shared__ __align__(128) uint32_t lut[81 * 32];
  for (int x = threadIdx.x; x < 81 * 32; x += blockDim.x) lut[x] = ((x >> 5) % 9 == 3 ? 384u : 192u) * 0x00010001u;
  __shared__ uint16_t ring[128 * 16];
  for (int x = threadIdx.x; x < 128 * 16; x += blockDim.x) ring[x] = (uint16_t)(((x * 2654435761u) >> 20) % 25u) << 7;
  __syncthreads();
  const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + 4u * (threadIdx.x & 31);
  uint32_t A[K], B[K], cur[K], Q[K], W[K];
#pragma unroll
  for (int m = 0; m < K; ++m) {
    A[m] = seed * (m + 1); B[m] = seed * (m + 7); cur[m] = 0;
    Q[m] = VAR == 2 ? lut_lane + (((seed >> m) % 25u) << 7) : (192u + 192u * ((seed >> m) & 1)) * 0x00010001u;
    W[m] = VAR == 0 ? 0u : (((seed >> (m + 3)) % 25u) << 7);
  }
  uint32_t floor_ = 0, e = 59u * 0x00010001u, up = seed, upPrev = seed + 1;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      if (VAR != 0) W[u] = ring[((it * K + u) & 15) * 128 + threadIdx.x];      // one window word per step, like the kernels' ring
#pragma unroll
      for (int m = K - 1; m >= 0; --m) {
        uint32_t sub;
        if (VAR == 0) sub = Q[m];
        else if (VAR == 1) sub = Q[m] + W[(u - m + K) % K];
        else { const uint32_t x = Q[m] + W[(u - m + K) % K]; asm("ld.shared.u32 %0, [%1];" : "=r"(sub) : "r"(x)); }
        uint32_t d, uu, l;
        if (u & 1) { d = m ? B[m - 1] : upPrev; uu = m ? A[m - 1] : up; l = A[m]; }
        else       { d = m ? A[m - 1] : upPrev; uu = m ? B[m - 1] : up; l = B[m]; }
        const uint32_t t1 = __viaddmax_s16x2(d, sub, uu);
        const uint32_t h  = __vimax3_s16x2(t1, l, floor_);
        if (u & 1) B[m] = h; else A[m] = h;
        cur[m] = __viaddmax_s16x2(h, e, cur[m]);
      }
      upPrev = up; up = (u & 1) ? B[K - 1] : A[K - 1];
      floor_ += 0x00800080u;
      e = __vsub2(e, 0x00810081u);
    }
    if ((it & 7) == 7) { floor_ = 0; e = 59u * 0x00010001u; }
  }
  const unsigned long long t1c = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int m = 0; m < K; ++m) r ^= A[m] ^ B[m] ^ cur[m];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1c - t0;
}

template <int VAR> static void run(int sms, int ctas_per_sm, const char* name)
{
  const int ctas = sms * ctas_per_sm, iters = 4000;
  uint32_t* out; unsigned long long* cyc;
  CK(cudaMalloc(&out, 4 * 128 * ctas)); CK(cudaMalloc(&cyc, 8 * ctas));
  loop_kernel<VAR><<<ctas, 128>>>(out, cyc, iters, 12345u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  loop_kernel<VAR><<<ctas, 128>>>(out, cyc, iters, 12345u);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long* h = (unsigned long long*)malloc(8 * ctas);
  CK(cudaMemcpy(h, cyc, 8 * ctas, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < ctas; ++i) avg += (double)h[i]; avg /= ctas;
  const double cells = (double)iters * K * K;                     // cell pairs per thread
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  // by wall time at the maximum SM clock (the CTAs of a wave do not all start together, so per-CTA clock64 spans overlap only partly)
  const double pairs_per_s = cells * 128.0 * ctas / (ms * 1e-3);
  const double dpx_per_clk_sm = 3.0 * pairs_per_s / (khz * 1e3) / sms;
  printf("\"%s_ctas%d\": {\"dpx_thread_instr_per_clk_per_sm\": %.2f, \"processed_gcups\": %.1f, \"ms\": %.3f, \"avg_cta_cycles\": %.0f}", name,
         ctas_per_sm, dpx_per_clk_sm, 2.0 * pairs_per_s / 1e9, ms, avg);
  CK(cudaFree(out)); CK(cudaFree(cyc)); free(h);
}

int main()
{
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, ", p.name, p.multiProcessorCount);
  run<0>(p.multiProcessorCount, 4, "dpx_only"); printf(", ");
  run<1>(p.multiProcessorCount, 4, "dpx_imad"); printf(", ");
  run<2>(p.multiProcessorCount, 4, "dpx_imad_lds"); printf(", ");
  run<0>(p.multiProcessorCount, 2, "dpx_only"); printf(", ");
  run<2>(p.multiProcessorCount, 2, "dpx_imad_lds");
  printf("}\n");
  return 0;
}
