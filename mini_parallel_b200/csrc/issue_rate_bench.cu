// issue_rate_bench.cu -- measures the per-SM issue rate of the integer / DPX
// instructions the Smith-Waterman cell update is built from (the
// roofline denominator "dpx_s16x2_thread_instr_per_clk_per_SM" is not in
// MEASURED_PEAKS.json and has to be measured on the box).
//
// Method: every thread runs NCHAIN independent dependent-chains of one
// instruction kind (so latency is hidden by ILP x TLP), a full grid of
// 1024-thread CTAs (2 per SM) keeps every SM sub-partition saturated, and the
// kernel brackets the loop with clock64() per CTA.  Rate = thread-instructions
// retired by the CTAs of one SM / cycles.  A second figure is computed from the
// CUDA-event wall time and the SM clock the driver reports under load.
//
// Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NCHAIN = 8;
constexpr int ITERS  = 2048;
constexpr int UNROLL = 8;

enum Op { OP_VIADDMNMX_RELU = 0, OP_VIADDMNMX, OP_VIMNMX3, OP_VIMNMX, OP_VIADD16X2, OP_IADD3,
          OP_LOP3, OP_IMAD, OP_VIADDMNMX_S32, OP_VIMNMX3_S32, OP_PRMT, OP_SHFL, OP_LDS,
          OP_MIX_CELL, OP_MIX_CELL_LDS, OP_MIX_ALU_FMA,
          OP_HFMA2, OP_HFMA2_RELU, OP_HMNMX2, OP_HADD2, OP_MIX_ALU_HFMA2, OP_MIX_ALU_HMNMX2, OP_MIX_CELL_TRK0, OP_MIX_CELL_TRK1, OP_COUNT };

static const char* op_name[OP_COUNT] = {
  "viaddmnmx_s16x2_relu", "viaddmnmx_s16x2", "vimnmx3_s16x2", "vimnmx_s16x2", "viadd_16x2", "iadd3",
  "lop3", "imad", "viaddmnmx_s32", "vimnmx3_s32", "prmt", "shfl", "lds",
  "mix_cell4", "mix_cell4_lds", "mix_alu_fma",
  "hfma2", "hfma2_relu", "hmnmx2", "hadd2", "mix_alu_hfma2", "mix_alu_hmnmx2", "mix_cell_trk0", "mix_cell_trk1" };
// thread-instructions issued per chain-iteration for each op kind
static const int op_instr[OP_COUNT] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 4, 5, 2, 1, 1, 1, 1, 2, 2, 5, 11 };

template <int OP>
__device__ __forceinline__ unsigned step(unsigned a, unsigned b, unsigned c, const unsigned* sm) {
  if (OP == OP_VIADDMNMX_RELU) return __viaddmax_s16x2_relu(a, b, c);
  if (OP == OP_VIADDMNMX)      return __viaddmax_s16x2(a, b, c);
  if (OP == OP_VIMNMX3)        return __vimax3_s16x2(a, b, c);
  if (OP == OP_VIMNMX)         { unsigned r; asm volatile("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_VIADD16X2)      return __vadd2(a, b);
  if (OP == OP_IADD3)          { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_LOP3)           { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_IMAD)           { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_VIADDMNMX_S32)  return (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);
  if (OP == OP_VIMNMX3_S32)    return (unsigned)__vimax3_s32((int)a, (int)b, (int)c);
  if (OP == OP_PRMT)           return __byte_perm(a, b, c);
  if (OP == OP_SHFL)           return __shfl_up_sync(0xffffffffu, a, 1);
  if (OP == OP_LDS)            return sm[(a & 63u)];
  if (OP == OP_MIX_CELL) {     // the 4-instruction cell update of the anti-diagonal kernel (no LUT fetch)
    unsigned x = a ^ b;                               // LOP3
    unsigned t = __viaddmax_s16x2(a, x, c);           // VIADDMNMX
    unsigned h = __vimax3_s16x2(t, b, c);             // VIMNMX3
    return __viaddmax_s16x2(h, c, a);                 // VIADDMNMX (tracking)
  }
  if (OP == OP_MIX_CELL_LDS) { // same plus the substitution LUT fetch from shared memory
    unsigned x = (a ^ b) & 63u;                       // LOP3
    unsigned s = sm[x];                               // LDS
    unsigned t = __viaddmax_s16x2(a, s, c);
    unsigned h = __vimax3_s16x2(t, b, c);
    return __viaddmax_s16x2(h, c, a);
  }
  if (OP == OP_MIX_ALU_FMA) {  // one ALU-pipe DPX op + one FMA-pipe IMAD: do they dual-issue?
    unsigned t = __viaddmax_s16x2(a, b, c);
    unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(t), "r"(b), "r"(c));
    return r;
  }
  // round 2: do the fp16x2 instructions issue on the FMA pipe beside the ALU-pipe DPX instructions?
  if (OP == OP_HFMA2)          { unsigned r; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_HFMA2_RELU)     { unsigned r; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  if (OP == OP_HMNMX2)         { unsigned r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_HADD2)          { unsigned r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  if (OP == OP_MIX_ALU_HFMA2) {
    unsigned t = __viaddmax_s16x2(a, b, c);
    unsigned r; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(t), "r"(b), "r"(c));
    return r;
  }
  if (OP == OP_MIX_ALU_HMNMX2) {
    unsigned t = __viaddmax_s16x2(a, b, c);
    unsigned r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(t), "r"(b));
    return r;
  }
  if (OP == OP_MIX_CELL_TRK0) { // the round-1 cell: ADD-indexed LUT fetch, two DPX for the cell, one DPX tracker
    unsigned x; asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(x) : "r"(a & 0xfcu), "r"(b & 0x3u));
    unsigned s = sm[(x >> 2) & 63u];
    unsigned t = __viaddmax_s16x2(a, s, c);
    unsigned h = __vimax3_s16x2(t, b, c);
    return __viaddmax_s16x2(h, c, a);
  }
  if (OP == OP_MIX_CELL_TRK1) { // two cells with the round-2 tracker: x = h + e as a plain add, one VIMNMX3 per two steps
    unsigned x0; asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(x0) : "r"(a & 0xfcu), "r"(b & 0x3u));
    unsigned s0 = sm[(x0 >> 2) & 63u];
    unsigned t0 = __viaddmax_s16x2(a, s0, c);
    unsigned h0 = __vimax3_s16x2(t0, b, c);
    unsigned y0; asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(y0) : "r"(h0), "r"(c));
    unsigned x1; asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(x1) : "r"(h0 & 0xfcu), "r"(b & 0x3u));
    unsigned s1 = sm[(x1 >> 2) & 63u];
    unsigned t1 = __viaddmax_s16x2(h0, s1, c);
    unsigned h1 = __vimax3_s16x2(t1, b, c);
    unsigned y1; asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(y1) : "r"(h1), "r"(b));
    return __vimax3_s16x2(a, y0, y1);
  }
  return a;
}

template <int OP>
__global__ void __launch_bounds__(1024, 2)
rate_kernel(unsigned* out, unsigned long long* cycles, unsigned seed_b, unsigned seed_c) {
  __shared__ unsigned sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x * 0x00010001u;
  __syncthreads();
  unsigned acc[NCHAIN];
#pragma unroll
  for (int k = 0; k < NCHAIN; ++k) acc[k] = threadIdx.x * 2654435761u + k * 40503u;
  unsigned b = seed_b + (threadIdx.x & 3), c = seed_c;
  unsigned long long t0 = clock64();
  for (int it = 0; it < ITERS / UNROLL; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < NCHAIN; ++k) acc[k] = step<OP>(acc[k], b, c, sm);
    }
  }
  unsigned long long t1 = clock64();
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < NCHAIN; ++k) r ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Result { double per_clk_sm_cycles; double per_clk_sm_wall; double ms; double med_cycles; };

template <int OP>
static Result run(int sms, double clock_hz) {
  const int ctas = sms * 2;             // 2 x 1024 threads = every SM fully occupied, one wave
  unsigned* out; unsigned long long* cyc;
  CK(cudaMalloc(&out, sizeof(unsigned) * 1024 * ctas));
  CK(cudaMalloc(&cyc, sizeof(unsigned long long) * ctas));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) rate_kernel<OP><<<ctas, 1024>>>(out, cyc, 0x00020002u, 0x00010003u);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f; std::vector<unsigned long long> h(ctas);
  double med = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    rate_kernel<OP><<<ctas, 1024>>>(out, cyc, 0x00020002u, 0x00010003u);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) {
      best_ms = ms;
      CK(cudaMemcpy(h.data(), cyc, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end()); med = (double)h[ctas / 2];
    }
  }
  CK(cudaGetLastError());
  const double instr_per_thread = (double)ITERS * NCHAIN * op_instr[OP];
  Result r;
  // two CTAs share one SM and run concurrently for ~med cycles
  r.per_clk_sm_cycles = 2.0 * 1024.0 * instr_per_thread / med;
  r.per_clk_sm_wall   = (double)ctas * 1024.0 * instr_per_thread / (best_ms * 1e-3) / clock_hz / sms;
  r.ms = best_ms; r.med_cycles = med;
  CK(cudaFree(out)); CK(cudaFree(cyc));
  return r;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  const double clock_hz = khz * 1e3;
  Result r[OP_COUNT];
  r[0]  = run<0>(p.multiProcessorCount, clock_hz);   r[1]  = run<1>(p.multiProcessorCount, clock_hz);
  r[2]  = run<2>(p.multiProcessorCount, clock_hz);   r[3]  = run<3>(p.multiProcessorCount, clock_hz);
  r[4]  = run<4>(p.multiProcessorCount, clock_hz);   r[5]  = run<5>(p.multiProcessorCount, clock_hz);
  r[6]  = run<6>(p.multiProcessorCount, clock_hz);   r[7]  = run<7>(p.multiProcessorCount, clock_hz);
  r[8]  = run<8>(p.multiProcessorCount, clock_hz);   r[9]  = run<9>(p.multiProcessorCount, clock_hz);
  r[10] = run<10>(p.multiProcessorCount, clock_hz);  r[11] = run<11>(p.multiProcessorCount, clock_hz);
  r[12] = run<12>(p.multiProcessorCount, clock_hz);  r[13] = run<13>(p.multiProcessorCount, clock_hz);
  r[14] = run<14>(p.multiProcessorCount, clock_hz);  r[15] = run<15>(p.multiProcessorCount, clock_hz);
  r[16] = run<16>(p.multiProcessorCount, clock_hz);  r[17] = run<17>(p.multiProcessorCount, clock_hz);
  r[18] = run<18>(p.multiProcessorCount, clock_hz);  r[19] = run<19>(p.multiProcessorCount, clock_hz);
  r[20] = run<20>(p.multiProcessorCount, clock_hz);  r[21] = run<21>(p.multiProcessorCount, clock_hz);
  r[22] = run<22>(p.multiProcessorCount, clock_hz);  r[23] = run<23>(p.multiProcessorCount, clock_hz);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_rate_mhz\": %.1f, \"rates\": {", p.name, p.multiProcessorCount, clock_hz / 1e6);
  for (int i = 0; i < OP_COUNT; ++i)
    printf("%s\"%s\": {\"thread_instr_per_clk_per_sm\": %.2f, \"by_wall_at_max_clock\": %.2f, \"ms\": %.4f}",
           i ? ", " : "", op_name[i], r[i].per_clk_sm_cycles, r[i].per_clk_sm_wall, r[i].ms);
  printf("}}\n");
  return 0;
}
