// swb_kernels.cu -- hand-written sm_100a kernels of the Smith-Waterman scoring engine.
//
// Scoring function (the reference's, smith_waterman/src/smith_waterman.cl):
//   s(i,j) = seq1[i]==seq2[j] ? +2 : -1          (:5-6, :114, raw byte equality)
//   H[i][j] = max(0, H[i-1][j-1]+s, H[i][j-1]-2, H[i-1][j]-2)   (:7, :116-125, zero borders)
// Output per pair: max H and the first cell (row-major) that reaches it.
//
// Kernels
//   pack2bit_kernel       ASCII -> 2 bits/base, 128-bit streaming loads, HBM-bound (1.25 B/base algorithmic)
//   classify_kernel       routes each pair: empty / short / long / bytes / generic, device-side work lists
//   chunk_prepare_kernel  offset rebase and window ends of one chunk of a host batch
//   sw_stream_kernel      DEFAULT short-read path (reads <= 160 bp): int16x2 DPX, two pairs per word, runs of pair
//                         couples stream through a lane group (no per-pair fill/drain), window ring in shared memory,
//                         conflict-free ADD-indexed substitution table
//   sw_short_kernel       the earlier one-couple-per-group int16x2 kernel (variants 0-3), kept as a cross-check
//   sw_long_kernel        32-bit banded wavefront, one warp per pair, bands stream through the warp; table variant for
//                         ACGT-only pairs, raw-byte variant (arithmetic substitution term) for everything else
//   sw_generic_kernel     one warp per pair, 32-bit, explicit tracking: pairs beyond 2^20 rows/columns, last-row maximum
//   ref_compat_kernel     the reference's LIVE kernel semantics (smith_waterman.cl:11-71)
//   synth_*               counter-RNG synthetic reads/windows
// (FASTQ.gz ingest kernels: swb_fastq_kernels.cu; start cell + CIGAR behind a score: swb_traceback.cu.)  -DSWB_ABLATE=n builds timing experiments of the stream kernel whose
// results are wrong on purpose; the product is always built with SWB_ABLATE=0.
#include "swb_kernels.cuh"
#ifndef SWB_ABLATE
#define SWB_ABLATE 0
#endif
#include <cstdio>
#include <cstdlib>

namespace swb {

// =====================================================================================
// 2-bit packing (pack4 in swb_kernels.cuh).  A word is flagged when any of its bytes is not exactly A/C/G/T.
// =====================================================================================
// One warp packs UNROLL groups of 32 words (512 bases each) per iteration: every lane issues UNROLL independent
// 128-bit loads before it touches the first result, which is what keeps enough bytes in flight for HBM3e.
constexpr int kPackUnroll = 4;

__device__ __forceinline__ uint4 pack_load(const uint8_t* __restrict__ bytes, uint64_t n, uint64_t w)
{
  const uint64_t base = w << 4;
  if (base + 16 <= n) return __ldcs(reinterpret_cast<const uint4*>(bytes + base));    // streamed once: 128-bit, evict-first
  uint32_t t[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};                 // ragged tail: pad with 'A'
  for (uint64_t k = base; k < n; ++k) {
    const uint32_t sh = (uint32_t)((k - base) & 3) * 8;
    t[(k - base) >> 2] = (t[(k - base) >> 2] & ~(0xFFu << sh)) | ((uint32_t)bytes[k] << sh);
  }
  return make_uint4(t[0], t[1], t[2], t[3]);
}

__global__ void __launch_bounds__(256)
pack2bit_kernel(const uint8_t* __restrict__ bytes, uint64_t n, uint32_t* __restrict__ words,
                uint32_t* __restrict__ bitmap)
{
  const uint64_t n_words = (n + 15) >> 4;
  const uint64_t n_groups = (n_words + 31) >> 5;                       // 32 words = one bitmap word
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t g0 = warp * kPackUnroll; g0 < n_groups; g0 += n_warps * kPackUnroll) {
    uint4 v[kPackUnroll];
#pragma unroll
    for (int u = 0; u < kPackUnroll; ++u) {
      const uint64_t w = ((g0 + u) << 5) + lane;
      v[u] = w < n_words ? pack_load(bytes, n, w) : make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
    }
#pragma unroll
    for (int u = 0; u < kPackUnroll; ++u) {
      const uint64_t w = ((g0 + u) << 5) + lane;
      uint32_t bad = 0;
      const uint32_t word = pack4(v[u].x, bad) | (pack4(v[u].y, bad) << 8) | (pack4(v[u].z, bad) << 16) | (pack4(v[u].w, bad) << 24);
      if (w < n_words) words[w] = word;
      const uint32_t ballot = __ballot_sync(0xffffffffu, bad != 0);
      if (lane == 0 && g0 + u < n_groups) bitmap[g0 + u] = ballot;
    }
  }
}

int launch_pack2bit(const uint8_t* bytes, uint64_t n, uint32_t* words, uint32_t* bitmap, cudaStream_t st)
{
  if (n == 0) return 0;
  const uint64_t n_words = (n + 15) >> 4;
  uint64_t blocks = (n_words + 256 * kPackUnroll - 1) / (256 * kPackUnroll);
  if (blocks > 148 * 16) blocks = 148 * 16;          // grid-stride: a multiple of the SM count
  pack2bit_kernel<<<(unsigned)blocks, 256, 0, st>>>(bytes, n, words, bitmap);
  return 1;
}

__device__ __forceinline__ uint32_t code_at(const uint32_t* __restrict__ pk, uint64_t pos)
{
  return (__ldg(pk + (pos >> 4)) >> (2u * (uint32_t)(pos & 15))) & 3u;
}

// any flagged 16-base word overlapping [lo, hi) ?
__device__ __forceinline__ bool range_flagged(const uint32_t* __restrict__ bitmap, uint64_t lo, uint64_t hi)
{
  if (hi <= lo) return false;
  const uint64_t w0 = lo >> 4, w1 = (hi - 1) >> 4;        // inclusive word range
  for (uint64_t bw = w0 >> 5; bw <= (w1 >> 5); ++bw) {
    uint32_t bits = __ldg(bitmap + bw);
    const uint64_t first = bw << 5;
    if (first < w0) bits &= ~0u << (uint32_t)(w0 - first);
    if (first + 31 > w1) bits &= ~0u >> (uint32_t)(first + 31 - w1);
    if (bits) return true;
  }
  return false;
}

__global__ void __launch_bounds__(256)
classify_kernel(BatchView b)
{
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t cls = 0xFF, m = 0;
  if (k < b.n_pairs) {
    const uint64_t q0 = b.q_beg[k], q1 = b.q_end[k], r0 = b.r_beg[k], r1 = b.r_end[k];
    const uint64_t n = q1 - q0; const uint64_t mm = r1 - r0;
    if (n == 0 || mm == 0) {                               // aligner.rs:413-416: empty input scores 0
      cls = CLASS_EMPTY;
      b.out[k] = swb_result{0, -1, -1};
    } else if (mm > b.max_window) {                        // longer than the bound the scratch rows were sized for (a wrong
      cls = CLASS_EMPTY;                                   //  max_r_len of swb_score_batch_device): not scored, reported
      b.out[k] = swb_result{INT32_MIN, -1, -1};
      atomicAdd(&b.counters->n_overflow, 1u);
    } else if (n <= kLongMaxLen && mm <= kLongMaxLen) {
      const bool acgt = !range_flagged(b.q_bad, q0, q1) && !range_flagged(b.r_bad, r0, r1);
      if (acgt && n <= b.short_max_read && mm <= kShortMaxWindow) { cls = CLASS_SHORT; m = (uint32_t)mm; }
      else if (acgt && n <= kMidMaxRead && mm <= kShortMaxWindow && b.mid_desc && !b.force_bytes) { cls = CLASS_MID; m = (uint32_t)mm; }
      else cls = (acgt && !b.force_bytes) ? CLASS_LONG : CLASS_BYTES;
    } else {
      cls = CLASS_GENERIC;
    }
  }
  // warp-aggregated append to the two work lists
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t ms = __ballot_sync(0xffffffffu, cls == CLASS_SHORT);
  const uint32_t mg = __ballot_sync(0xffffffffu, cls == CLASS_GENERIC);
  const uint32_t ml = __ballot_sync(0xffffffffu, cls == CLASS_LONG);
  const uint32_t mb = __ballot_sync(0xffffffffu, cls == CLASS_BYTES);
  const uint32_t mm_ = __ballot_sync(0xffffffffu, cls == CLASS_MID);
  uint32_t base_s = 0, base_g = 0, base_l = 0, base_b = 0, base_m = 0;
  if (lane == 0) {
    if (ms) base_s = atomicAdd(&b.counters->n_short, __popc(ms));
    if (mg) base_g = atomicAdd(&b.counters->n_generic, __popc(mg));
    if (ml) base_l = atomicAdd(&b.counters->n_long, __popc(ml));
    if (mb) base_b = atomicAdd(&b.counters->n_bytes, __popc(mb));
    if (mm_) base_m = atomicAdd(&b.counters->n_mid, __popc(mm_));
  }
  base_s = __shfl_sync(0xffffffffu, base_s, 0);
  base_g = __shfl_sync(0xffffffffu, base_g, 0);
  base_l = __shfl_sync(0xffffffffu, base_l, 0);
  base_b = __shfl_sync(0xffffffffu, base_b, 0);
  base_m = __shfl_sync(0xffffffffu, base_m, 0);
  const uint32_t below = (1u << lane) - 1;
  if (cls == CLASS_SHORT) {
    const uint32_t slot = base_s + __popc(ms & below);
    b.short_list[slot] = (uint32_t)k;
    const uint64_t q0 = b.q_beg[k], r0 = b.r_beg[k];
    uint4* d = reinterpret_cast<uint4*>(b.short_desc + slot);
    d[0] = make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)r0, (uint32_t)(r0 >> 32));
    d[1] = make_uint4((uint32_t)(b.q_end[k] - q0), m, (uint32_t)k, 0u);
  }
  if (cls == CLASS_MID) {
    const uint32_t slot = base_m + __popc(mm_ & below);
    const uint64_t q0 = b.q_beg[k], r0 = b.r_beg[k];
    uint4* d = reinterpret_cast<uint4*>(b.mid_desc + slot);
    d[0] = make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)r0, (uint32_t)(r0 >> 32));
    d[1] = make_uint4((uint32_t)(b.q_end[k] - q0), m, (uint32_t)k, 0u);
  }
  if (cls == CLASS_GENERIC) b.generic_list[base_g + __popc(mg & below)] = (uint32_t)k;
  if (cls == CLASS_LONG)    b.long_list[base_l + __popc(ml & below)] = (uint32_t)k;
  if (cls == CLASS_BYTES)   b.bytes_list[base_b + __popc(mb & below)] = (uint32_t)k;
  uint32_t wm = cls == CLASS_SHORT ? m : 0u, wmid = cls == CLASS_MID ? m : 0u;
  uint32_t rmid = cls == CLASS_MID ? (uint32_t)(b.q_end[k] - b.q_beg[k]) : 0u;
  for (int o = 16; o; o >>= 1) {
    wm = max(wm, __shfl_xor_sync(0xffffffffu, wm, o)); wmid = max(wmid, __shfl_xor_sync(0xffffffffu, wmid, o));
    rmid = max(rmid, __shfl_xor_sync(0xffffffffu, rmid, o));
  }
  if (lane == 0 && wm) atomicMax(&b.counters->max_short_window, wm);
  if (lane == 0 && wmid) atomicMax(&b.counters->max_mid_window, wmid);
  if (lane == 0 && rmid) atomicMax(&b.counters->max_mid_read, rmid);
}

// Chunk preparation of the host path: CSR offsets arrive as absolute positions in the caller's arrays and are
// rebased to the chunk's own buffers; windows of the resident reference get their end positions.
// len_a / len_b != 0: every sequence of that side has this length, its offsets were not uploaded: they are k * len.
__global__ void __launch_bounds__(256)
chunk_prepare_kernel(uint64_t* __restrict__ off_a, uint64_t n_a, uint64_t base_a, uint64_t len_a,
                     uint64_t* __restrict__ off_b, uint64_t n_b, uint64_t base_b, uint64_t len_b,
                     uint64_t* __restrict__ win_beg, const uint32_t* __restrict__ win_len, uint32_t len_w,
                     uint64_t* __restrict__ win_end, uint64_t n_w, uint64_t win_base)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_a || k < n_b || k < n_w; k += stride) {
    if (k < n_a) off_a[k] = len_a ? k * len_a : off_a[k] - base_a;
    if (k < n_b) off_b[k] = len_b ? k * len_b : off_b[k] - base_b;
    if (k < n_w) { const uint64_t w0 = win_beg[k] - win_base; win_beg[k] = w0; win_end[k] = w0 + (len_w ? len_w : win_len[k]); }
  }
}

int launch_chunk_prepare(uint64_t* off_a, uint64_t n_a, uint64_t base_a, uint64_t len_a, uint64_t* off_b, uint64_t n_b, uint64_t base_b,
                         uint64_t len_b, uint64_t* win_beg, const uint32_t* win_len, uint32_t len_w, uint64_t* win_end, uint64_t n_w, uint64_t win_base,
                         cudaStream_t st)
{
  const uint64_t n = n_a > n_b ? (n_a > n_w ? n_a : n_w) : (n_b > n_w ? n_b : n_w);
  if (n == 0) return 0;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  chunk_prepare_kernel<<<(unsigned)blocks, 256, 0, st>>>(off_a, n_a, base_a, len_a, off_b, n_b, base_b, len_b, win_beg, win_len, len_w, win_end, n_w, win_base);
  return 1;
}

int launch_classify(const BatchView& b, cudaStream_t st)
{
  if (b.n_pairs == 0) return 0;
  classify_kernel<<<(unsigned)((b.n_pairs + 255) / 256), 256, 0, st>>>(b);
  return 1;
}

// =====================================================================================
// Inter-task int16x2 kernel (reads <= G*K rows).
//
// Work decomposition.  A group of G lanes owns two pairs at once: pair A in the high
// int16 half of every score word, pair B in the low half (identical control flow, no
// dependency between the halves).  Lane L holds K consecutive rows i = K*L+m, m<K.  At
// step t slot m computes cell (i, j = t - i): all K*G cells of a step lie on ONE
// anti-diagonal, so the K cells of a lane are independent (ILP = K) and the row above
// lane L's first row arrives from lane L-1 with one SHFL per step.
//
// Arithmetic.  Stored value  V = 64 * (H + 2*tau),  tau = step inside the current block
// of BLOCK steps.  With that bias the gap additions disappear:
//     H = max(0, D+s, U-2, L-2)      <=>     V = max(V_D + 64*(s+4), V_U, V_L, floor)
// where floor = 64*2*tau is the image of H = 0.  Two DPX instructions per cell pair:
//     t = VIADDMNMX.S16x2(V_D, sub, V_U);   V = VIMNMX3.S16x2(t, V_L, floor)
// sub = {384,192} (match/mismatch) comes from a 64-entry shared-memory table indexed by
// XOR of the 3-bit base codes of both pairs (one LOP3 + one LDS).  Every BLOCK steps all
// values are rebased by -128*BLOCK so they stay inside int16 (max 64*(2*160+2*59) = 28032).
//
// End cell.  Per row, cur = max(cur, V + e) with e = tag - floor, tag = BLOCK-1-tau: the
// low 6 bits carry the (inverted) step of the first occurrence of the row maximum inside
// the block, the high bits carry 64*H.  At block end the K row trackers are folded into one
// 32-bit key per pair  H<<21 | (255-i)<<13 | (8191-NPAD-j)  whose maximum is exactly
// (max H, then min i, then min j).
// =====================================================================================
#ifdef SWB_ALL_VARIANTS      // sw_short_kernel: test build only (an independent implementation the parity tests cross-check)
struct ShortArgs {
  const uint32_t* q_pk; const uint64_t* q_beg; const uint64_t* q_end;
  const uint32_t* r_pk; const uint64_t* r_beg; const uint64_t* r_end;
  const uint32_t* list; const uint32_t* n_list;     // device-side count of listed pairs
  swb_result* out;
  uint32_t w_pad;          // columns processed per pair (multiple of K, >= longest window)
  uint32_t wbuf_stride;    // bytes of shared memory per group
};

constexpr uint32_t CODE_QPAD = 5, CODE_WPAD = 6;

template <int G, int K, bool SPLIT>
__global__ void __launch_bounds__(128)
sw_short_kernel(ShortArgs a)
{
  static_assert(K % 2 == 0, "K must be even (in-place double buffering)");
  static_assert(32 % G == 0, "G must divide the warp");
  constexpr int NPAD  = G * K;
  constexpr int BLOCK = (62 / K) * K;                   // steps between rebases, multiple of K, <= 62
  constexpr uint32_t REBASE = (uint32_t)(128 * BLOCK) * 0x00010001u;
  constexpr int GPW = 32 / G;                           // groups per warp

  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(256) uint32_t lut[64];           // 256-aligned: table address | index == address + index
  if (threadIdx.x < 64) {
    const uint32_t xa = threadIdx.x >> 3, xb = threadIdx.x & 7;
    lut[threadIdx.x] = ((xa == 0 ? 384u : 192u) << 16) | (xb == 0 ? 384u : 192u);
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t g = lane / G, L = lane % G;
  const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));
  const uint32_t n_list = *a.n_list;
  const uint32_t n_pp = (n_list + 1) >> 1;
  const uint32_t pp = (blockIdx.x * 4 + warp) * GPW + g;
  if (pp >= n_pp) return;                               // whole group idle (shuffles are group-masked)

  const uint32_t pA = a.list[2 * pp];
  const bool hasB = (2 * pp + 1) < n_list;
  const uint32_t pB = hasB ? a.list[2 * pp + 1] : pA;
  const uint64_t qA0 = a.q_beg[pA], qB0 = a.q_beg[pB], rA0 = a.r_beg[pA], rB0 = a.r_beg[pB];
  const uint32_t nA = (uint32_t)(a.q_end[pA] - qA0), nB = (uint32_t)(a.q_end[pB] - qB0);
  const uint32_t mA = (uint32_t)(a.r_end[pA] - rA0), mB = (uint32_t)(a.r_end[pB] - rB0);

  // ---- stage the combined window stream of both pairs in shared memory ----
  uint8_t* wbuf = smem + (size_t)(warp * GPW + g) * a.wbuf_stride;
  const uint32_t n_iters = (NPAD + a.w_pad + K - 1) / K;
  const uint32_t wlen = NPAD + n_iters * K;             // indices touched: [K, NPAD + steps)
  for (uint32_t x = L; x < wlen; x += G) {
    const int32_t j = (int32_t)x - NPAD;
    const uint32_t ca = (j >= 0 && (uint32_t)j < mA) ? code_at(a.r_pk, rA0 + j) : CODE_WPAD;
    const uint32_t cb = (j >= 0 && (uint32_t)j < mB) ? code_at(a.r_pk, rB0 + j) : CODE_WPAD;
    wbuf[x] = (uint8_t)(((ca << 3) | cb) << 2);
  }
  // ---- query codes of this lane's K rows ----
  const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(lut);
  uint32_t Q[K];
#pragma unroll
  for (int m = 0; m < K; ++m) {
    const uint32_t i = K * L + m;
    const uint32_t ca = i < nA ? code_at(a.q_pk, qA0 + i) : CODE_QPAD;
    const uint32_t cb = i < nB ? code_at(a.q_pk, qB0 + i) : CODE_QPAD;
    Q[m] = lut_addr | (((ca << 3) | cb) << 2);          // XOR with a window byte yields the LDS address itself
  }
  __syncwarp(gmask);

  uint32_t A[K], B[K], W[K], cur[K];
#pragma unroll
  for (int m = 0; m < K; ++m) {
    A[m] = 0xFF00FF00u;            // image of H=0 two steps before tau=0  (-256)
    B[m] = 0xFF80FF80u;            // image of H=0 one step before tau=0   (-128)
    W[m] = ((CODE_WPAD << 3) | CODE_WPAD) << 2;
    cur[m] = 0;
  }
  uint32_t recA = 0, recB = 0;     // best key of pair A (hi half) / pair B (lo half)
  uint32_t floor_ = 0, fm1 = 0xFF80FF80u, upPrev = 0xFF00FF00u;
  uint32_t e = SPLIT ? (uint32_t)(BLOCK - 1) * 65537u : (uint32_t)(BLOCK - 1) * 0x00010001u;
  const uint8_t* wp = wbuf + (NPAD - K * L);
  int32_t blockStart = 0;
  int bit = 0;

  auto block_end = [&]() {
    // fold the K row trackers of this block into the two per-pair keys, then rebase
    const int32_t P0 = (int32_t)(((255u - K * L) << 13) + K * L) + (8192 - NPAD - BLOCK - blockStart);
#pragma unroll
    for (int m = 0; m < K; ++m) {
      const int32_t Pm = P0 - 8191 * m;
      const uint32_t hi = cur[m] >> 16, lo = cur[m] & 0xFFFFu;
      const uint32_t kh = hi * 32768u - (hi & 63u) * 32767u + (uint32_t)Pm;
      const uint32_t kl = lo * 32768u - (lo & 63u) * 32767u + (uint32_t)Pm;
      recA = max(recA, kh);
      recB = max(recB, kl);
      cur[m] = 0;
      A[m] = __vsub2(A[m], REBASE);
      B[m] = __vsub2(B[m], REBASE);
    }
    upPrev = __vsub2(upPrev, REBASE);
    floor_ = 0; fm1 = 0xFF80FF80u;
    e = SPLIT ? (uint32_t)(BLOCK - 1) * 65537u : (uint32_t)(BLOCK - 1) * 0x00010001u;
    blockStart += BLOCK;
  };

  for (uint32_t it = 0; it < n_iters; ++it) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      W[u] = wp[u];
      uint32_t up = __shfl_up_sync(gmask, (u & 1) ? A[K - 1] : B[K - 1], 1, G);
      if (L == 0) up = fm1;                     // row -1: image of H = 0 at the previous step
#pragma unroll
      for (int m = K - 1; m >= 0; --m) {
        const uint32_t x = Q[m] ^ W[(u - m + K) % K];
        uint32_t sub;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(sub) : "r"(x));
        uint32_t d, uu, l;
        if (u & 1) { d = m ? B[m - 1] : upPrev; uu = m ? A[m - 1] : up; l = A[m]; }
        else       { d = m ? A[m - 1] : upPrev; uu = m ? B[m - 1] : up; l = B[m]; }
        const uint32_t t1 = __viaddmax_s16x2(d, sub, uu);
        const uint32_t h  = __vimax3_s16x2(t1, l, floor_);
        if (u & 1) B[m] = h; else A[m] = h;
        if (SPLIT) cur[m] = __vmaxs2(cur[m], h + e);
        else       cur[m] = __viaddmax_s16x2(h, e, cur[m]);
      }
      upPrev = up;
      fm1 = floor_;
      floor_ += 0x00800080u;
      e -= SPLIT ? 129u * 65537u : 0u;
      if (!SPLIT) e = __vsub2(e, 0x00810081u);
    }
    wp += K;
    if (++bit == BLOCK / K) { bit = 0; block_end(); }
  }
  if (bit) block_end();

#pragma unroll
  for (int o = G / 2; o; o >>= 1) {
    recA = max(recA, __shfl_xor_sync(gmask, recA, o, G));
    recB = max(recB, __shfl_xor_sync(gmask, recB, o, G));
  }
  if (L == 0) {
    const uint32_t sA = recA >> 21, sB = recB >> 21;
    swb_result ra{0, -1, -1}, rb{0, -1, -1};
    if (sA) ra = swb_result{(int32_t)sA, 255 - (int32_t)((recA >> 13) & 255u), 8191 - NPAD - (int32_t)(recA & 8191u)};
    if (sB) rb = swb_result{(int32_t)sB, 255 - (int32_t)((recB >> 13) & 255u), 8191 - NPAD - (int32_t)(recB & 8191u)};
    a.out[pA] = ra;
    if (hasB) a.out[pB] = rb;
  }
}

template <int G, int K, bool SPLIT>
static int launch_short_t(const BatchView& b, uint32_t window_cap, LaunchCfg& lc, int slot, cudaStream_t st)
{
  constexpr int NPAD = G * K, GPW = 32 / G;
  ShortArgs a;
  a.q_pk = b.q_pk; a.q_beg = b.q_beg; a.q_end = b.q_end; a.r_pk = b.r_pk; a.r_beg = b.r_beg; a.r_end = b.r_end;
  a.list = b.short_list; a.n_list = &b.counters->n_short; a.out = b.out;
  a.w_pad = (window_cap + K - 1) / K * K;
  const uint32_t n_iters = (NPAD + a.w_pad + K - 1) / K;
  a.wbuf_stride = (NPAD + n_iters * K + 15) & ~15u;
  const size_t smem = (size_t)a.wbuf_stride * 4 * GPW;
  if (!lc.attr_set[slot]) {                              // function attributes are per device: once per context
    cudaFuncSetAttribute(sw_short_kernel<G, K, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    lc.attr_set[slot] = true;
  }
  // the grid covers the worst case (every pair short); surplus groups read n_short and leave
  const uint64_t n_pp = (b.n_pairs + 1) / 2;
  const uint64_t blocks = (n_pp + 4 * GPW - 1) / (4 * GPW);
  sw_short_kernel<G, K, SPLIT><<<(unsigned)blocks, 128, smem, st>>>(a);
  return 1;
}

#endif  // SWB_ALL_VARIANTS

template <int G, int K, int MINB, int FLAGS> static int launch_stream_t(const BatchView& b, LaunchCfg& lc, int slot, cudaStream_t st,
                                                                        uint32_t sel_lo = 0, uint32_t sel_hi = 0);

void launch_cfg_init(LaunchCfg& lc, int sm_count)
{
  lc = LaunchCfg();
  lc.sm_count = sm_count;
  if (const char* v = getenv("SWB_STREAM_CTAS_PER_SM")) lc.stream_ctas_per_sm = atoi(v);
  if (const char* v = getenv("SWB_STREAM_GRID")) lc.stream_grid = atol(v);
  if (const char* v = getenv("SWB_LONG_K")) lc.long_k = atoi(v);
  lc.debug = getenv("SWB_DEBUG") != nullptr;
}

// Short-read kernel variants.  9 (default): sw_stream_kernel<16 lanes x 10 rows, 4 CTAs/SM> with the two-step tracker and
// dynamic couple distribution.  The others exist only in builds with -DSWB_ALL_VARIANTS (tests/native target: the parity
// tests cross-check every one of them; the product library carries the default only): 4 = the round-1 stream kernel,
// 5 / 6 = 5 CTAs per SM / 8 lanes x 20 rows, 7 = two-step tracker only, 8 = dynamic distribution only, 10 / 11 = 8 / 9 at
// 5 CTAs per SM, 0..3 = sw_short_kernel (bit0: 0 = G8/K20, 1 = G16/K10; bit1: split tracking).
int launch_short(const BatchView& b, uint32_t window_cap, int variant, LaunchCfg& lc, cudaStream_t st)
{
  if (b.n_pairs == 0) return 0;
  (void)window_cap;
  switch (variant & 15) {
#ifdef SWB_ALL_VARIANTS
    case 0: return launch_short_t<8, 20, false>(b, window_cap, lc, 0, st);
    case 1: return launch_short_t<16, 10, false>(b, window_cap, lc, 1, st);
    case 2: return launch_short_t<8, 20, true>(b, window_cap, lc, 2, st);
    case 3: return launch_short_t<16, 10, true>(b, window_cap, lc, 3, st);
    case 4: return launch_stream_t<16, 10, 4, 0>(b, lc, 4, st);
    case 5: return launch_stream_t<16, 10, 5, 0>(b, lc, 5, st);
    case 6: return launch_stream_t<8, 20, 3, 0>(b, lc, 6, st);
    case 7: return launch_stream_t<16, 10, 4, 1>(b, lc, 7, st);
    case 8: return launch_stream_t<16, 10, 4, 2>(b, lc, 8, st);
    case 10: return launch_stream_t<16, 10, 5, 2>(b, lc, 10, st);
    case 11: return launch_stream_t<16, 10, 5, 3>(b, lc, 11, st);
#endif
    default:
      // 16 lanes x 8 rows when no read of the batch is longer than 128 bp (125 bp reads: 2 % pad rows instead of 22 %)
      if (b.short_max_read <= 128) return launch_stream_t<16, 8, 4, 3>(b, lc, 13, st);
      return launch_stream_t<16, 10, 4, 3>(b, lc, 9, st);
  }
}

// reads of 161..320 bp: one group of 32 lanes x 10 rows per warp, value scale 32 (two-step tracker, dynamic distribution)
int launch_mid(const BatchView& b, LaunchCfg& lc, cudaStream_t st)
{
  if (b.n_pairs == 0 || !b.mid_desc) return 0;
  // the list's longest read decides on the device: up to 256 bp the 32 x 8-row instantiation (250 bp reads: 2 % pad rows
  // instead of 22 %), beyond that the 32 x 10-row one; the other launch returns at once
  return launch_stream_t<32, 8, 4, 3 + 8>(b, lc, 14, st, 0, 256) + launch_stream_t<32, 10, 4, 3 + 8>(b, lc, 15, st, 256, kMidMaxRead);
}

bool short_variant_available(int variant)
{
#ifdef SWB_ALL_VARIANTS
  return variant >= 0 && variant <= 11;
#else
  return variant == 9;
#endif
}

// =====================================================================================
// Streaming inter-task kernel (sw_stream_kernel): same cell arithmetic as sw_short_kernel,
// three changes that remove its structural losses.
//
// 1. Pairs STREAM through a lane group.  The anti-diagonal wavefront of sw_short_kernel pays
//    G*K fill/drain steps per pair (160 of 660 at 150x500).  Here a group owns a run of pair
//    couples whose windows are laid end to end on one column stream with a fixed stride of
//    Wp columns per pair (Wp = window cap + K-1 rounded up to K, pad columns never match).
//    Lane L switches to the next pair when its slot 0 reaches that pair's column 0 -- K steps
//    after lane L-1 did -- so the wavefront never drains: one fill/drain per RUN of pairs.
//    At its switch a lane folds its row trackers, reloads its K query codes and resets its
//    K cells to the image of H=0; the K-1 slots that are still on pad columns compute zeros.
//    A lane of the old pair reads zeros (not junk) from the lane above once that one has
//    switched; it only does so on pad columns, whose values are never the maximum.
// 2. The window stream lives in a small per-group RING in shared memory (4*G*K entries),
//    refilled G*K columns at a time, so shared memory no longer grows with the window
//    length and the staging cost is spread over the run.
// 3. The substitution table is indexed by an ADD (IMAD on the idle FMA pipe, the ALU pipe is
//    the one that saturates) instead of a LOP3, and replicated per lane so that the LDS of a
//    warp never has a bank conflict: entry(idx) of lane l sits at lut + 128*idx + 4*l,
//    idx = 9*a + b with a = (3 - q_code_A) + w_code_A in 0..8 (match iff a == 3; pad codes are
//    4 on both sides so a pad never matches), b likewise for pair B.
//
// Work distribution: persistent grid (a multiple of the SM count), group g scores the pair
// couples g, g+NG, g+2NG, ... of the short list (NG = groups in the grid).
// =====================================================================================
// The groups of a warp read their rings at the same phase; skew the ring bases so that their
// bank sets are disjoint (lane stride is K*2 bytes: G=8,K=20 -> banks {0,10,20,30,8,18,28,6}+c,
// translates by 0,1,16,17 words are disjoint; G=16,K=10 -> 16 distinct banks, translate by 16).
__device__ __forceinline__ uint32_t ring_skew(uint32_t g) { return (g & 1u) * 4u + (g >> 1) * 64u; }
template <int G> __device__ __forceinline__ uint32_t ring_skew_t(uint32_t g) { return G == 16 ? g * 64u : ring_skew(g); }

struct StreamArgs {
  const uint32_t* q_pk; const uint32_t* r_pk;
  const ShortDesc* desc;                    // the listed pairs, in list order
  const uint32_t* n_list;                   // how many (device-side count)
  const uint32_t* max_window;               // longest window among them
  uint32_t* cursor;                         // couples handed out beyond every group's static ones (zeroed per batch)
  swb_result* out;
  // Two instantiations can be launched on one list and decide on the device which of them runs (no host round trip for a
  // count): the kernel leaves at once unless sel_lo < *sel_value <= sel_hi.  sel_value == nullptr: always runs.
  const uint32_t* sel_value; uint32_t sel_lo, sel_hi;
};

// 2-bit codes of bases [pos, pos+32) of a packed array, code k in bits 2k..2k+1
__device__ __forceinline__ uint64_t codes32(const uint32_t* __restrict__ pk, uint64_t pos)
{
  const uint32_t* w = pk + (pos >> 4);
  const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);   // arenas carry 64 B of slack
  const uint32_t sh = 2u * (uint32_t)(pos & 15);
  return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}

__device__ __forceinline__ ShortDesc load_desc(const ShortDesc* __restrict__ d)
{
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(d)), b = __ldg(reinterpret_cast<const uint4*>(d) + 1);
  ShortDesc r;
  r.q0 = (uint64_t)a.x | ((uint64_t)a.y << 32); r.r0 = (uint64_t)a.z | ((uint64_t)a.w << 32);
  r.n = b.x; r.m = b.y; r.pair = b.z; r.pad = b.w;
  return r;
}

constexpr int kLutEntries = 81;
constexpr int kLutBytes = kLutEntries * 128;

template <int G, int K, int MINB, int FLAGS>
__global__ void __launch_bounds__(128, MINB)
sw_stream_kernel(StreamArgs a)
{
  constexpr bool TRK = (FLAGS & 1) != 0;                 // two-step tracker (x = h + e as a plain add, one VIMNMX3 per two steps)
  constexpr bool DYN = (FLAGS & 2) != 0;                 // couples beyond a group's first kStaticCouples come from a device-wide cursor
  constexpr uint32_t NSTAT = 5;                          // couples per group assigned statically; the cursor is read NSTAT ahead
  static_assert(K % 2 == 0 && K <= 32, "K even, at most 32 codes per 64-bit code word");
  static_assert(32 % G == 0, "G must divide the warp");
  constexpr int NPAD  = G * K;
  // Value scale.  V = SC * (H + 2*tau) must stay inside int16: SC = 64 (6 tag bits, blocks of 60 steps) holds H <= 2*160 + the
  // block's bias; the 320-row instantiation (FLAGS & 8: reads of 161..320 bp, H <= 640) runs at SC = 32 with 5 tag bits and
  // blocks of 30 steps: 32 * (640 + 2*29) = 22 336.  Keys: H << HSHIFT | (RMAX - i) << 13 | (8191 - NPAD - j).
  constexpr int TB    = (FLAGS & 8) ? 5 : 6;
  constexpr uint32_t SC = 1u << TB;
  constexpr int BLOCK = (((1 << TB) - 2) / K) * K;
  constexpr int IPB   = BLOCK / K;                       // iterations per block
  constexpr uint32_t P2 = 0x00010001u;                   // both halves
  constexpr uint32_t REBASE = (2u * SC * BLOCK) * P2;
  constexpr int ROWBITS = NPAD > 256 ? 9 : 8;
  constexpr int HSHIFT = 13 + ROWBITS;
  constexpr uint32_t RMAX = (1u << ROWBITS) - 1u;
  constexpr uint32_t FOLDM = 1u << (HSHIFT - TB);        // key = c * FOLDM - tag * (FOLDM - 1) + position
  static_assert(NPAD <= 320 && (NPAD <= 160 || TB == 5), "int16 range: 160 rows at scale 64, 320 rows at scale 32");
  constexpr uint32_t V0A = (0u - 4u * SC) & 0xFFFFu, V0B = (0u - 2u * SC) & 0xFFFFu;   // image of H = 0 two steps / one step before tau = 0
  constexpr int GPW = 32 / G;
  constexpr int RSLOTS = 4 * G;                          // ring = RSLOTS slots of K columns
  constexpr int RING = RSLOTS * K;                       // uint16 entries
  constexpr int GSTRIDE = RING * 2 + 128;                // bytes between group rings (a multiple of 128, room for ring_skew)
  constexpr uint32_t WPADV = (9u * 4u + 4u) << 7;        // ring value of a pad column

  if (a.sel_value) { const uint32_t v = *a.sel_value; if (v <= a.sel_lo || v > a.sel_hi) return; }      // (the whole grid alike, before any barrier)
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem);
  for (uint32_t x = threadIdx.x; x < kLutEntries * 32; x += blockDim.x) {
    const uint32_t idx = x >> 5, ia = idx / 9, ib = idx % 9;
    lut[x] = ((ia == 3 ? 6u * SC : 3u * SC) << 16) | (ib == 3 ? 6u * SC : 3u * SC);
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t g = lane / G, L = lane % G;
  const uint32_t n_list = *a.n_list;
  const uint32_t n_pp = (n_list + 1) >> 1;
  const uint32_t NG = gridDim.x * 4 * GPW;
  const uint32_t gidx0 = (blockIdx.x * 4 + warp) * GPW, gidx = gidx0 + g;
  if (gidx0 >= n_pp) return;                             // no pair couple for any group of this warp
  const uint32_t warpN = (n_pp - 1 - gidx0) / NG + 1;    // static distribution: the warp's first group has the longest run
  uint32_t Wp = (*a.max_window + 2 * K - 2) / K * K;
  if (Wp < NPAD + K) Wp = NPAD + K;                      // at most two pairs in flight per group
  const uint32_t ipp = Wp / K;                           // iterations per pair

  uint16_t* ring = reinterpret_cast<uint16_t*>(smem + kLutBytes + (size_t)(warp * GPW + g) * GSTRIDE + ring_skew_t<G>(g));
  const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + 4u * lane;
  // per-thread scratch (word m of thread t at [m*128 + t]: conflict-free): the query codes prepared for the
  // lane's NEXT pair, and the row trackers a lane parks at its switch until the group folds them together
  uint32_t* qnext = reinterpret_cast<uint32_t*>(smem + kLutBytes + (size_t)4 * GPW * GSTRIDE) + threadIdx.x;
  uint32_t* csave = qnext + K * 128;
  // Which couple is the group's n-th?  Static: gidx + n*NG.  DYN: the first NSTAT like that, the rest are taken from
  // counters->stream_cursor one per couple period, NSTAT couples ahead of the one being finished (the ring refill and the
  // query prefetch never look further), so that a warp the schedulers favour scores more couples instead of leaving early
  // and every warp of the grid stays resident until the list is empty.  Ids only grow, a group's first id >= n_pp ends its run.
  uint32_t* cidt = reinterpret_cast<uint32_t*>(smem + kLutBytes + (size_t)4 * GPW * GSTRIDE + (size_t)2 * K * 128 * 4) + (warp * GPW + g) * 8;
  auto couple_of = [&](uint32_t n) -> uint32_t { return DYN ? cidt[n & 7] : gidx + n * NG; };
  if (DYN) {
    if (L < 8) cidt[L] = L < NSTAT ? gidx + L * NG : 0xFFFFFFFFu;
    __syncwarp();
  }

  // ---- producer side of the ring: columns [k*NPAD + L*K, +K) of the stream per call ----
  uint32_t stg_n = 0, stg_j = L * K;
  auto stage = [&](uint32_t slot) {
    uint64_t cA = 0, cB = 0; int32_t vA = 0, vB = 0;
    const uint32_t pp = couple_of(stg_n);
    if (pp < n_pp) {
      const ShortDesc dA = load_desc(a.desc + 2 * (uint64_t)pp);
      vA = (int32_t)dA.m - (int32_t)stg_j;
      if (vA > 0) cA = codes32(a.r_pk, dA.r0 + stg_j);
      if (2 * pp + 1 < n_list) {
        const ShortDesc dB = load_desc(a.desc + 2 * (uint64_t)pp + 1);
        vB = (int32_t)dB.m - (int32_t)stg_j;
        if (vB > 0) cB = codes32(a.r_pk, dB.r0 + stg_j);
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(ring + slot * K);
#pragma unroll
    for (int x = 0; x < K; x += 2) {
      const uint32_t a0 = x < vA ? (uint32_t)(cA >> (2 * x)) & 3u : 4u, a1 = x + 1 < vA ? (uint32_t)(cA >> (2 * x + 2)) & 3u : 4u;
      const uint32_t b0 = x < vB ? (uint32_t)(cB >> (2 * x)) & 3u : 4u, b1 = x + 1 < vB ? (uint32_t)(cB >> (2 * x + 2)) & 3u : 4u;
      dst[x >> 1] = ((9u * a0 + b0) << 7) | ((9u * a1 + b1) << 23);
    }
    stg_j += NPAD;
    if (stg_j >= Wp) { stg_j -= Wp; ++stg_n; }
  };

  // ---- query codes of this lane's K rows for the lane's pair number n ----
  uint32_t Q[K];
  auto load_query = [&](uint32_t n, bool to_regs) {
    uint64_t cA = 0, cB = 0; int32_t vA = 0, vB = 0;
    const uint32_t pp = couple_of(n);
    if (pp < n_pp) {
      const ShortDesc dA = load_desc(a.desc + 2 * (uint64_t)pp);
      vA = (int32_t)dA.n - (int32_t)(K * L);
      if (vA > 0) cA = codes32(a.q_pk, dA.q0 + K * L);
      if (2 * pp + 1 < n_list) {
        const ShortDesc dB = load_desc(a.desc + 2 * (uint64_t)pp + 1);
        vB = (int32_t)dB.n - (int32_t)(K * L);
        if (vB > 0) cB = codes32(a.q_pk, dB.q0 + K * L);
      }
    }
#pragma unroll
    for (int m = 0; m < K; ++m) {
      const uint32_t qa = m < vA ? 3u - ((uint32_t)(cA >> (2 * m)) & 3u) : 4u;
      const uint32_t qb = m < vB ? 3u - ((uint32_t)(cB >> (2 * m)) & 3u) : 4u;
      const uint32_t v = lut_lane + ((9u * qa + qb) << 7);
      if (to_regs) Q[m] = v; else qnext[m * 128] = v;
    }
  };

  // ---- prologue: pad columns [-NPAD, 0), stream chunk 0, query of pair 0 ----
  {
    uint32_t* dst = reinterpret_cast<uint32_t*>(ring + L * K);
#pragma unroll
    for (int x = 0; x < K / 2; ++x) dst[x] = WPADV | (WPADV << 16);
  }
  stage(G + L);
  load_query(0, true);
  load_query(1, false);
  __syncwarp();

  uint32_t A[K], B[K], W[K], cur[K], X[K];
#pragma unroll
  for (int m = 0; m < K; ++m) { A[m] = V0A * P2; B[m] = V0B * P2; W[m] = WPADV; cur[m] = 0; X[m] = 0; }
  uint32_t recA = 0, recB = 0, prevA = 0, prevB = 0;
  uint32_t floor_ = 0, fm1 = V0B * P2, upPrev = V0A * P2;
  // e = tag - floor per half.  TRK == 1 adds it to h as ONE 32-bit integer: both halves of h are >= floor, so
  // h_half + e_half >= tag >= 0 and the carry out of the low half cancels the borrow of a negative e_half
  // exactly when e is held as e_half * 65537 (two's complement)
  uint32_t e = (uint32_t)(BLOCK - 1) * P2;               // (BLOCK-1) * 65537 == (BLOCK-1) * 0x00010001: both forms start equal
  int32_t blockStart = 0;
  int32_t pairBase = 0;                                  // first stream column of the lane's current pair
  uint32_t fin_n = 0;
  uint32_t sw_it = ipp + L, fin_it = ipp + G - 1, stage_it = 0, stage_k = 1;
  int bit = 0;

  // fold the K row trackers into the two per-pair keys  H<<HSHIFT | (RMAX-i)<<13 | (8191-NPAD-j)
  auto fold_word = [&](uint32_t c, int32_t Pm, uint32_t& ra, uint32_t& rb) {
    const uint32_t hi = c >> 16, lo = c & 0xFFFFu;
    ra = max(ra, hi * FOLDM - (hi & (SC - 1u)) * (FOLDM - 1u) + (uint32_t)Pm);
    rb = max(rb, lo * FOLDM - (lo & (SC - 1u)) * (FOLDM - 1u) + (uint32_t)Pm);
  };
  const int32_t Pconst = (int32_t)(((RMAX - K * L) << 13) + K * L) + (8192 - NPAD - BLOCK);

  const uint32_t n_iters = DYN ? 0xFFFFFFFFu : warpN * ipp + G;
  for (uint32_t it = 0; it < n_iters; ++it) {
    // ---- events at the iteration boundary ----
    if (bit == IPB) {                                    // block end (all lanes): fold, rebase
      bit = 0;
      const int32_t P0 = Pconst - blockStart + pairBase;
#pragma unroll
      for (int m = 0; m < K; ++m) {
#if SWB_ABLATE != 3
        fold_word(cur[m], P0 - 8191 * m, recA, recB);
#else
        recA |= cur[m];
#endif
        cur[m] = 0;
        A[m] = __vsub2(A[m], REBASE); B[m] = __vsub2(B[m], REBASE);
      }
      upPrev = __vsub2(upPrev, REBASE);
      floor_ = 0; fm1 = V0B * P2;
      e = (uint32_t)(BLOCK - 1) * P2;
      blockStart += BLOCK;
    }
    if (it == stage_it) {                                // refill the ring one chunk ahead (all lanes)
      stage((stage_k * G + G + L) & (RSLOTS - 1));
      ++stage_k; stage_it += G;
      __syncwarp();
    }
    if (it == sw_it) {                                   // this lane moves on to its next pair (one lane per group:
      prevA = recA; prevB = recB; recA = 0; recB = 0;    //  divergent, so it only parks / fetches / resets)
      pairBase += (int32_t)Wp; sw_it += ipp;
      const uint32_t z2 = __vsub2(floor_, (4u * SC) * P2);
#pragma unroll
      for (int m = 0; m < K; ++m) {
        csave[m * 128] = cur[m]; cur[m] = 0;
        Q[m] = qnext[m * 128];
        A[m] = z2; B[m] = fm1;
      }
    }
    if (it == fin_it) {                                  // every lane of the group has left pair fin_n (all lanes)
      // fold the trackers each lane parked at its switch (iteration fin_it - (G-1) + L, after that
      // boundary's block end) into the keys of pair fin_n
      {
        const uint32_t it_sw = it - (G - 1) + L;
        const int32_t P0 = Pconst - (int32_t)(it_sw / IPB) * BLOCK + (int32_t)(fin_n * Wp);
#pragma unroll
        for (int m = 0; m < K; ++m) fold_word(csave[m * 128], P0 - 8191 * m, prevA, prevB);
      }
      uint32_t ka = prevA, kb = prevB;
#pragma unroll
      for (int o = G / 2; o; o >>= 1) {
        ka = max(ka, __shfl_xor_sync(0xffffffffu, ka, o, G));
        kb = max(kb, __shfl_xor_sync(0xffffffffu, kb, o, G));
      }
      const uint32_t pp = couple_of(fin_n);
      if (L == 0 && pp < n_pp) {
        const uint32_t sA = ka >> HSHIFT, sB = kb >> HSHIFT;
        swb_result ra{0, -1, -1}, rb{0, -1, -1};
        if (sA) ra = swb_result{(int32_t)sA, (int32_t)RMAX - (int32_t)((ka >> 13) & RMAX), 8191 - NPAD - (int32_t)(ka & 8191u)};
        if (sB) rb = swb_result{(int32_t)sB, (int32_t)RMAX - (int32_t)((kb >> 13) & RMAX), 8191 - NPAD - (int32_t)(kb & 8191u)};
        a.out[a.desc[2 * (uint64_t)pp].pair] = ra;
        if (2 * pp + 1 < n_list) a.out[a.desc[2 * (uint64_t)pp + 1].pair] = rb;
      }
      load_query(fin_n + 2, false);                      // every lane has consumed the codes of pair fin_n+1
      if (DYN) {
        // one cursor read per warp and couple period: the id of each group's couple fin_n + NSTAT
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, (uint32_t)GPW);
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t next_valid = couple_of(fin_n + 1) < n_pp;
        __syncwarp();
        if (L == 0) cidt[(fin_n + NSTAT) & 7] = NSTAT * NG + base + g;
        __syncwarp();
        if (!__any_sync(0xffffffffu, next_valid)) break;  // no group of this warp has another couple
      }
      ++fin_n; fin_it += ipp;
    }
    ++bit;

    // ---- K steps of the wavefront ----
    const uint16_t* wp = ring + ((it + G - L) & (RSLOTS - 1)) * K;
#pragma unroll
    for (int u = 0; u < K; ++u) {
#if SWB_ABLATE == 5
      W[u] = WPADV + u;                                 /* no ring read */
#else
      W[u] = wp[u];
#endif
#if SWB_ABLATE == 4
      uint32_t up = fm1 + ((u & 1) ? A[K - 1] : B[K - 1]);   /* no shuffle, no select */
#else
      uint32_t up = __shfl_up_sync(0xffffffffu, (u & 1) ? A[K - 1] : B[K - 1], 1, G);
      if (L == 0) up = fm1;
#endif
#pragma unroll
      for (int m = K - 1; m >= 0; --m) {
#if SWB_ABLATE == 1                                  /* timing experiments only: results are wrong */
        const uint32_t sub = Q[m] ^ W[(u - m + K) % K];
#else
        const uint32_t x = Q[m] + W[(u - m + K) % K];
        uint32_t sub;
        asm("ld.shared.u32 %0, [%1];" : "=r"(sub) : "r"(x));
#endif
        uint32_t d, uu, l;
        if (u & 1) { d = m ? B[m - 1] : upPrev; uu = m ? A[m - 1] : up; l = A[m]; }
        else       { d = m ? A[m - 1] : upPrev; uu = m ? B[m - 1] : up; l = B[m]; }
        const uint32_t t1 = __viaddmax_s16x2(d, sub, uu);
        const uint32_t h  = __vimax3_s16x2(t1, l, floor_);
        if (u & 1) B[m] = h; else A[m] = h;
#if SWB_ABLATE != 2
        if (TRK) {                                       // two steps per tracker update: x = h + e is a plain add (FMA pipe)
          const uint32_t xe = h + e;
          if (u & 1) cur[m] = __vimax3_s16x2(cur[m], X[m], xe);
          else       X[m] = xe;
        } else {
          cur[m] = __viaddmax_s16x2(h, e, cur[m]);
        }
#else
        cur[m] |= h;
#endif
      }
      upPrev = up;
#if SWB_ABLATE != 6
      fm1 = floor_;
      floor_ += (2u * SC) * P2;
      e = TRK ? e - (2u * SC + 1u) * 65537u : __vsub2(e, (2u * SC + 1u) * P2);
#endif
    }
  }
}

template <int G, int K, int MINB, int FLAGS>
static int launch_stream_t(const BatchView& b, LaunchCfg& lc, int slot, cudaStream_t st, uint32_t sel_lo, uint32_t sel_hi)
{
  constexpr bool MID = (FLAGS & 8) != 0;                 // the 320-row instantiation scores the mid list
  constexpr int GPW = 32 / G;
  constexpr int GSTRIDE = 4 * G * K * 2 + 128;
  StreamArgs a;
  a.q_pk = b.q_pk; a.r_pk = b.r_pk; a.out = b.out;
  a.desc = MID ? b.mid_desc : b.short_desc;
  a.n_list = MID ? &b.counters->n_mid : &b.counters->n_short;
  a.max_window = MID ? &b.counters->max_mid_window : &b.counters->max_short_window;
  a.cursor = MID ? &b.counters->mid_cursor : &b.counters->stream_cursor;
  a.sel_value = (MID && sel_hi) ? &b.counters->max_mid_read : nullptr; a.sel_lo = sel_lo; a.sel_hi = sel_hi;
  const size_t smem = kLutBytes + (size_t)GSTRIDE * 4 * GPW + (size_t)2 * K * 128 * 4 + (size_t)4 * GPW * 8 * 4;
  int& resident = lc.resident[slot];                     // CTAs of this kernel one SM holds (asked once per context)
  if (!resident) {
    cudaFuncSetAttribute(sw_stream_kernel<G, K, MINB, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, sw_stream_kernel<G, K, MINB, FLAGS>, 128, smem) != cudaSuccess || resident < 1)
      resident = 1;
    if (lc.stream_ctas_per_sm >= 1 && lc.stream_ctas_per_sm <= resident) resident = lc.stream_ctas_per_sm;
    if (lc.debug) fprintf(stderr, "sw_stream_kernel<%d,%d,%d,%d>: %d resident CTAs/SM, %zu B smem\n", G, K, MINB, FLAGS, resident, smem);
  }
  // persistent grid: every resident CTA slot of every SM; never more groups than pair couples in the worst case
  const uint64_t n_pp = (b.n_pairs + 1) / 2;
  uint64_t blocks = (uint64_t)lc.sm_count * resident;
  if (lc.stream_grid >= 1) blocks = (uint64_t)lc.stream_grid;
  const uint64_t need = (n_pp + 4 * GPW - 1) / (4 * GPW);
  if (blocks > need) blocks = need;
  sw_stream_kernel<G, K, MINB, FLAGS><<<(unsigned)blocks, 128, smem, st>>>(a);
  return 1;
}

// =====================================================================================
// Long-pair kernel (sw_long_kernel): intra-task anti-diagonal wavefront, one warp per pair,
// 32-bit cells, for ACGT-only pairs that the int16x2 kernel cannot take (reads > 160 rows,
// windows > 4096 columns, 10 kb x 10 kb pairs).  Same cell arithmetic as the streaming kernel
// with 32-bit DPX (VIADDMNMX / VIMNMX3), one pair per word:  V = 256*(H + 2*tau), 8 tag bits,
// blocks of up to 254 steps, sub = {1536, 768} from a 9-entry per-lane table (idx = (3-q) + w).
//
// The read is cut into BANDS of 32*K rows.  The bands of a pair -- and then the next pair the warp
// steals -- stream through the warp exactly like pairs stream through a group in
// sw_stream_kernel: lane L enters the next band K steps after lane L-1, so the wavefront of
// a 10 kb x 10 kb pair (32 bands) fills and drains once, not 32 times.  The last row of a band
// leaves lane 31 one value per step into a per-warp scratch row in global memory (unbiased);
// the ring refill of the next band loads it back next to the window codes, and lane 0 reads it as
// its "row above" (zero for the first band).  A band's column stride is at least 3 wavefronts
// (3*32*K) so the stored value is always older than the refill that needs it.
// =====================================================================================
struct LongArgs {
  const uint32_t* q_pk; const uint32_t* r_pk;
  const uint8_t* q_bytes; const uint8_t* r_bytes;
  const uint64_t* q_beg; const uint64_t* q_end; const uint64_t* r_beg; const uint64_t* r_end;
  const uint32_t* list; Counters* counters;
  swb_result* out; int32_t* scratch; uint64_t scratch_stride;
};

struct LongJob {                     // one band of one pair (uniform per warp, lives in shared memory)
  uint64_t q0, r0;                   // first base of the band's rows / of the window
  uint32_t rows, n2;                 // valid rows in this band (<= 32*K; 0 = dummy job), window length
  uint32_t row0;                     // first row of the band inside the read
  uint32_t pair;                     // pair index in the batch
  uint32_t base_it;                  // iteration at which lane 0 enters the job (stream column = base_it*K)
  uint32_t ipp;                      // iterations per job (column stride / K)
  uint32_t first, last;              // first / last band of its pair
};


// BYTES = true is the same kernel on RAW BYTES (any alphabet: 'N' == 'N', 'a' != 'A' as in smith_waterman.cl:114): the
// ring holds window bytes, Q the read bytes (pads 0x200 / 0x100 never compare equal) and the substitution term is
// arithmetic instead of a table fetch -- d = q - w, sub = 6S - 3S*min(d*d, 1): three FMA-pipe IMADs and one ALU min.
template <int K, int MINB, bool BYTES>
__global__ void __launch_bounds__(128, MINB)
sw_long_kernel(LongArgs a)
{
  static_assert(K % 2 == 0 && K <= 32, "K even, at most 32 codes per 64-bit code word");
  constexpr int G = 32;
  constexpr int NPAD  = G * K;
  constexpr int TB = 8;                                  // tag bits
  constexpr int BLOCK = (254 / K) * K;
  constexpr int IPB   = BLOCK / K;
  constexpr uint32_t SC = 1u << TB;                      // value scale
  constexpr uint32_t REBASE = 2u * SC * BLOCK;
  constexpr int RSLOTS = 4 * G;
  constexpr int RING = RSLOTS * K;
  constexpr uint32_t WPADV = BYTES ? 0x200u : (4u << 7);
  constexpr uint32_t NOEVENT = 0xFFFFFFFFu;
  constexpr uint64_t M21 = (1ull << 21) - 1;
  constexpr int JT = 16;                                 // job table entries (live: lane 31's job .. LA jobs ahead of lane 0)
  constexpr int LA = 4;                                  // jobs created ahead of lane 0 (the ring refill runs up to 3 chunks ahead)

  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem);     // 9 entries x 32 lanes
  for (uint32_t x = threadIdx.x; x < 9 * 32; x += blockDim.x) lut[x] = ((x >> 5) == 3) ? 6u * SC : 3u * SC;
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, L = lane;
  constexpr size_t WARP_BYTES = (size_t)RING * 2 + (size_t)RING * 4 + JT * sizeof(LongJob);
  uint8_t* wbase = smem + 9 * 128 + warp * WARP_BYTES;
  uint32_t* bring = reinterpret_cast<uint32_t*>(wbase);                       // boundary row values per stream column
  uint16_t* ring  = reinterpret_cast<uint16_t*>(wbase + (size_t)RING * 4);    // window codes per stream column
  LongJob*  jobs  = reinterpret_cast<LongJob*>(wbase + (size_t)RING * 6);
  uint32_t* qnext = reinterpret_cast<uint32_t*>(smem + 9 * 128 + 4 * WARP_BYTES) + threadIdx.x;
  uint32_t* csave = qnext + K * 128;
  const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + 4u * lane;
  int32_t* scratch = a.scratch + (uint64_t)(blockIdx.x * 4 + warp) * a.scratch_stride;
  const uint32_t n_list = BYTES ? a.counters->n_bytes : a.counters->n_long;
  uint32_t* const cursor = BYTES ? &a.counters->bytes_cursor : &a.counters->long_cursor;

  // ---- job creation (uniform): next band of the current pair, else steal the next pair ----
  uint32_t created = 0, next_base_it = 0;                // jobs created so far, base_it of the next one
  uint32_t cp_pair = 0, cp_n1 = 0, cp_n2 = 0, cp_row = 0; uint64_t cp_q0 = 0, cp_r0 = 0;   // pair being cut into bands
  bool cp_live = false, exhausted = false;
  uint32_t first_dummy = NOEVENT;
  auto create_job = [&]() {
    if (!cp_live && !exhausted) {
      uint32_t item = 0;
      if (lane == 0) item = atomicAdd(cursor, 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item < n_list) {
        cp_pair = a.list[item];
        cp_q0 = a.q_beg[cp_pair]; cp_r0 = a.r_beg[cp_pair];
        cp_n1 = (uint32_t)(a.q_end[cp_pair] - cp_q0); cp_n2 = (uint32_t)(a.r_end[cp_pair] - cp_r0);
        cp_row = 0; cp_live = true;
      } else {
        exhausted = true;
      }
    }
    LongJob j;
    if (cp_live) {
      j.q0 = cp_q0 + cp_row; j.r0 = cp_r0; j.n2 = cp_n2; j.row0 = cp_row; j.pair = cp_pair;
      j.rows = min(cp_n1 - cp_row, (uint32_t)NPAD);
      j.first = cp_row == 0; j.last = cp_row + NPAD >= cp_n1;
      // column stride: the window plus K-1 pad columns; a band that hands its last row to the next band (or takes one)
      // needs three wavefronts between producer and consumer; every job needs > 32 iterations so lanes 0 and 31 are
      // never more than one job apart
      uint32_t wp = (cp_n2 + 2 * K - 2) / K * K;
      const uint32_t wmin = (j.first && j.last) ? 34u * K : 3u * NPAD + K;
      if (wp < wmin) wp = wmin;
      j.ipp = wp / K;
      cp_row += NPAD;
      if (j.last) cp_live = false;
    } else {
      if (first_dummy == NOEVENT) first_dummy = created;
      j.q0 = 0; j.r0 = 0; j.n2 = 0; j.row0 = 0; j.pair = 0; j.rows = 0; j.first = 1; j.last = 1; j.ipp = 34;
    }
    j.base_it = next_base_it;
    next_base_it += j.ipp;
    if (lane == 0) jobs[created % JT] = j;
    ++created;
    __syncwarp();
  };

  // ---- ring refill: columns [k*NPAD + L*K, +K) of the stream: window codes + boundary values ----
  uint32_t stg_job = 0;                                  // job the lane's next refill piece belongs to
  auto stage = [&](uint32_t chunk) {
    const uint32_t it0 = chunk * G + L;                  // the piece starts at stream column it0*K
    while (it0 >= jobs[stg_job % JT].base_it + jobs[stg_job % JT].ipp) ++stg_job;
    const LongJob j = jobs[stg_job % JT];
    const uint32_t col0 = (it0 - j.base_it) * K;
    const int32_t v = j.rows ? (int32_t)j.n2 - (int32_t)col0 : 0;       // valid columns in this piece
    uint64_t cw = 0;
    if (!BYTES && v > 0) cw = codes32(a.r_pk, j.r0 + col0);
    const uint8_t* wb = a.r_bytes + j.r0 + col0;
    const uint32_t slot = (chunk * G + G + L) & (RSLOTS - 1);
    uint32_t* dst = reinterpret_cast<uint32_t*>(ring + slot * K);
    uint32_t* bdst = bring + slot * K;
    const bool has_above = v > 0 && !j.first;
#pragma unroll
    for (int x = 0; x < K; x += 2) {
      if (BYTES) {
        const uint32_t w0 = x < v ? (uint32_t)__ldg(wb + x) : WPADV, w1 = x + 1 < v ? (uint32_t)__ldg(wb + x + 1) : WPADV;
        dst[x >> 1] = w0 | (w1 << 16);
      } else {
        const uint32_t w0 = x < v ? (uint32_t)(cw >> (2 * x)) & 3u : 4u, w1 = x + 1 < v ? (uint32_t)(cw >> (2 * x + 2)) & 3u : 4u;
        dst[x >> 1] = (w0 << 7) | (w1 << 23);
      }
      bdst[x]     = (has_above && x < v)     ? (uint32_t)__ldcg(scratch + col0 + x)     : 0u;
      bdst[x + 1] = (has_above && x + 1 < v) ? (uint32_t)__ldcg(scratch + col0 + x + 1) : 0u;
    }
  };

  // ---- query codes of the lane's K rows in job jn, parked in shared memory until the lane switches ----
  uint32_t Q[K];
  auto load_query = [&](uint32_t jn, bool to_regs) {
    const LongJob j = jobs[jn % JT];
    const int32_t v = (int32_t)j.rows - (int32_t)(K * L);
    uint64_t cq = 0;
    if (!BYTES && v > 0) cq = codes32(a.q_pk, j.q0 + K * L);
    const uint8_t* qb = a.q_bytes + j.q0 + K * L;
#pragma unroll
    for (int m = 0; m < K; ++m) {
      uint32_t val;
      if (BYTES) {
        val = m < v ? (uint32_t)__ldg(qb + m) : 0x100u;
      } else {
        const uint32_t qc = m < v ? 3u - ((uint32_t)(cq >> (2 * m)) & 3u) : 4u;
        val = lut_lane + (qc << 7);
      }
      if (to_regs) Q[m] = val; else qnext[m * 128] = val;
    }
  };

  // ---- prologue ----
#pragma unroll 1
  for (int k = 0; k <= LA; ++k) create_job();
  if (jobs[0].rows == 0) return;                         // nothing to steal for this warp
  {
    uint32_t* dst = reinterpret_cast<uint32_t*>(ring + L * K);
#pragma unroll
    for (int x = 0; x < K / 2; ++x) dst[x] = WPADV | (WPADV << 16);
#pragma unroll
    for (int x = 0; x < K; ++x) bring[L * K + x] = 0;
  }
  stage(0);
  load_query(0, true);
  load_query(1, false);
  __syncwarp();

  uint32_t A[K], B[K], W[K], cur[K];
#pragma unroll
  for (int m = 0; m < K; ++m) { A[m] = 0u - 4u * SC; B[m] = 0u - 2u * SC; W[m] = WPADV; cur[m] = 0; }
  uint64_t rec = 0, prev = 0, pair_best = 0;
  uint32_t floor_ = 0, fm1 = 0u - 2u * SC, upPrev = 0u - 4u * SC;
  uint32_t e = (uint32_t)(BLOCK - 1);
  int32_t blockStart = 0;
  uint32_t head = 0, tail = 0, ljob = 0;                 // job of lane 0 / of lane 31 (uniform), of this lane
  uint32_t it_head = jobs[1].base_it, sw_it = jobs[1].base_it + L, fin_it = jobs[1].base_it + (G - 1);
  uint32_t stage_it = 0, stage_k = 1;
  // per-lane context of the job the lane is in: key offset and what lane 31 stores
  int32_t jobCol = 0;                                    // stream column of the job's column 0
  uint32_t rowBase = K * L;                              // absolute row of slot 0
  uint32_t store_n2 = jobs[0].last ? 0u : jobs[0].n2;    // lane 31: columns to park for the next band
  int bit = 0;

  // key = H << 42 | (M21 - i) << 21 | (M21 - NPAD - j)
  auto fold_word = [&](uint32_t c, uint32_t i, int32_t tcol, uint64_t& r) {
    const uint32_t H = c >> TB, tag = c & (SC - 1);
    const int32_t j = tcol - (int32_t)tag;               // tcol = column of this slot at tag 0
    const uint64_t key = ((uint64_t)H << 42) | ((M21 - i) << 21) | (uint64_t)(uint32_t)((int32_t)M21 - NPAD - j);
    r = max(r, key);
  };

  for (uint32_t it = 0;; ++it) {
    // ---- events at the iteration boundary ----
    if (bit == IPB) {                                    // block end (all lanes)
      bit = 0;
      const int32_t tcol = blockStart + (BLOCK - 1) - (int32_t)(K * L) - jobCol;
#pragma unroll
      for (int m = 0; m < K; ++m) {
        fold_word(cur[m], rowBase + m, tcol - m, rec);
        cur[m] = 0;
        A[m] -= REBASE; B[m] -= REBASE;
      }
      upPrev -= REBASE;
      floor_ = 0; fm1 = 0u - 2u * SC;
      e = (uint32_t)(BLOCK - 1);
      blockStart += BLOCK;
    }
    if (it == stage_it) {                                // refill the ring one chunk ahead (all lanes)
      stage(stage_k);
      ++stage_k; stage_it += G;
      __syncwarp();
    }
    if (it + 1 == it_head) load_query(head + 1, false);  // every lane has taken the codes of job `head` (all lanes)
    if (it == it_head) {                                 // lane 0 enters job head+1: look one more job ahead (all lanes)
      ++head;
      create_job();
      it_head = jobs[(head + 1) % JT].base_it;
    }
    if (it == sw_it) {                                   // this lane enters its next job (one lane: park / fetch / reset)
      ++ljob;
      const LongJob j = jobs[ljob % JT];
      prev = rec; rec = 0;
      jobCol = (int32_t)(j.base_it * K); rowBase = j.row0 + K * L;
      store_n2 = j.last ? 0u : j.n2;
      sw_it = jobs[(ljob + 1) % JT].base_it + L;         // jobs are created LA ahead of lane 0
      const uint32_t z2 = floor_ - 4u * SC;
#pragma unroll
      for (int m = 0; m < K; ++m) {
        csave[m * 128] = cur[m]; cur[m] = 0;
        Q[m] = qnext[m * 128];
        A[m] = z2; B[m] = fm1;
      }
    }
    if (it == fin_it) {                                  // every lane has left job `tail` (all lanes)
      const LongJob j = jobs[tail % JT];
      {
        const uint32_t it_sw = it - (G - 1) + L;         // the iteration at which this lane parked its trackers
        const int32_t tcol = (int32_t)(it_sw / IPB) * BLOCK + (BLOCK - 1) - (int32_t)(K * L) - (int32_t)(j.base_it * K);
#pragma unroll
        for (int m = 0; m < K; ++m) fold_word(csave[m * 128], j.row0 + K * L + m, tcol - m, prev);
      }
      uint64_t k = prev;
#pragma unroll
      for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
      pair_best = j.first ? k : max(pair_best, k);
      if (j.last && j.rows && lane == 0) {
        const uint32_t sc = (uint32_t)(pair_best >> 42);
        swb_result r{0, -1, -1};
        if (sc) r = swb_result{(int32_t)sc, (int32_t)(M21 - ((pair_best >> 21) & M21)), (int32_t)M21 - NPAD - (int32_t)(pair_best & M21)};
        a.out[j.pair] = r;
      }
      ++tail;
      if (tail == first_dummy) break;                    // every real job of this warp is written
      fin_it = jobs[(tail + 1) % JT].base_it + (G - 1);
    }
    ++bit;

    // ---- K steps of the wavefront ----
    const uint32_t slot = (it + G - L) & (RSLOTS - 1);
    const uint16_t* wp = ring + slot * K;
    const uint32_t* bp = bring + ((it + G) & (RSLOTS - 1)) * K;     // lane 0's columns
    const int32_t scol = (int32_t)(it * K) - (NPAD - 1) - jobCol;   // column of lane 31's last slot at step 0
#pragma unroll
    for (int u = 0; u < K; ++u) {
      W[u] = wp[u];
      uint32_t up = __shfl_up_sync(0xffffffffu, (u & 1) ? A[K - 1] : B[K - 1], 1);
      if (L == 0) up = bp[u] + fm1;
#pragma unroll
      for (int m = K - 1; m >= 0; --m) {
        uint32_t sub;
        if (BYTES) {
          const uint32_t dq = Q[m] - W[(u - m + K) % K];
          sub = 6u * SC - 3u * SC * min(dq * dq, 1u);
        } else {
          const uint32_t x = Q[m] + W[(u - m + K) % K];
          asm("ld.shared.u32 %0, [%1];" : "=r"(sub) : "r"(x));
        }
        uint32_t d, uu, l;
        if (u & 1) { d = m ? B[m - 1] : upPrev; uu = m ? A[m - 1] : up; l = A[m]; }
        else       { d = m ? A[m - 1] : upPrev; uu = m ? B[m - 1] : up; l = B[m]; }
        const uint32_t t1 = (uint32_t)__viaddmax_s32((int32_t)d, (int32_t)sub, (int32_t)uu);
        const uint32_t h  = (uint32_t)__vimax3_s32((int32_t)t1, (int32_t)l, (int32_t)floor_);
        if (u & 1) B[m] = h; else A[m] = h;
        cur[m] = (uint32_t)__viaddmax_s32((int32_t)h, (int32_t)e, (int32_t)cur[m]);
      }
      if (L == G - 1 && (uint32_t)(scol + u) < store_n2)
        scratch[scol + u] = (int32_t)(((u & 1) ? B[K - 1] : A[K - 1]) - floor_);
      upPrev = up;
      fm1 = floor_;
      floor_ += 2u * SC;
      e -= 2u * SC + 1u;
    }
  }
}

template <int K, int MINB, bool BYTES>
static int launch_long_t(const BatchView& b, int ctas, LaunchCfg& lc, int slot, cudaStream_t st)
{
  LongArgs a;
  a.q_pk = b.q_pk; a.r_pk = b.r_pk; a.q_bytes = b.q_bytes; a.r_bytes = b.r_bytes;
  a.q_beg = b.q_beg; a.q_end = b.q_end; a.r_beg = b.r_beg; a.r_end = b.r_end;
  a.list = BYTES ? b.bytes_list : b.long_list; a.counters = b.counters; a.out = b.out;
  a.scratch = b.scratch; a.scratch_stride = b.scratch_stride;
  constexpr size_t RING = 4 * 32 * K;
  const size_t smem = 9 * 128 + 4 * (RING * 6 + 16 * sizeof(LongJob)) + (size_t)2 * K * 128 * 4;
  if (!lc.attr_set[slot]) {                              // function attributes are per device: once per context
    cudaFuncSetAttribute(sw_long_kernel<K, MINB, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lc.attr_set[slot] = true;
  }
  sw_long_kernel<K, MINB, BYTES><<<ctas, 128, smem, st>>>(a);
  return 1;
}

// persistent grids (ctas = a multiple of the SM count unless the scratch clamp reduced it), work-stealing over the long /
// the bytes list; the two launches use disjoint scratch halves
// max_read_len <= 192: every job is a single band, 32 x 6 rows waste fewer lanes than 32 x 10 on 150 bp reads
int launch_long(const BatchView& b, int ctas, uint32_t max_read_len, LaunchCfg& lc, cudaStream_t st)
{
  if (max_read_len <= 192) return launch_long_t<6, 4, false>(b, ctas, lc, 16, st);
#ifdef SWB_ALL_VARIANTS
  switch (lc.long_k) {                                   // SWB_LONG_K, measured on 10 kb pairs: K=10 4019, 12 4113, 14 3959, 16 3658 GCUPS
    case 10: return launch_long_t<10, 4, false>(b, ctas, lc, 17, st);
    case 14: return launch_long_t<14, 4, false>(b, ctas, lc, 18, st);
    case 16: return launch_long_t<16, 3, false>(b, ctas * 3 / 4, lc, 19, st);
    default: break;
  }
#endif
  return launch_long_t<12, 4, false>(b, ctas, lc, 20, st);
}
int launch_long_bytes(const BatchView& b, int ctas, uint32_t max_read_len, LaunchCfg& lc, cudaStream_t st)
{
  return max_read_len <= 192 ? launch_long_t<6, 4, true>(b, ctas, lc, 21, st) : launch_long_t<10, 4, true>(b, ctas, lc, 22, st);
}

// =====================================================================================
// Generic 32-bit kernel: one warp per pair, any length, any bytes (raw byte equality).
// Anti-diagonal wavefront: lane L holds KG rows of a 32*KG-row band, slot m of lane L
// computes cell (i, j = t - (KG*L+m)) at step t; the row above a lane's first row comes
// from lane L-1 by SHFL; the last row of a band is parked in a scratch row in global
// memory and becomes the top boundary of the next band.  Tracking is exact and explicit:
// per row (score, first column), folded per band with strict '>' in ascending row order.
// =====================================================================================
constexpr int KG = 8;
constexpr int BAND = 32 * KG;

__global__ void __launch_bounds__(128)
sw_generic_kernel(BatchView b, const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_list_ptr,
                  int32_t* __restrict__ last_row_out)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int32_t* scratch = b.scratch + (uint64_t)warp_global * b.scratch_stride;
  const uint32_t n_list = *n_list_ptr;

  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(&b.counters->generic_cursor, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_list) break;
    const uint32_t p = list[item];
    const uint8_t* __restrict__ q = b.q_bytes + b.q_beg[p];
    const uint8_t* __restrict__ r = b.r_bytes + b.r_beg[p];
    const uint64_t n1 = b.q_end[p] - b.q_beg[p];
    const uint64_t n2 = b.r_end[p] - b.r_beg[p];

    int32_t gbest = 0; int64_t gi = -1, gj = -1; int32_t lastrow = 0;
    const uint64_t n_bands = (n1 + BAND - 1) / BAND;
    for (uint64_t band = 0; band < n_bands; ++band) {
      const uint64_t row0 = band * BAND + (uint64_t)KG * lane;
      int32_t qv[KG], A[KG], B[KG], W[KG], best[KG]; int64_t bt[KG];
#pragma unroll
      for (int m = 0; m < KG; ++m) {
        qv[m] = (row0 + m < n1) ? (int32_t)q[row0 + m] : 0x100;     // sentinel never equals a byte
        A[m] = 0; B[m] = 0; W[m] = 0x200; best[m] = 0; bt[m] = 0;
      }
      int32_t upPrev = 0;
      const uint64_t steps = (n2 + BAND - 1 + KG - 1) / KG * KG;
      for (uint64_t t0 = 0; t0 < steps; t0 += KG) {
#pragma unroll
        for (int u = 0; u < KG; ++u) {
          const uint64_t t = t0 + u;
          const int64_t jc = (int64_t)t - (int64_t)(KG * lane);     // column of this lane's slot 0
          W[u] = (jc >= 0 && (uint64_t)jc < n2) ? (int32_t)__ldg(r + jc) : 0x200;
          int32_t up = __shfl_up_sync(0xffffffffu, (u & 1) ? A[KG - 1] : B[KG - 1], 1);
          if (lane == 0) up = (band > 0 && t < n2) ? __ldcg(scratch + t) : 0;
#pragma unroll
          for (int m = KG - 1; m >= 0; --m) {
            const int32_t s = (qv[m] == W[(u - m + KG) % KG]) ? kMatch : kMismatch;
            int32_t d, uu, l;
            if (u & 1) { d = m ? B[m - 1] : upPrev; uu = m ? A[m - 1] : up; l = A[m]; }
            else       { d = m ? A[m - 1] : upPrev; uu = m ? B[m - 1] : up; l = B[m]; }
            const int32_t x = max(uu, l) + kGap;
            int32_t h = __viaddmax_s32_relu(d, s, x);
            // cells left of column 0 / right of the last column must stay neutral
            const int64_t j = (int64_t)t - (int64_t)(KG * lane + m);
            if (j < 0 || (uint64_t)j >= n2) h = 0;
            if (u & 1) B[m] = h; else A[m] = h;
            if (h > best[m]) { best[m] = h; bt[m] = j; }
          }
          // park the band's last row for the next band
          if (lane == 31) {
            const int64_t j = (int64_t)t - (int64_t)(BAND - 1);
            if (j >= 0 && (uint64_t)j < n2) scratch[j] = (u & 1) ? B[KG - 1] : A[KG - 1];
          }
          upPrev = up;
        }
      }
#pragma unroll
      for (int m = 0; m < KG; ++m) {
        if (best[m] > gbest) { gbest = best[m]; gi = (int64_t)(row0 + m); gj = bt[m]; }
        if (row0 + m == n1 - 1) lastrow = best[m];
      }
      __syncwarp();
    }
    // warp reduction: max score, then min i, then min j
    int32_t lr = lastrow;
    for (int o = 16; o; o >>= 1) {
      const int32_t os = __shfl_xor_sync(0xffffffffu, gbest, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, gi, o), oj = __shfl_xor_sync(0xffffffffu, gj, o);
      if (os > gbest || (os == gbest && os > 0 && (oi < gi || (oi == gi && oj < gj)))) { gbest = os; gi = oi; gj = oj; }
      lr = max(lr, __shfl_xor_sync(0xffffffffu, lr, o));
    }
    if (lane == 0) {
      b.out[p] = gbest > 0 ? swb_result{gbest, (int32_t)gi, (int32_t)gj} : swb_result{0, -1, -1};
      if (last_row_out) last_row_out[item] = lr;
    }
    __syncwarp();
  }
}

int launch_generic(const BatchView& b, int sm_count, int /*warps_resident*/, cudaStream_t st)
{
  // persistent grid: 4 CTAs x 4 warps per SM, work-stealing over the generic list
  sw_generic_kernel<<<sm_count * 4, 128, 0, st>>>(b, b.generic_list, &b.counters->n_generic, nullptr);
  return 1;
}

__global__ void single_pair_setup_kernel(Counters* c, uint32_t* list)
{
  c->n_short = 0; c->n_generic = 1; c->max_short_window = 0; c->generic_cursor = 0; c->n_long = 0; c->long_cursor = 0; c->n_bytes = 0; c->bytes_cursor = 0; c->stream_cursor = 0; c->n_overflow = 0; c->n_mid = 0; c->max_mid_window = 0; c->mid_cursor = 0; c->max_mid_read = 0; list[0] = 0;
}

// exposed for the C API: run the generic kernel on a prepared single-pair view
int launch_generic_single(const BatchView& b, int32_t* last_row_out, cudaStream_t st)
{
  single_pair_setup_kernel<<<1, 1, 0, st>>>(b.counters, b.generic_list);
  sw_generic_kernel<<<1, 32, 0, st>>>(b, b.generic_list, &b.counters->n_generic, last_row_out);
  return 2;
}

// =====================================================================================
// The reference's LIVE kernel, restated for CUDA (smith_waterman.cl:11-71): work-group g
// owns [g*chunk, min((g+1)*chunk, L)), work-item lid walks it with stride wgs keeping a
// clamped running sum of +2/-1; the maximum over everything lands in *result.
// =====================================================================================
__global__ void ref_compat_kernel(const uint8_t* __restrict__ s1, const uint8_t* __restrict__ s2,
                                  uint64_t len, uint64_t chunk, int32_t* result)
{
  __shared__ int32_t wmax[32];
  const uint64_t start = (uint64_t)blockIdx.x * chunk;                 // cl:27
  int32_t mx = 0, curv = 0;
  if (start < len) {                                                   // cl:30-32
    const uint64_t end = min(start + chunk, len);                      // cl:28
    for (uint64_t i = start + threadIdx.x; i < end; i += blockDim.x) { // cl:39
      const int32_t s = (s1[i] == s2[i]) ? kMatch : kMismatch;         // cl:43-47
      curv = max(curv + s, 0);                                         // cl:50
      mx = max(mx, curv);                                              // cl:51
    }
  }
  for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    int32_t v = (threadIdx.x < (blockDim.x + 31) / 32) ? wmax[threadIdx.x] : 0;
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) atomicMax(result, v);                        // cl:68-70
  }
}

int launch_ref_compat(const uint8_t* s1, const uint8_t* s2, uint64_t len, uint32_t wgs, uint64_t groups,
                      int32_t* result, cudaStream_t st)
{
  cudaMemsetAsync(result, 0, sizeof(int32_t), st);
  if (len == 0) return 0;
  const uint64_t chunk = (len + groups - 1) / groups;                  // cl:26
  ref_compat_kernel<<<(unsigned)groups, wgs, 0, st>>>(s1, s2, len, chunk, result);
  return 1;
}

// =====================================================================================
// Synthetic workload.  Counter RNG: the k-th draw of stream `seed` for pair
// p is splitmix64 evaluated at state  (seed ^ p*0x9E3779B97F4A7C15) + (k+1)*0x9E3779B97F4A7C15.
// Windows: iid ACGT, 32 bases per draw of stream 0xB200.  Reads, distribution 0 (related):
// cut from the window at offset o = draw(0xB201,0) % (W-n+1), then per read base i one draw
// x = draw(0xB201, 1+i):  x%1000 == 0 -> 1-base insertion (random base, cursor stays),
// x%1000 == 1 -> 1-base deletion (cursor skips one), and (x>>10)%100 == 0 -> substitution.
// Distribution 1 (unrelated): read base i = 2 bits of draw(0xB201, 1 + i/32).
// With a reference (launch_synth_ref) the window of pair p is instead cut from it at draw(0xB202, 0) % (ref_len - W + 1)
// -- windows of neighbouring reads overlap, as in a genome -- and the read is made from that window by the same rule.
// =====================================================================================
__host__ __device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t p, uint64_t k)
{
  uint64_t z = (seed ^ (p * 0x9E3779B97F4A7C15ull)) + (k + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t window_code(uint64_t p, uint32_t j)
{
  return (uint32_t)(splitmix_at(0xB200ull, p, j >> 5) >> (2 * (j & 31))) & 3u;
}

__global__ void synth_window_kernel(uint64_t first_pair, uint64_t n_pairs, uint32_t wlen,
                                    uint8_t* __restrict__ r_bytes, uint64_t* __restrict__ r_off,
                                    const uint8_t* __restrict__ ref, uint64_t ref_len, uint64_t* __restrict__ win_start)
{
  const uint64_t total = n_pairs * wlen;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += stride) {
    const uint64_t k = x / wlen; const uint32_t j = (uint32_t)(x - k * wlen);
    if (ref) r_bytes[x] = ref[splitmix_at(0xB202ull, first_pair + k, 0) % (ref_len - wlen + 1) + j];
    else     r_bytes[x] = "ACGT"[window_code(first_pair + k, j)];
  }
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_pairs; k += stride) {
    r_off[k] = k * wlen;
    if (ref && win_start && k < n_pairs) win_start[k] = splitmix_at(0xB202ull, first_pair + k, 0) % (ref_len - wlen + 1);
  }
}

// the read of pair k is made from ITS window as synth_window_kernel wrote it (w_bytes + k*wlen), whatever filled it
__global__ void synth_read_kernel(uint64_t first_pair, uint64_t n_pairs, uint32_t rlen, uint32_t wlen, int dist,
                                  uint8_t* __restrict__ q_bytes, uint64_t* __restrict__ q_off, const uint8_t* __restrict__ w_bytes)
{
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_pairs; k += stride) {
    q_off[k] = k * rlen;
    if (k == n_pairs) break;
    const uint64_t p = first_pair + k;
    uint8_t* dst = q_bytes + k * rlen;
    if (dist == 1) {
      for (uint32_t i = 0; i < rlen; ++i)
        dst[i] = "ACGT"[(uint32_t)(splitmix_at(0xB201ull, p, 1 + (i >> 5)) >> (2 * (i & 31))) & 3u];
    } else {
      const uint32_t span = wlen >= rlen ? wlen - rlen + 1 : 1;
      uint32_t c = (uint32_t)(splitmix_at(0xB201ull, p, 0) % span);
      for (uint32_t i = 0; i < rlen; ++i) {
        const uint64_t x = splitmix_at(0xB201ull, p, 1 + i);
        const uint32_t ev = (uint32_t)(x % 1000u);
        uint32_t code;
        if (ev == 0) {
          code = (uint32_t)(x >> 32) & 3u;                       // insertion: cursor stays
        } else {
          if (ev == 1) ++c;                                      // deletion: skip one window base
          if (c < wlen) { const uint32_t t = ((uint32_t)w_bytes[k * wlen + c] >> 1) & 3u; code = t ^ (t >> 1); }   // A C G T -> 0 1 2 3
          else code = (uint32_t)(x >> 34) & 3u;
          ++c;
          if ((uint32_t)((x >> 10) % 100u) == 0) code = (code + 1u + (uint32_t)((x >> 20) % 3u)) & 3u;
        }
        dst[i] = "ACGT"[code];
      }
    }
  }
}

int launch_synth(uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len, int distribution,
                 uint8_t* q_bytes, uint64_t* q_off, uint8_t* r_bytes, uint64_t* r_off, cudaStream_t st)
{
  synth_window_kernel<<<148 * 8, 256, 0, st>>>(first_pair, n_pairs, window_len, r_bytes, r_off, nullptr, 0, nullptr);
  synth_read_kernel<<<148 * 8, 256, 0, st>>>(first_pair, n_pairs, read_len, window_len, distribution, q_bytes, q_off, r_bytes);
  return 2;
}

int launch_synth_ref(const uint8_t* ref, uint64_t ref_len, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len,
                     int distribution, uint8_t* q_bytes, uint64_t* q_off, uint8_t* r_bytes, uint64_t* r_off, uint64_t* win_start, cudaStream_t st)
{
  synth_window_kernel<<<148 * 8, 256, 0, st>>>(first_pair, n_pairs, window_len, r_bytes, r_off, ref, ref_len, win_start);
  synth_read_kernel<<<148 * 8, 256, 0, st>>>(first_pair, n_pairs, read_len, window_len, distribution, q_bytes, q_off, r_bytes);
  return 2;
}

}  // namespace swb
