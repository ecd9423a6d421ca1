// swb_traceback.cu -- the alignment behind a score: start cell and CIGAR of pairs whose best local
// score and end cell are known (swb_score_*).  The reference has no such output (gpu_align returns one i32,
// aligner.rs:410); the rule is this repository's (include/swb200.h, restated by the checker): walk back from the end cell while H > 0,
// at every cell the first predecessor that explains its value in the order diagonal, up ('I'), left ('D').
//
// One warp per pair (work stealing).  Only the part of the matrix an alignment ending at (end_i, end_j) can touch is
// recomputed: rows 0..end_i, and the last 2*(end_i+1) columns up to end_j -- a local alignment with a positive score has
// fewer gap steps than matches (2m - x - 2g >= 1), so it spans at most twice as many columns as rows.  Inside that
// rectangle (zero borders) the values ON the path equal the full matrix's and every off-path neighbour is <= its true
// value, so each "does this predecessor explain H" test comes out as in the full matrix.
// A row is computed in spans of 32*CPL columns, CPL consecutive columns per lane: T = max(0, diag + s, up + gap) per cell,
// then the left dependency H[j] = max(T[j], H[j-1] + gap) as a prefix maximum of T[j] - gap*j -- inside the lane's
// cells, one five-shuffle scan across the warp per span, and a fix-up.  The 2-bit directions of a lane's cells are one
// word (global scratch, L2-resident: 19 KB per 150 x 300 rectangle); the walk back reads them, run-length encodes the
// operations and writes them in alignment order at a place reserved with one atomic add.  The two H rows and the runs
// live in shared memory when they fit (6 KB per warp at that shape, so 32 warps per SM hide the walk's load latency).
#include "swb_kernels.cuh"
#include <cstdint>

namespace swb {

// CPL = consecutive columns per lane (4, 8, 12 or 16: the narrowest that covers the widest rectangle of the batch in one span
// of 32*CPL columns, else 16 and several spans).  H values of one span of a row: lane l's cells at l*(CPL+1) + 1 .. + CPL
// (an odd stride: no bank conflicts), the cell left of them at l*(CPL+1).
__host__ __device__ __forceinline__ uint32_t tb_row_words(uint32_t cpl) { return 32 * (cpl + 1) + 1; }
__host__ __device__ __forceinline__ uint64_t tb_spans(uint64_t width, uint32_t cpl) { return (width + 32 * cpl - 1) / (32 * cpl); }

uint32_t tb_pick_cpl(uint64_t max_width) { return max_width <= 128 ? 4 : max_width <= 256 ? 8 : max_width <= 384 ? 12 : 16; }
// per-warp scratch of a pair with `rows` rows and `width` columns: the two H rows + the runs (small: shared memory when it
// fits), and the directions (one word per lane, span and row: global, L2-resident)
uint64_t tb_rows_bytes(uint64_t rows, uint64_t width, uint32_t cpl) { return 2 * tb_spans(width, cpl) * tb_row_words(cpl) * 4 + (rows + width + 1) * 4 + 64; }
uint64_t tb_dirs_bytes(uint64_t rows, uint64_t width, uint32_t cpl) { return rows * tb_spans(width, cpl) * 32 * 4; }

template <uint32_t CPL>
__global__ void __launch_bounds__(128)
traceback_kernel(TracebackArgs a)
{
  extern __shared__ __align__(16) uint8_t tb_smem[];
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  constexpr uint32_t kTbCols = CPL, kTbSpan = 32 * CPL, kTbRow = 32 * (CPL + 1) + 1;
  uint8_t* my_dirs = a.scratch + warp * a.scratch_per_warp;
  uint8_t* my = a.rows_in_smem ? tb_smem + (threadIdx.x >> 5) * a.rows_per_warp : my_dirs + a.dirs_per_warp;
  for (;;) {
    unsigned long long pair = 0;
    if (lane == 0) pair = atomicAdd(&a.cursor[0], 1ull);
    pair = __shfl_sync(0xffffffffu, pair, 0);
    if (pair >= a.n_pairs) break;
    if (a.handled[pair]) continue;                                            // traceback_diag_kernel wrote this alignment
    const swb_result res = a.res[pair];
    const uint8_t* q = a.q + a.qo[pair]; const uint64_t n1 = a.qo[pair + 1] - a.qo[pair];
    const uint8_t* r = a.r + a.ro[pair]; const uint64_t n2 = a.ro[pair + 1] - a.ro[pair];
    swb_alignment al; al.start_i = -1; al.start_j = -1; al.cigar_len = 0; al.status = 0; al.cigar_off = 0;
    if (res.score <= 0) { if (lane == 0) a.out[pair] = al; continue; }       // nothing aligned: (-1,-1), no operations
    if (res.end_i < 0 || res.end_j < 0 || (uint64_t)res.end_i >= n1 || (uint64_t)res.end_j >= n2) {
      al.status = 1; if (lane == 0) a.out[pair] = al; continue;                // not a result of this pair
    }
    const uint32_t R = (uint32_t)res.end_i + 1;
    const uint64_t w_all = (uint64_t)res.end_j + 1, w_cap = 2ull * R;
    const uint32_t Wd = (uint32_t)(w_all < w_cap ? w_all : w_cap);
    const uint32_t c0 = (uint32_t)res.end_j + 1 - Wd, SP = (uint32_t)tb_spans(Wd, CPL);
    int32_t* Hp = reinterpret_cast<int32_t*>(my);                              // the row above, span by span (layout: kTbRow)
    int32_t* Hc = Hp + (uint64_t)SP * kTbRow;
    uint32_t* runs = reinterpret_cast<uint32_t*>(Hc + (uint64_t)SP * kTbRow);
    uint32_t* dirs = reinterpret_cast<uint32_t*>(my_dirs);                     // dirs[(i*SP + span)*32 + lane]: CPL cells x 2 bits
    for (uint32_t x = lane; x < SP * kTbRow; x += 32) Hp[x] = 0;
    __syncwarp();
    // the window bytes and the valid-cell mask of a lane's cells do not depend on the row: with one span (the usual case)
    // they are loaded once per pair.  A cell past the rectangle's last column holds a byte that equals nothing: such cells
    // lie right of every real cell of their row, so what they compute reaches nothing that is kept.
    uint32_t wb[kTbCols], vmask = 0;
    auto load_window = [&](uint32_t sp) {
      vmask = 0;
#pragma unroll
      for (uint32_t k = 0; k < kTbCols; ++k) {
        const uint32_t jj = sp * kTbSpan + lane * kTbCols + k;
        wb[k] = jj < Wd ? (uint32_t)r[c0 + jj] : 0x100u;
        vmask |= (jj < Wd ? 3u : 0u) << (2 * k);
      }
    };
    if (SP == 1) load_window(0);
    for (uint32_t i = 0; i < R; ++i) {
      const uint32_t qi = q[i];
      int32_t carry = -kGapAbs;                      // (H + gap*j) one column left of the row: the zero border at j = -1
      int32_t left_h = 0;                            // H left of the span's first cell
      for (uint32_t sp = 0; sp < SP; ++sp) {
        if (SP > 1) load_window(sp);
        const int32_t g0 = kGapAbs * (int32_t)(sp * kTbSpan + lane * kTbCols);  // gap * (this lane's first column)
        const int32_t* hp = Hp + (uint64_t)sp * kTbRow + lane * (kTbCols + 1);
        int32_t* hc = Hc + (uint64_t)sp * kTbRow + lane * (kTbCols + 1);
        int32_t up[kTbCols + 1];                                               // up[0] = the cell diagonal to the lane's first
#pragma unroll
        for (uint32_t k = 0; k <= kTbCols; ++k) up[k] = hp[k];
        int32_t b[kTbCols], ts[kTbCols];                                       // prefix maxima of T + gap*j; diag + s per cell
        int32_t run = INT32_MIN / 2;
#pragma unroll
        for (uint32_t k = 0; k < kTbCols; ++k) {
          ts[k] = up[k] + (qi == wb[k] ? kMatch : kMismatch);
          const int32_t t = max(max(ts[k], up[k + 1] + kGap), 0);
          run = max(run, t + g0 + kGapAbs * (int32_t)k);
          b[k] = run;
        }
        // exclusive prefix maximum of the lanes' totals, seeded with the carry of the previous span
        int32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc = max(inc, v); }
        int32_t before = __shfl_up_sync(0xffffffffu, inc, 1);
        before = lane ? max(before, carry) : carry;
        carry = max(__shfl_sync(0xffffffffu, inc, 31), carry);
        int32_t prev_h = __shfl_up_sync(0xffffffffu, max(b[kTbCols - 1], before) - g0 - kGapAbs * (int32_t)(kTbCols - 1), 1);
        if (lane == 0) prev_h = left_h;                                        // H of the cell left of the lane's first
        uint32_t word = 0;
        int32_t h = 0;
#pragma unroll
        for (uint32_t k = 0; k < kTbCols; ++k) {
          h = max(b[k], before) - g0 - kGapAbs * (int32_t)k;
          const uint32_t d = h == 0 ? 0u : (h == ts[k] ? 1u : (h == up[k + 1] + kGap ? 2u : 3u));
          word |= d << (2 * k);
          hc[k + 1] = h;
        }
        hc[0] = prev_h;
        left_h = __shfl_sync(0xffffffffu, h, 31);
        dirs[((uint64_t)i * SP + sp) * 32 + lane] = word & vmask;
      }
      __syncwarp();
      int32_t* t2 = Hp; Hp = Hc; Hc = t2;
    }
    {
      const uint32_t je = Wd - 1;
      const int32_t h_end = Hp[(uint64_t)(je / kTbSpan) * kTbRow + ((je % kTbSpan) / kTbCols) * (kTbCols + 1) + 1 + (je % kTbCols)];
      if (h_end != res.score) { al.status = 1; if (lane == 0) a.out[pair] = al; __syncwarp(); continue; }   // not this pair's end cell
    }
    // ---- walk back (every lane, same steps), operations run-length encoded end -> start ----
    int64_t i = res.end_i, jj = (int64_t)Wd - 1;
    uint32_t n_runs = 0, cur_op = 0, cur_len = 0;
    while (i >= 0 && jj >= 0) {
      const uint32_t word = dirs[((uint64_t)i * SP + (uint64_t)jj / kTbSpan) * 32 + ((uint64_t)jj % kTbSpan) / kTbCols];
      const uint32_t d = (word >> (2 * ((uint32_t)jj % kTbCols))) & 3u;
      if (d == 0) break;
      al.start_i = (int32_t)i; al.start_j = (int32_t)(c0 + jj);
      uint32_t op;
      if (d == 1) { op = q[i] == r[c0 + jj] ? 7u : 8u; --i; --jj; }
      else if (d == 2) { op = 1u; --i; }
      else { op = 2u; --jj; }
      if (op == cur_op) ++cur_len;
      else { if (cur_len && lane == 0) runs[n_runs] = (cur_len << 4) | cur_op; n_runs += cur_len ? 1 : 0; cur_op = op; cur_len = 1; }
    }
    if (cur_len) { if (lane == 0) runs[n_runs] = (cur_len << 4) | cur_op; ++n_runs; }
    __syncwarp();
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(&a.cursor[1], (unsigned long long)n_runs);
    off = __shfl_sync(0xffffffffu, off, 0);
    al.cigar_len = n_runs; al.cigar_off = off;
    if (off + n_runs <= a.cigar_cap)
      for (uint32_t k = lane; k < n_runs; k += 32) a.cigar[off + k] = runs[n_runs - 1 - k];
    else al.status = 2;                                                        // the caller's buffer is too small (cursor[1] says how much is needed)
    if (lane == 0) a.out[pair] = al;
    __syncwarp();
  }
}

// Gapless alignments without the matrix (one thread per pair).  Walking back along the diagonal from the end cell, let
// sum(t) be the score of its last t cells.  If sum(k) == score for some k, the alignment IS those k cells: every H on the
// diagonal is at least the score of the diagonal segment that ends there (H[t] >= score - sum(t)) and, by the recurrence,
// at most its successor's value minus the successor's substitution score (H[t] <= H[t-1] - s), which from H[0] = score
// pins H[t] = score - sum(t); so the diagonal predecessor explains every cell (it is asked first), and the walk ends at the
// first k with sum(k) == score, where H reaches 0.  (sum(t) > score would mean H[end] > score: not this pair's result.)
// Most reads of a real run align without a gap, and are done here in a few hundred byte compares.
__global__ void __launch_bounds__(256)
traceback_diag_kernel(TracebackArgs a)
{
  const uint64_t pair = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= a.n_pairs) return;
  a.handled[pair] = 0;
  const swb_result res = a.res[pair];
  const uint8_t* q = a.q + a.qo[pair]; const uint64_t n1 = a.qo[pair + 1] - a.qo[pair];
  const uint8_t* r = a.r + a.ro[pair]; const uint64_t n2 = a.ro[pair + 1] - a.ro[pair];
  if (res.score <= 0 || res.end_i < 0 || res.end_j < 0 || (uint64_t)res.end_i >= n1 || (uint64_t)res.end_j >= n2) return;
  const uint32_t ei = (uint32_t)res.end_i, ej = (uint32_t)res.end_j, max_k = min(ei, ej) + 1;
  int32_t sum = 0;
  uint32_t k = 0, n_runs = 0, prev = 2;
  for (uint32_t t = 0; t < max_k; ++t) {
    const uint32_t eq = q[ei - t] == r[ej - t];
    sum += eq ? kMatch : kMismatch;
    n_runs += eq != prev; prev = eq;
    if (sum == res.score) { k = t + 1; break; }
    if (sum > res.score || sum + kMatch * (int32_t)(max_k - 1 - t) < res.score) break;
  }
  if (!k) return;                                              // a gap, or not a result of this pair: the matrix decides
  swb_alignment al;
  al.start_i = (int32_t)(ei + 1 - k); al.start_j = (int32_t)(ej + 1 - k); al.cigar_len = n_runs; al.status = 0;
  al.cigar_off = atomicAdd(&a.cursor[1], (unsigned long long)n_runs);
  if (al.cigar_off + n_runs <= a.cigar_cap) {
    uint64_t at = al.cigar_off;
    uint32_t cur = q[al.start_i] == r[al.start_j], len = 0;
    for (uint32_t t = 0; t < k; ++t) {
      const uint32_t eq = q[al.start_i + t] == r[al.start_j + t];
      if (eq != cur) { a.cigar[at++] = (len << 4) | (cur ? 7u : 8u); cur = eq; len = 0; }
      ++len;
    }
    a.cigar[at] = (len << 4) | (cur ? 7u : 8u);
  } else al.status = 2;
  a.out[pair] = al;
  a.handled[pair] = 1;
}

int launch_traceback(const TracebackArgs& a, uint32_t cpl, int warps, cudaStream_t st)
{
  if (a.n_pairs == 0) return 0;
  traceback_diag_kernel<<<(unsigned)((a.n_pairs + 255) / 256), 256, 0, st>>>(a);
  const int ctas = (warps + 3) / 4;
  const size_t smem = a.rows_in_smem ? (size_t)a.rows_per_warp * 4 : 0;
  switch (cpl) {
    case 4:  traceback_kernel<4><<<ctas, 128, smem, st>>>(a); break;
    case 8:  traceback_kernel<8><<<ctas, 128, smem, st>>>(a); break;
    case 12: traceback_kernel<12><<<ctas, 128, smem, st>>>(a); break;
    default: traceback_kernel<16><<<ctas, 128, smem, st>>>(a); break;
  }
  return 2;
}

}  // namespace swb
