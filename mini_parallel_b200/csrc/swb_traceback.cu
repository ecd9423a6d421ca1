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
// Rows are computed 32 columns at a time: T = max(0, diag + s, up + gap) per lane, then the left dependency
// H[j] = max(T[j], H[j-1] + gap) as a prefix maximum of T[j] - gap*j across the warp (five shuffles).  Two ballots per 32
// cells store the 2-bit directions; the walk back reads them, run-length encodes the operations and writes them in
// alignment order at a place reserved with one atomic add.
#include "swb_kernels.cuh"
#include <cstdint>

namespace swb {

__host__ __device__ __forceinline__ uint64_t tb_groups(uint64_t width) { return (width + 31) >> 5; }

// bytes of per-warp scratch a pair with `rows` rows and `width` columns needs
uint64_t tb_scratch_bytes(uint64_t rows, uint64_t width)
{
  return 2 * (width + 2) * 4 + rows * tb_groups(width) * 8 + (rows + width + 1) * 4 + 64;
}

__global__ void __launch_bounds__(128)
traceback_kernel(TracebackArgs a)
{
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint8_t* my = a.scratch + warp * a.scratch_per_warp;
  for (;;) {
    unsigned long long pair = 0;
    if (lane == 0) pair = atomicAdd(&a.cursor[0], 1ull);
    pair = __shfl_sync(0xffffffffu, pair, 0);
    if (pair >= a.n_pairs) break;
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
    const uint32_t c0 = (uint32_t)res.end_j + 1 - Wd, G = (uint32_t)tb_groups(Wd);
    int32_t* Hp = reinterpret_cast<int32_t*>(my);                              // Hp[jj+1] = H[i-1][jj], Hp[0] = the border
    int32_t* Hc = Hp + (Wd + 2);
    uint32_t* dirs = reinterpret_cast<uint32_t*>(Hc + (Wd + 2));               // two ballots per 32 cells
    uint32_t* runs = dirs + (uint64_t)R * G * 2;
    for (uint32_t x = lane; x <= Wd; x += 32) Hp[x] = 0;
    __syncwarp();
    for (uint32_t i = 0; i < R; ++i) {
      const uint32_t qi = q[i];
      int32_t carry = -kGapAbs;                                                // H[i][-1] - gap*(-1): the zero border, one column out
      for (uint32_t g = 0; g < G; ++g) {
        const uint32_t jj = g * 32 + lane;
        const bool valid = jj < Wd;
        const int32_t diag = valid ? Hp[jj] : 0, up = valid ? Hp[jj + 1] : 0;
        const int32_t s = (valid && qi == (uint32_t)r[c0 + jj]) ? kMatch : kMismatch;
        const int32_t t = max(max(diag + s, up + kGap), 0);
        int32_t b = valid ? t + kGapAbs * (int32_t)jj : INT32_MIN / 2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t v = __shfl_up_sync(0xffffffffu, b, o); if (lane >= (uint32_t)o) b = max(b, v); }
        b = max(b, carry);
        carry = __shfl_sync(0xffffffffu, b, 31);
        const int32_t h = b - kGapAbs * (int32_t)jj;
        const uint32_t d = !valid || h == 0 ? 0u : (h == diag + s ? 1u : (h == up + kGap ? 2u : 3u));
        if (valid) Hc[jj + 1] = h;
        const uint32_t b0 = __ballot_sync(0xffffffffu, d & 1u), b1 = __ballot_sync(0xffffffffu, d & 2u);
        if (lane == 0) { dirs[((uint64_t)i * G + g) * 2] = b0; dirs[((uint64_t)i * G + g) * 2 + 1] = b1; }
      }
      if (lane == 0) Hc[0] = 0;
      __syncwarp();
      int32_t* t2 = Hp; Hp = Hc; Hc = t2;
    }
    if (Hp[Wd] != res.score) { al.status = 1; if (lane == 0) a.out[pair] = al; continue; }   // not this pair's end cell
    // ---- walk back (every lane, same steps), operations run-length encoded end -> start ----
    int64_t i = res.end_i, jj = (int64_t)Wd - 1;
    uint32_t n_runs = 0, cur_op = 0, cur_len = 0;
    while (i >= 0 && jj >= 0) {
      const uint64_t wi = ((uint64_t)i * G + ((uint64_t)jj >> 5)) * 2;
      const uint32_t sh = (uint32_t)jj & 31u;
      const uint32_t d = ((dirs[wi] >> sh) & 1u) | (((dirs[wi + 1] >> sh) & 1u) << 1);
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

int launch_traceback(const TracebackArgs& a, int warps, cudaStream_t st)
{
  if (a.n_pairs == 0) return 0;
  const int ctas = (warps + 3) / 4;
  traceback_kernel<<<ctas, 128, 0, st>>>(a);
  return 1;
}

}  // namespace swb
