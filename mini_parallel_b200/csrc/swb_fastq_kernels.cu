// swb_fastq_kernels.cu -- FASTQ.gz ingest on the GPU: blocked-gzip (BGZF) members are inflated one warp
// per block, the text is indexed in place (newline count -> scan -> sequence-line ranges) and everything that is not
// a base of a sequence line is overwritten with 'A', so the scoring pipeline can take the reads where they lie:
// read k = text[seq_beg[k], seq_end[k]).  Replaces the `zcat` child + per-line String loop of
// process_fastq_file_in_chunks (smith_waterman/src/aligner.rs:107-178) when the input is BGZF.
#include "swb_kernels.cuh"
#include "swb_inflate.cuh"
#include <cstdlib>

namespace swb {

// ------------------------------------------------------------------------------------------------
// inflate: one warp per BGZF block, 8 warps per CTA, decode tables in shared memory
// ------------------------------------------------------------------------------------------------
template <bool SOLO>
__global__ void __launch_bounds__(256, 4)            // 4 CTAs = 32 warps = 32 blocks in flight per SM (shared memory allows no more)
inflate_bgzf_kernel(const uint8_t* __restrict__ comp, const swb_bgzf_block* __restrict__ blocks, uint64_t n_blocks,
                    const uint64_t* __restrict__ out_off, uint8_t* __restrict__ text, uint32_t* __restrict__ n_failed)   // n_failed[0] count, [2] first status, [3] its produced
{
  __shared__ swi::Tables tables[8];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t b = (uint64_t)blockIdx.x * 8 + warp;
  if (b >= n_blocks) return;
  const swb_bgzf_block blk = blocks[b];
  swi::Lanes L{SOLO ? 0 : (int)lane, SOLO ? 1 : 32};   // SOLO (debug): every lane does all the work redundantly
  uint32_t produced = 0;
  int st = swi::inflate_member(comp + blk.in_off, blk.in_len, text + out_off[b], blk.out_len, &produced, tables[warp], L);
  if (st == swi::OK && produced != blk.out_len) st = swi::ERR_LENGTH_MISMATCH;
  if (st != swi::OK && lane == 0) { if (atomicAdd(n_failed, 1u) == 0) { n_failed[2] = (uint32_t)st; n_failed[3] = produced; } }
}

int launch_inflate_bgzf(const uint8_t* comp, const swb_bgzf_block* blocks, uint64_t n_blocks, const uint64_t* out_off, uint8_t* text,
                        uint32_t* n_failed, cudaStream_t st)
{
  if (n_blocks == 0) return 0;
  if (getenv("SWB_INFLATE_SOLO")) inflate_bgzf_kernel<true><<<(unsigned)((n_blocks + 7) / 8), 256, 0, st>>>(comp, blocks, n_blocks, out_off, text, n_failed);
  else inflate_bgzf_kernel<false><<<(unsigned)((n_blocks + 7) / 8), 256, 0, st>>>(comp, blocks, n_blocks, out_off, text, n_failed);
  return 1;
}

// ------------------------------------------------------------------------------------------------
// FASTQ index.  The text of a segment is text[begin, end); it starts at the first byte of a record.  Tiles of 4096 bytes
// are aligned to 16 bytes of the buffer (128-bit loads); bytes outside [begin, end) count as filler.
// Line index of a byte = number of '\n' before it; the sequence line of record r is line 4r+1 (aligner.rs:138:
// line_count % 4 == 2 with a 1-based count).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTile = 4096;

__device__ __forceinline__ void load16(const uint8_t* __restrict__ text, uint64_t pos, uint64_t begin, uint64_t end, uint8_t (&b)[16])
{
  const uint4 v = (pos + 16 <= begin || pos >= end) ? make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u)
                                                    : *reinterpret_cast<const uint4*>(text + pos);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint8_t c = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    b[k] = (pos + k >= begin && pos + k < end) ? c : (uint8_t)'A';
  }
}

__global__ void __launch_bounds__(256)
fq_count_kernel(const uint8_t* __restrict__ text, uint64_t begin, uint64_t end, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ flags)
{
  const uint64_t base = (begin & ~15ull) + (uint64_t)blockIdx.x * kTile + threadIdx.x * 16;
  uint8_t b[16];
  load16(text, base, begin, end, b);
  uint32_t n = 0, hi = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) { n += b[k] == '\n'; hi |= b[k]; }
  __shared__ uint32_t wsum[8];
  for (int o = 16; o; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); hi |= __shfl_xor_sync(0xffffffffu, hi, o); }
  if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = n; if (hi & 0x80u) atomicOr(flags, 1u); }   // not ASCII: host path
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += wsum[w];
    tile_count[blockIdx.x] = t;
  }
}

// exclusive scan of the tile counts by one CTA (a few hundred thousand tiles at most); total[0] = number of newlines
__global__ void __launch_bounds__(1024)
fq_scan_kernel(const uint32_t* __restrict__ tile_count, uint64_t n_tiles, uint64_t* __restrict__ tile_prefix, uint64_t* __restrict__ total)
{
  __shared__ uint64_t part[1024];
  const uint64_t per = (n_tiles + 1023) / 1024;
  const uint64_t lo = min(n_tiles, (uint64_t)threadIdx.x * per), hi = min(n_tiles, lo + per);
  uint64_t s = 0;
  for (uint64_t t = lo; t < hi; ++t) s += tile_count[t];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t run = 0;
    for (int k = 0; k < 1024; ++k) { const uint64_t v = part[k]; part[k] = run; run += v; }
    total[0] = run;
  }
  __syncthreads();
  uint64_t run = part[threadIdx.x];
  for (uint64_t t = lo; t < hi; ++t) { tile_prefix[t] = run; run += tile_count[t]; }
}

// line index at the first byte of this thread's 16 bytes
__device__ __forceinline__ uint64_t thread_line_base(const uint8_t (&b)[16], uint64_t tile_prefix)
{
  uint32_t n = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) n += b[k] == '\n';
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = n;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += v; }
  __shared__ uint32_t wtot[8];
  if (lane == 31) wtot[warp] = inc;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += wtot[w];
  return tile_prefix + before + (inc - n);
}

// Only record ends in the last kTailSearch bytes compete for tail_start: the tail a caller can carry is at most 1 MiB
// (swb_fastq_bgzf_score), so an older record end can only belong to a segment that is declined anyway -- and one atomic per
// record on a single address (2 M per segment) was most of this kernel's time.
constexpr uint64_t kTailSearch = 4ull << 20;

// sequence-line ranges (read-only pass).  seq_beg[r] / seq_end[r] are positions in the buffer; tail_start = first byte
// after the last complete record (atomicMax); with `final` an unterminated last sequence line still counts as a read.
__global__ void __launch_bounds__(256)
fq_extract_kernel(const uint8_t* __restrict__ text, uint64_t begin, uint64_t end, const uint64_t* __restrict__ tile_prefix,
                  uint64_t* __restrict__ seq_beg, uint64_t* __restrict__ seq_end, uint64_t n_records_cap,
                  unsigned long long* __restrict__ tail_start, int final_segment)
{
  const uint64_t base = (begin & ~15ull) + (uint64_t)blockIdx.x * kTile + threadIdx.x * 16;
  uint8_t b[16];
  load16(text, base, begin, end, b);
  uint64_t line = thread_line_base(b, tile_prefix[blockIdx.x]);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint64_t p = base + k;
    if (b[k] == '\n') {
      const uint64_t rec = line >> 2; const uint32_t ph = (uint32_t)(line & 3);
      if (rec < n_records_cap) {
        if (ph == 0) seq_beg[rec] = p + 1;
        if (ph == 1) {
          const bool cr = p > begin && (k ? b[k - 1] : text[p - 1]) == '\r';      // lines() strips "\r\n"
          seq_end[rec] = p - (cr ? 1 : 0);
        }
      }
      if (ph == 3 && p + kTailSearch >= end) atomicMax(tail_start, (unsigned long long)(p + 1));
      ++line;
    } else if (final_segment && p + 1 == end && (line & 3) == 1 && (line >> 2) < n_records_cap) {
      seq_end[line >> 2] = end;                                                 // last line without a newline (BufRead::lines yields it)
    }
  }
}

// everything that is not a base of a sequence line becomes 'A' (so the packing kernel flags real non-ACGT bases only)
// Bytes from *keep_from on (the incomplete record at the end of a non-final segment) stay as they are: the host carries
// them, unmasked, into the next segment.
__global__ void __launch_bounds__(256)
fq_mask_kernel(uint8_t* __restrict__ text, uint64_t begin, uint64_t end, const uint64_t* __restrict__ tile_prefix,
               const unsigned long long* __restrict__ keep_from)
{
  const uint64_t base = (begin & ~15ull) + (uint64_t)blockIdx.x * kTile + threadIdx.x * 16;
  const uint64_t keep = keep_from ? (uint64_t)*keep_from : ~0ull;
  uint8_t b[16];
  load16(text, base, begin, end, b);
  uint64_t line = thread_line_base(b, tile_prefix[blockIdx.x]);
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint8_t c = b[k];
    const bool nl = c == '\n';
    if ((nl || (line & 3) != 1 || c == '\r') && base + k < keep) c = 'A';
    if (nl) ++line;
    w[k >> 2] |= (uint32_t)c << (8 * (k & 3));
  }
  if (base + 16 > begin && base < end) *reinterpret_cast<uint4*>(text + base) = make_uint4(w[0], w[1], w[2], w[3]);
}

int launch_fq_index(const uint8_t* text, uint64_t begin, uint64_t end, uint32_t* tile_count, uint64_t* tile_prefix, uint64_t* total,
                    uint32_t* flags, cudaStream_t st)
{
  if (end <= begin) return 0;
  const uint64_t n_tiles = fq_tiles(begin, end);
  fq_count_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(text, begin, end, tile_count, flags);
  fq_scan_kernel<<<1, 1024, 0, st>>>(tile_count, n_tiles, tile_prefix, total);
  return 2;
}

int launch_fq_extract_mask(uint8_t* text, uint64_t begin, uint64_t end, const uint64_t* tile_prefix, uint64_t* seq_beg, uint64_t* seq_end,
                           uint64_t n_records_cap, unsigned long long* tail_start, int final_segment, cudaStream_t st)
{
  if (end <= begin) return 0;
  const uint64_t n_tiles = fq_tiles(begin, end);
  fq_extract_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(text, begin, end, tile_prefix, seq_beg, seq_end, n_records_cap, tail_start, final_segment);
  fq_mask_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(text, begin, end, tile_prefix, final_segment ? nullptr : tail_start);
  return 2;
}

uint64_t fq_tiles(uint64_t begin, uint64_t end) { return end > begin ? (end - (begin & ~15ull) + kTile - 1) / kTile : 0; }

// ------------------------------------------------------------------------------------------------
// per-batch helpers: the window every read is paired with (this engine's --full-wgs pairing rule, rustseq_host.cpp)
// and the reduction of a batch's results
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64_dev(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256)
fq_windows_kernel(uint64_t file_index, uint64_t first_read, uint64_t n, uint64_t ref_len, uint32_t w, uint64_t* __restrict__ seq_beg,
                  uint64_t* __restrict__ seq_end, uint64_t* __restrict__ win_beg, uint64_t* __restrict__ win_end)
{
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (seq_end[k] < seq_beg[k]) seq_end[k] = seq_beg[k];      // never hand the scoring kernels a negative length
  const uint64_t g = (file_index << 40) + first_read + k;
  const uint64_t s = splitmix64_dev(g ^ 0xB202ull) % (ref_len - w + 1);
  win_beg[k] = s; win_end[k] = s + w;
}

__global__ void __launch_bounds__(256)
fq_reduce_kernel(const swb_result* __restrict__ res, const uint64_t* __restrict__ seq_beg, const uint64_t* __restrict__ seq_end, uint64_t n,
                 unsigned long long* __restrict__ sums /* [0] score sum, [1] bases */)
{
  unsigned long long sc = 0, bs = 0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    sc += (unsigned long long)(long long)res[k].score; bs += seq_end[k] - seq_beg[k];
  }
  for (int o = 16; o; o >>= 1) { sc += __shfl_xor_sync(0xffffffffu, sc, o); bs += __shfl_xor_sync(0xffffffffu, bs, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], sc); atomicAdd(&sums[1], bs); }
}

int launch_fq_windows(uint64_t file_index, uint64_t first_read, uint64_t n, uint64_t ref_len, uint32_t w, uint64_t* seq_beg, uint64_t* seq_end,
                      uint64_t* win_beg, uint64_t* win_end, cudaStream_t st)
{
  if (n == 0) return 0;
  fq_windows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(file_index, first_read, n, ref_len, w, seq_beg, seq_end, win_beg, win_end);
  return 1;
}

int launch_fq_reduce(const swb_result* res, const uint64_t* seq_beg, const uint64_t* seq_end, uint64_t n, unsigned long long* sums, cudaStream_t st)
{
  if (n == 0) return 0;
  fq_reduce_kernel<<<148 * 4, 256, 0, st>>>(res, seq_beg, seq_end, n, sums);
  return 1;
}

}  // namespace swb
