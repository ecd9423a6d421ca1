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
// start at a multiple of 512 bytes of the buffer (128-bit loads; one warp = 32 chunks of 16 bytes = one word of the
// non-ACGT bitmap); bytes outside [begin, end) count as filler.
// Line index of a byte = number of '\n' before it; the sequence line of record r is line 4r+1 (aligner.rs:138:
// line_count % 4 == 2 with a 1-based count).
// Everything works on 32-bit words (4 bytes per instruction), not bytes: newlines are rare (4 per record), so a thread
// walks the 0-2 newlines of its 16 bytes instead of its 16 bytes.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kTile = 4096;
constexpr uint32_t kFillWord = 0x41414141u;                 // "AAAA"

__host__ __device__ __forceinline__ uint64_t fq_tile0(uint64_t begin) { return begin & ~511ull; }

// 0x80 in every byte of x that equals the byte replicated in pat (exact: no borrow between bytes)
__device__ __forceinline__ uint32_t eq_flags(uint32_t x, uint32_t pat)
{
  const uint32_t y = x ^ pat;
  return ~(((y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | y | 0x7F7F7F7Fu);
}
// the four 0x80 flags of a word as bits 0..3
__device__ __forceinline__ uint32_t flags_to_nibble(uint32_t f) { return ((f >> 7) * 0x01020408u) >> 24; }
// bits 0..3 as four byte masks (0xFF / 0x00)
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t n) { return (((n & 15u) * 0x00204081u) & 0x01010101u) * 0xFFu; }

struct FqChunk { uint32_t w[4]; uint32_t nl; };             // 16 bytes of text; nl: bit k = byte k is '\n'

__device__ __forceinline__ FqChunk fq_load(const uint8_t* __restrict__ text, uint64_t pos, uint64_t begin, uint64_t end)
{
  FqChunk c;
  if (pos + 16 <= begin || pos >= end) { c.w[0] = c.w[1] = c.w[2] = c.w[3] = kFillWord; c.nl = 0; return c; }
  const uint4 v = *reinterpret_cast<const uint4*>(text + pos);
  c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
  if (pos < begin || pos + 16 > end) {                       // the first / last chunk of the text: filler outside the range
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (pos + k < begin || pos + k >= end) c.w[k >> 2] = (c.w[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (0x41u << (8 * (k & 3)));
  }
  c.nl = flags_to_nibble(eq_flags(c.w[0], 0x0A0A0A0Au)) | (flags_to_nibble(eq_flags(c.w[1], 0x0A0A0A0Au)) << 4) |
         (flags_to_nibble(eq_flags(c.w[2], 0x0A0A0A0Au)) << 8) | (flags_to_nibble(eq_flags(c.w[3], 0x0A0A0A0Au)) << 12);
  return c;
}

__global__ void __launch_bounds__(256)
fq_count_kernel(const uint8_t* __restrict__ text, uint64_t begin, uint64_t end, uint32_t* __restrict__ tile_count, uint32_t* __restrict__ flags,
                int final_segment)
{
  const uint64_t pos = fq_tile0(begin) + (uint64_t)blockIdx.x * kTile + threadIdx.x * 16;
  const FqChunk c = fq_load(text, pos, begin, end);
  uint32_t n = __popc(c.nl), hi = (c.w[0] | c.w[1] | c.w[2] | c.w[3]) & 0x80808080u;
  // A '\r' that is not the first half of "\r\n" (mid-line, or ending an unterminated last line) is a byte of its line for
  // BufRead::lines; the in-place masking below only knows line terminators, so such a file is left to the host reader.
  const uint32_t cr = flags_to_nibble(eq_flags(c.w[0], 0x0D0D0D0Du)) | (flags_to_nibble(eq_flags(c.w[1], 0x0D0D0D0Du)) << 4) |
                      (flags_to_nibble(eq_flags(c.w[2], 0x0D0D0D0Du)) << 8) | (flags_to_nibble(eq_flags(c.w[3], 0x0D0D0D0Du)) << 12);
  if (cr) {
    uint32_t stray = cr & ~(c.nl >> 1) & 0x7FFFu;           // bytes 0..14: the next byte is in this chunk
    if (cr & 0x8000u) {                                     // byte 15: the next byte is the neighbour's (text is still untouched here)
      if (pos + 16 < end) stray |= text[pos + 16] != '\n';
      else stray |= final_segment ? 1u : 0u;                // last byte of the text: of a non-final segment it is carried and seen again
    }
    if (pos + 16 > end && !final_segment) stray &= (1u << (uint32_t)(end - 1 - pos)) - 1u;   // same for a chunk the text ends in
    if (stray) hi |= 0x80u;                                 // (the extra '\r' check after the last real byte sees filler 'A', not '\n')
  }
  __shared__ uint32_t wsum[8];
  n = __reduce_add_sync(0xffffffffu, n); hi = __reduce_or_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = n; if (hi) atomicOr(flags, 1u); }   // not ASCII: host path
  __syncthreads();
  if (threadIdx.x == 255) {                                   // newlines of the tile (<= 4096) | its last byte << 24: the next tile's
    uint32_t t = 0;                                           // first thread needs it to strip "\r\n" without reading text that
    for (int w = 0; w < 8; ++w) t += wsum[w];                 // fq_extract_kernel is rewriting
    tile_count[blockIdx.x] = t | (c.w[3] & 0xFF000000u);
  }
}

// exclusive scan of the tile counts by one CTA (a few hundred thousand tiles at most), 1024 tiles per round: coalesced
// loads, a shuffle scan per warp, the warp totals scanned by the first warp; total[0] = number of newlines
__global__ void __launch_bounds__(1024)
fq_scan_kernel(const uint32_t* __restrict__ tile_count, uint64_t n_tiles, uint64_t* __restrict__ tile_prefix, uint64_t* __restrict__ total)
{
  __shared__ uint32_t wsum[32], round_total;
  __shared__ uint64_t base;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (uint64_t t0 = 0; t0 < n_tiles; t0 += 1024) {
    const uint64_t t = t0 + threadIdx.x;
    const uint32_t v = t < n_tiles ? tile_count[t] & 0xFFFFFFu : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += x; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= (uint32_t)o) winc += x; }
      wsum[lane] = winc - w;                                   // exclusive: newlines of the warps before this one
      if (lane == 31) round_total = winc;
    }
    __syncthreads();
    const uint64_t b0 = base;
    if (t < n_tiles) tile_prefix[t] = b0 + wsum[warp] + (inc - v);
    __syncthreads();
    if (threadIdx.x == 0) base = b0 + round_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = base;
}

// One pass over the text that does everything the scoring pipeline needs from it:
//  * sequence-line ranges: seq_beg[r] / seq_end[r] are positions in the buffer ("\r\n" stripped like BufRead::lines);
//    with `final_segment` an unterminated last sequence line still counts as a read;
//  * tail_start = first byte after the last complete record of a non-final segment (the line index says which newline
//    that is: newline 4*floor(N/4), N = all newlines of the text -- no atomics); the bytes from there on stay as they
//    are, the host carries them into the next segment;
//  * everything else that is not a base of a sequence line becomes 'A' in the text (the byte-compare kernels read it in
//    place), so that only real non-ACGT bases are flagged;
//  * the 2-bit packed text and its non-ACGT bitmap (what pack2bit_kernel would make of the masked text).
__global__ void __launch_bounds__(256)
fq_extract_kernel(uint8_t* __restrict__ text, uint64_t begin, uint64_t end, const uint32_t* __restrict__ tile_count, const uint64_t* __restrict__ tile_prefix,
                  const uint64_t* __restrict__ n_newlines, uint64_t* __restrict__ seq_beg, uint64_t* __restrict__ seq_end, uint64_t n_records_cap,
                  unsigned long long* __restrict__ tail_start, int final_segment, uint32_t* __restrict__ pk_words, uint32_t* __restrict__ pk_bitmap)
{
  const uint64_t pos = fq_tile0(begin) + (uint64_t)blockIdx.x * kTile + threadIdx.x * 16;
  const FqChunk c = fq_load(text, pos, begin, end);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // line index at the first byte of this thread's chunk
  const uint32_t n = __popc(c.nl);
  uint32_t inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += v; }
  __shared__ uint32_t wtot[8], wlast[8];
  if (lane == 31) { wtot[warp] = inc; wlast[warp] = c.w[3] >> 24; }
  __syncthreads();
  uint32_t before = 0;
#pragma unroll
  for (uint32_t w = 0; w < 8; ++w) before += w < warp ? wtot[w] : 0u;
  // the byte before this chunk, as it was before anybody masked it (for "\r\n" at a chunk boundary)
  uint32_t prev_byte = __shfl_up_sync(0xffffffffu, c.w[3] >> 24, 1);
  if (lane == 0) prev_byte = warp ? wlast[warp - 1] : (blockIdx.x ? tile_count[blockIdx.x - 1] >> 24 : 0u);
  uint64_t line = tile_prefix[blockIdx.x] + before + (inc - n);
  const uint64_t keep_line = final_segment ? ~0ull : (*n_newlines & ~3ull);     // lines from here on are the carried tail

  // walk the newlines of the chunk: seq (bit k: byte k is a base of a sequence line), tail (bit k: byte k is carried)
  uint32_t seq = 0, tail = 0, rest = c.nl, a = 0;
  for (;;) {
    const uint32_t b = rest ? (uint32_t)__ffs((int)rest) - 1u : 16u;      // next newline, or the end of the chunk
    const uint32_t span = ((1u << b) - 1u) & ~((1u << a) - 1u);            // bytes [a, b) lie on line `line`
    if (line >= keep_line) tail |= span | (b < 16 ? 1u << b : 0u);
    else if ((line & 3) == 1) seq |= span;
    if (b == 16) break;
    const uint64_t p = pos + b, rec = line >> 2;
    const uint32_t ph = (uint32_t)line & 3u;
    if (rec < n_records_cap) {
      if (ph == 0) seq_beg[rec] = p + 1;
      if (ph == 1) {
        const uint32_t prev = b ? (c.w[(b - 1) >> 2] >> (8 * ((b - 1) & 3))) & 0xFFu : prev_byte;
        seq_end[rec] = p - (prev == '\r' && p > begin ? 1 : 0);           // lines() strips "\r\n"
      }
    }
    if (line + 1 == keep_line) *tail_start = (unsigned long long)(p + 1);
    rest &= rest - 1; a = b + 1; ++line;
  }
  if (final_segment && end > pos && end <= pos + 16) {       // the text's last byte: a last line without a newline (BufRead::lines yields it)
    const uint32_t k = (uint32_t)(end - 1 - pos);
    const uint64_t l = line - __popc(c.nl >> k);               // `line` is past the whole chunk: back to byte k's line ...
    if (!((c.nl >> k) & 1u) && (l & 3) == 1 && (l >> 2) < n_records_cap) seq_end[l >> 2] = end;
  }
  // '\r' on a sequence line is not a base
  const uint32_t crf[4] = {eq_flags(c.w[0], 0x0D0D0D0Du), eq_flags(c.w[1], 0x0D0D0D0Du), eq_flags(c.w[2], 0x0D0D0D0Du), eq_flags(c.w[3], 0x0D0D0D0Du)};
  if (crf[0] | crf[1] | crf[2] | crf[3])
    seq &= ~(flags_to_nibble(crf[0]) | (flags_to_nibble(crf[1]) << 4) | (flags_to_nibble(crf[2]) << 8) | (flags_to_nibble(crf[3]) << 12));
  const uint32_t keep = seq | tail;
  uint32_t m[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { const uint32_t km = nibble_to_bytes(keep >> (4 * i)); m[i] = (c.w[i] & km) | (kFillWord & ~km); }
  if (pos + 16 > begin && pos < end) *reinterpret_cast<uint4*>(text + pos) = make_uint4(m[0], m[1], m[2], m[3]);
  uint32_t bad = 0;
  const uint32_t word = pack4(m[0], bad) | (pack4(m[1], bad) << 8) | (pack4(m[2], bad) << 16) | (pack4(m[3], bad) << 24);
  pk_words[pos >> 4] = word;
  const uint32_t ballot = __ballot_sync(0xffffffffu, bad != 0);
  if (lane == 0) pk_bitmap[pos >> 9] = ballot;
}

int launch_fq_index(const uint8_t* text, uint64_t begin, uint64_t end, uint32_t* tile_count, uint64_t* tile_prefix, uint64_t* total,
                    uint32_t* flags, int final_segment, cudaStream_t st)
{
  if (end <= begin) return 0;
  const uint64_t n_tiles = fq_tiles(begin, end);
  fq_count_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(text, begin, end, tile_count, flags, final_segment);
  fq_scan_kernel<<<1, 1024, 0, st>>>(tile_count, n_tiles, tile_prefix, total);
  return 2;
}

// pk_words / pk_bitmap: room for every 16-byte word / 512-byte group of text[0, end rounded up to a tile)
int launch_fq_extract(uint8_t* text, uint64_t begin, uint64_t end, const uint32_t* tile_count, const uint64_t* tile_prefix, const uint64_t* n_newlines, uint64_t* seq_beg,
                      uint64_t* seq_end, uint64_t n_records_cap, unsigned long long* tail_start, int final_segment, uint32_t* pk_words,
                      uint32_t* pk_bitmap, cudaStream_t st)
{
  if (end <= begin) return 0;
  const uint64_t n_tiles = fq_tiles(begin, end);
  fq_extract_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(text, begin, end, tile_count, tile_prefix, n_newlines, seq_beg, seq_end, n_records_cap, tail_start,
                                                       final_segment, pk_words, pk_bitmap);
  return 1;
}

uint64_t fq_tiles(uint64_t begin, uint64_t end) { return end > begin ? (end - fq_tile0(begin) + kTile - 1) / kTile : 0; }

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
                 unsigned long long* __restrict__ sums /* [0] score sum, [1] bases, [4] longest read */)
{
  unsigned long long sc = 0, bs = 0, mx = 0;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long len = seq_end[k] - seq_beg[k];
    sc += (unsigned long long)(long long)res[k].score; bs += len; mx = len > mx ? len : mx;
  }
  for (int o = 16; o; o >>= 1) {
    sc += __shfl_xor_sync(0xffffffffu, sc, o); bs += __shfl_xor_sync(0xffffffffu, bs, o);
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, mx, o); mx = other > mx ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], sc); atomicAdd(&sums[1], bs); if (mx) atomicMax(&sums[4], mx); }
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
