// swb_kernels.cuh -- device-side interface of the Smith-Waterman engine (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/swb200.h"

namespace swb {

// Scoring constants: smith_waterman.cl:5-7.
constexpr int kMatch = 2, kMismatch = -1, kGap = -2;
constexpr int kGapAbs = -kGap;

// ---- routing classes written by classify_pairs ----
enum : uint8_t { CLASS_EMPTY = 0, CLASS_SHORT = 1, CLASS_GENERIC = 2, CLASS_LONG = 3, CLASS_BYTES = 4, CLASS_MID = 5 };

// Limits of the int16x2 inter-task kernel (see sw_short_kernel).
constexpr uint32_t kShortMaxRead   = 160;    // rows held by one lane group (G x K)
constexpr uint32_t kMidMaxRead     = 320;    // rows of the one-group-per-warp instantiation (value scale 32, see sw_stream_kernel)
constexpr uint32_t kShortMaxWindow = 4096;   // columns staged in shared memory per group
constexpr uint32_t kLongMaxLen = (1u << 20) - 4096;   // rows / columns the 64-bit end-cell key of sw_long_kernel can hold

struct Counters {             // device-resident, zeroed per batch
  uint32_t n_short;           // pairs routed to the int16x2 kernel
  uint32_t n_generic;         // pairs routed to the 32-bit kernel
  uint32_t max_short_window;  // longest window among the short pairs
  uint32_t generic_cursor;    // work-stealing cursor of the generic kernel
  uint32_t n_long;            // ACGT-only pairs too long for the int16x2 kernel: 32-bit banded wavefront kernel
  uint32_t long_cursor;       // work-stealing cursor of the long kernel
  uint32_t n_bytes;           // pairs touching a non-ACGT byte (any length below kLongMaxLen): the same kernel on raw bytes
  uint32_t bytes_cursor;
  uint32_t stream_cursor;     // sw_stream_kernel: couples handed out beyond every group's static ones
  uint32_t n_overflow;        // pairs whose window is longer than BatchView::max_window (result INT32_MIN): the caller's bound was wrong
  uint32_t n_mid;             // ACGT-only pairs with reads of 161..320 bp: the 320-row int16x2 instantiation of sw_stream_kernel
  uint32_t max_mid_window;    // longest window among them
  uint32_t mid_cursor;        // its couple cursor
  uint32_t max_mid_read;      // longest read among them: <= 256 bp and the 256-row instantiation scores the list, else the 320-row one
  uint32_t pad_[2];
};

struct ShortDesc {            // one entry per short-listed pair, written by classify_kernel (32 B, 16-aligned)
  uint64_t q0, r0;            // first base of the read / of the window (base coordinates of the packed arrays)
  uint32_t n, m;              // read length, window length
  uint32_t pair, pad;         // index of the pair in the batch
};

struct BatchView {            // everything the kernels need about one batch (device pointers)
  // ASCII reads / windows.  Sequence p occupies bytes [beg[p], end[p]).  For CSR offsets end == beg + 1; for
  // windows cut from a device-resident reference beg/end are independent arrays (windows may overlap).
  const uint8_t*  q_bytes;  const uint64_t* q_beg;  const uint64_t* q_end;
  const uint8_t*  r_bytes;  const uint64_t* r_beg;  const uint64_t* r_end;
  const uint32_t* q_pk;     const uint32_t* q_bad;     // 2-bit packed reads  + non-ACGT bitmap (1 bit / 16-base word)
  const uint32_t* r_pk;     const uint32_t* r_bad;     // 2-bit packed windows + bitmap
  uint64_t        n_pairs;
  uint32_t*       short_list;   // pair ids, n_short entries
  ShortDesc*      short_desc;   // descriptors, same order as short_list
  ShortDesc*      mid_desc;     // descriptors of the 161..320 bp reads (n_mid entries); nullptr: no such path, they go to the long kernel
  uint32_t*       generic_list; // pair ids, n_generic entries
  uint32_t*       long_list;    // pair ids, n_long entries
  uint32_t*       bytes_list;   // pair ids, n_bytes entries
  int             force_bytes;  // debug/bench knob: route every non-short pair to the byte-compare kernel
  uint32_t        short_max_read;   // rows of the stream-kernel instantiation this batch uses: 160, or 128 when the caller's bound on
                                    // the read length allows it (2 x 100 / 2 x 125 bp runs); longer reads are routed past it
  Counters*       counters;
  swb_result*     out;
  int32_t*        scratch;      // generic kernel: one boundary row per resident warp
  uint64_t        scratch_stride;
  uint32_t        max_window;   // windows longer than this are refused (the scratch rows hold scratch_stride >= max_window columns)
};

#if defined(__CUDACC__)
// 2-bit packing of four bytes.  code = (byte >> 1) & 3 : A->0 C->1 T->2 G->3 (any bijection works, only equality
// matters).  `bad` collects the bits of every byte that is not exactly A/C/G/T, because the reference compares raw bytes
// (cl:114): 'a' != 'A', 'N' == 'N'.
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t& bad)
{
  const uint32_t c0 = (x >> 1) & 0x01010101u, c1 = (x >> 2) & 0x01010101u;
  const uint32_t is2 = c1 & ~c0;
  const uint32_t canon = 0x41414141u + 2u * c0 + 4u * c1 + 15u * is2;   // A,C,G,T rebuilt from the code
  bad |= x ^ canon;
  uint32_t t = (x >> 1) & 0x03030303u;
  t |= t >> 6;
  return (t & 0xFu) | ((t >> 12) & 0xF0u);
}
#endif

// Launch state of one context (function attributes set, occupancy asked, tuning knobs from the environment): owned by the
// swb_ctx, which one thread uses at a time -- no process-wide mutable state in the launchers.
struct LaunchCfg {
  int  sm_count = 0;
  int  resident[16] = {};          // sw_stream_kernel variant -> resident CTAs per SM (0 = not asked yet)
  bool attr_set[32] = {};          // kernel slot -> cudaFuncSetAttribute done
  int  stream_ctas_per_sm = 0;     // SWB_STREAM_CTAS_PER_SM
  long stream_grid = 0;            // SWB_STREAM_GRID
  int  long_k = 0;                 // SWB_LONG_K (builds with -DSWB_ALL_VARIANTS)
  bool debug = false;              // SWB_DEBUG
};
void launch_cfg_init(LaunchCfg& lc, int sm_count);
bool short_variant_available(int variant);

// launchers (all asynchronous on `st`); each returns the number of kernels it launched
int launch_pack2bit(const uint8_t* bytes, uint64_t n, uint32_t* words, uint32_t* bitmap, cudaStream_t st);
int launch_classify(const BatchView& b, cudaStream_t st);
int launch_chunk_prepare(uint64_t* off_a, uint64_t n_a, uint64_t base_a, uint64_t len_a, uint64_t* off_b, uint64_t n_b, uint64_t base_b,
                         uint64_t len_b, uint64_t* win_beg, const uint32_t* win_len, uint32_t len_w, uint64_t* win_end, uint64_t n_w, uint64_t win_base, cudaStream_t st);
int launch_short(const BatchView& b, uint32_t window_cap, int variant, LaunchCfg& lc, cudaStream_t st);
int launch_mid(const BatchView& b, LaunchCfg& lc, cudaStream_t st);
int launch_generic(const BatchView& b, int sm_count, int warps_resident, cudaStream_t st);
int launch_long(const BatchView& b, int ctas, uint32_t max_read_len, LaunchCfg& lc, cudaStream_t st);
int launch_long_bytes(const BatchView& b, int ctas, uint32_t max_read_len, LaunchCfg& lc, cudaStream_t st);
int launch_ref_compat(const uint8_t* s1, const uint8_t* s2, uint64_t len, uint32_t wgs, uint64_t groups,
                      int32_t* result, cudaStream_t st);
int launch_synth(uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len, int distribution,
                 uint8_t* q_bytes, uint64_t* q_off, uint8_t* r_bytes, uint64_t* r_off, cudaStream_t st);
int launch_synth_ref(const uint8_t* ref, uint64_t ref_len, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len,
                     int distribution, uint8_t* q_bytes, uint64_t* q_off, uint8_t* r_bytes, uint64_t* r_off, uint64_t* win_start, cudaStream_t st);
// alignments behind the scores (swb_traceback.cu)
struct TracebackArgs {
  const uint8_t* q; const uint64_t* qo; const uint8_t* r; const uint64_t* ro;     // CSR pairs (device)
  const swb_result* res;                     // their scores and end cells
  swb_alignment* out;
  uint32_t* cigar; uint64_t cigar_cap;       // operation words; alignments reserve their slices with an atomic add
  unsigned long long* cursor;                // [0] next pair, [1] operation words reserved (may exceed cigar_cap)
  uint64_t n_pairs;
  uint8_t* scratch; uint64_t scratch_per_warp;   // global: directions (dirs_per_warp bytes) [+ rows and runs when not in shared memory]
  uint64_t dirs_per_warp, rows_per_warp; int rows_in_smem;
  uint8_t* handled;                          // per pair: done by traceback_diag_kernel (gapless), the matrix kernel skips it
};
uint32_t tb_pick_cpl(uint64_t max_width);
uint64_t tb_rows_bytes(uint64_t rows, uint64_t width, uint32_t cpl);
uint64_t tb_dirs_bytes(uint64_t rows, uint64_t width, uint32_t cpl);
int launch_traceback(const TracebackArgs& a, uint32_t cpl, int warps, cudaStream_t st);
// FASTQ.gz ingest on the GPU (swb_fastq_kernels.cu)
uint64_t fq_tiles(uint64_t begin, uint64_t end);
int launch_inflate_bgzf(const uint8_t* comp, const swb_bgzf_block* blocks, uint64_t n_blocks, const uint64_t* out_off, uint8_t* text,
                        uint32_t* n_failed, cudaStream_t st);
int launch_fq_index(const uint8_t* text, uint64_t begin, uint64_t end, uint32_t* tile_count, uint64_t* tile_prefix, uint64_t* total,
                    uint32_t* flags, int final_segment, cudaStream_t st);
int launch_fq_extract(uint8_t* text, uint64_t begin, uint64_t end, const uint32_t* tile_count, const uint64_t* tile_prefix, const uint64_t* n_newlines, uint64_t* seq_beg,
                      uint64_t* seq_end, uint64_t n_records_cap, unsigned long long* tail_start, int final_segment, uint32_t* pk_words,
                      uint32_t* pk_bitmap, cudaStream_t st);
int launch_fq_windows(uint64_t file_index, uint64_t first_read, uint64_t n, uint64_t ref_len, uint32_t w, uint64_t* seq_beg, uint64_t* seq_end,
                      uint64_t* win_beg, uint64_t* win_end, cudaStream_t st);
int launch_fq_reduce(const swb_result* res, const uint64_t* seq_beg, const uint64_t* seq_end, uint64_t n, unsigned long long* sums, cudaStream_t st);

}  // namespace swb
