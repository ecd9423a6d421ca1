/*
 * swb200.h -- C ABI of the B200-native Smith-Waterman scoring engine.
 *
 * This is the drop-in boundary for the alignment path of bmwoolf/mini_parallel.  The
 * reference exposes no FFI; its seam is one Rust function:
 *
 *     pub fn gpu_align(seq1: &str, seq2: &str, device: &GpuDevice) -> Result<i32, String>
 *                                                     smith_waterman/src/aligner.rs:410
 *
 * Every entry point below cites the reference item it replaces.  Conventions:
 *   - return 0 = Ok, non-zero = Err; the message (the Result<_, String> text) is in
 *     swb_last_error() (thread-local);
 *   - inputs are borrowed for the duration of the call, outputs are caller-allocated;
 *   - a swb_ctx is used from one thread at a time (the reference serialises on one
 *     queue behind a Mutex, gpu.rs:13-14, :97-115); distinct contexts are independent;
 *   - there is NO CPU fallback: without a CUDA device swb_create fails, like
 *     main.rs:76-79 / :160-163;
 *   - i indexes s1 / the read (rows), j indexes s2 / the window (columns), as
 *     seq1[i] / seq2[j] in smith_waterman.cl:114.  Coordinates are 0-based, inclusive;
 *     (-1,-1) when the score is 0.  Tie-break: max score, then smallest i, then smallest j
 *     (first maximum in a row-major scan; the reference has no coordinates).
 */
#ifndef SWB200_H
#define SWB200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct swb_ctx swb_ctx;                                   /* opaque: device, streams, pinned + device arenas */
typedef struct { int32_t score; int32_t end_i; int32_t end_j; } swb_result;
typedef struct { int32_t match, mismatch, gap; } swb_params;      /* smith_waterman.cl:5-7: {2,-1,-2}; only these are accepted */

/* Device probe: replaces gpu::is_gpu_available / get_gpu_devices (gpu.rs:33-94). */
int  swb_device_count(void);
/* name / memory / max work-group size of a device: the GpuDevice struct, gpu.rs:17-23. */
int  swb_device_info(int device_id, char* name, size_t name_cap, double* memory_gb, int* max_work_group_size);

/* Free / total device memory in bytes (system_info.rs:236-243 budgets 80 % of VRAM; the WGS report states what is in use). */
int  swb_memory_info(int device_id, uint64_t* free_bytes, uint64_t* total_bytes);

/* Context: replaces get_opencl_context/init_opencl (gpu.rs:97-132).  One device per context. */
int  swb_create(swb_ctx** out, int device_id, const swb_params* params /* NULL = reference constants */);
void swb_destroy(swb_ctx*);

/* One pair, full Smith-Waterman (the recurrence of smith_waterman.cl:114-125, global max + end cell). */
int  swb_score_pair(swb_ctx*, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, swb_result* out);

/* A batch of independent pairs, HOST buffers (ASCII bytes, CSR offsets with n_pairs+1 entries).
 * H2D, packing, scoring and D2H all happen inside the call.  This is what a chunk of reads from
 * process_fastq_file_in_chunks (aligner.rs:107-178) is handed to instead of one gpu_align per chunk. */
int  swb_score_batch(swb_ctx*, const uint8_t* q_bytes, const uint64_t* q_off,
                     const uint8_t* r_bytes, const uint64_t* r_off,
                     uint64_t n_pairs, swb_result* out);

/* Reads against windows of a DEVICE-RESIDENT reference (north_star: windows must not stream over PCIe).
 * swb_set_reference uploads and 2-bit-packs the reference once; each call then ships only the reads plus one
 * (start,len) per read.  Window k is reference[win_start[k], win_start[k]+win_len[k]).  The reference has no
 * read-vs-reference notion (it self-compares a chunk, aligner.rs:270-274); this is the pairing the
 * --full-wgs driver uses here. */
int  swb_set_reference(swb_ctx*, const uint8_t* ref_bytes, uint64_t n);
int  swb_score_batch_vs_reference(swb_ctx*, const uint8_t* q_bytes, const uint64_t* q_off, uint64_t n_pairs,
                                  const uint64_t* win_start, const uint32_t* win_len, swb_result* out);

/* Reads against windows that are RANGES of one HOST buffer: window k is w_bytes[win_start[k], win_start[k]+win_len[k]).
 * Ranges may overlap, repeat and come in any order -- the candidate windows a read mapper cuts from a genome.  The part of
 * the buffer the windows touch is uploaded and packed ONCE per call (not once per window), the reads are pipelined against
 * it in chunks like in swb_score_batch_vs_reference; results are identical to swb_score_batch on the materialised windows.
 * swb_last_ranges_info: bytes of the buffer the last such call uploaded, and the sum of its window lengths (what a CSR
 * layout of the same windows would have uploaded).  gpu_align copies both whole strings per call (aligner.rs:478-492). */
int  swb_score_batch_ranges(swb_ctx*, const uint8_t* q_bytes, const uint64_t* q_off, uint64_t n_pairs,
                            const uint8_t* w_bytes, uint64_t w_total_bytes, const uint64_t* win_start, const uint32_t* win_len,
                            swb_result* out);
int  swb_last_ranges_info(swb_ctx*, uint64_t* bytes_uploaded, uint64_t* window_bytes);

/* Several devices behind one handle (north_star: each chunk is partitioned across the GPUs of one box with per-GPU
 * streams; the reference uses devices[0] only, gpu.rs:117-131, main.rs:95).  device_ids == NULL or n_devices <= 0: every
 * visible device.  One persistent host thread and one swb_ctx (three pipeline streams) per device, created concurrently.
 * swb_multi_score_batch* cut the batch into contiguous slices of pairs of about equal bytes, one per device; every device
 * writes its slice of out[]: no inter-GPU traffic, no collective.  Results are identical to the single-device calls.
 * swb_multi_ctx gives the k-th device's context (tuning knobs, timings); do not score on it while a multi call runs. */
typedef struct swb_multi swb_multi;
int  swb_create_multi(swb_multi** out, const int* device_ids, int n_devices, const swb_params* params);
void swb_destroy_multi(swb_multi*);
int  swb_multi_device_count(swb_multi*);
swb_ctx* swb_multi_ctx(swb_multi*, int k);
int  swb_multi_score_batch(swb_multi*, const uint8_t* q_bytes, const uint64_t* q_off, const uint8_t* r_bytes, const uint64_t* r_off,
                           uint64_t n_pairs, swb_result* out);
int  swb_multi_set_reference(swb_multi*, const uint8_t* ref_bytes, uint64_t n);
int  swb_multi_score_batch_vs_reference(swb_multi*, const uint8_t* q_bytes, const uint64_t* q_off, uint64_t n_pairs,
                                        const uint64_t* win_start, const uint32_t* win_len, swb_result* out);

/* FASTQ.gz ingest on the GPU for blocked gzip (BGZF: bgzip, BCL Convert): replaces the `zcat` child and the per-line
 * String loop of process_fastq_file_in_chunks (aligner.rs:107-178) for the --full-wgs path.  The host only walks the
 * block headers; a segment of whole blocks is copied to the device, inflated one warp per block, indexed (every 4th
 * line + 2 is a read, aligner.rs:138) and scored against windows of the resident reference (window of read g of file f:
 * start = splitmix64(((f << 40) + g) ^ 0xB202) mod (ref_len - window_len + 1), the pairing rule of the WGS driver).
 *   blocks[k]      deflate payload of block k inside comp[] and its inflated size (the member's ISIZE)
 *   carry          text left over from the previous segment of the file: the bytes after its last complete record
 *   final_segment  no more data follows: an unterminated last line still counts (BufRead::lines)
 * Outputs: sum of the scores, reads, bases and lines consumed, and the new carry (at most carry_cap bytes).
 * *status = 0 ok; 1 = the data needs the host path (an inflate error, a non-ASCII byte, a carry larger than carry_cap):
 * nothing was scored, the caller falls back to zlib + rsm_process_fastq_file_in_chunks semantics. */
typedef struct { uint64_t in_off; uint32_t in_len; uint32_t out_len; } swb_bgzf_block;
/* Optional: start copying and inflating the NEXT segment on a second stream while the current one is being scored.  The
 * buffers must stay untouched until swb_fastq_bgzf_score is called with the same comp pointer and returns. */
int  swb_fastq_bgzf_prefetch(swb_ctx*, const uint8_t* comp, uint64_t comp_bytes, const swb_bgzf_block* blocks, uint64_t n_blocks);
/* A prefetched segment that will NOT be scored after all: waits for its copy and frees the slot (the buffer may then be reused). */
int  swb_fastq_bgzf_cancel(swb_ctx*, const uint8_t* comp);
int  swb_fastq_bgzf_score(swb_ctx*, const uint8_t* comp, uint64_t comp_bytes, const swb_bgzf_block* blocks, uint64_t n_blocks,
                          const uint8_t* carry, uint64_t carry_len, int final_segment,
                          uint64_t file_index, uint64_t first_read, uint32_t window_len,
                          int64_t* score_sum, uint64_t* n_reads, uint64_t* n_bases, uint64_t* n_lines,
                          uint8_t* carry_out, uint64_t carry_cap, uint64_t* carry_out_len, int* status);

/* Same, DEVICE-resident inputs and outputs (pointers from cudaMalloc / a torch tensor's data_ptr);
 * runs on the context's stream, returns after the work is enqueued; swb_sync() waits.
 * max_r_len is a HARD upper bound of the window lengths (the boundary-row scratch of the long-pair kernels is sized from
 * it before the device has seen the offsets): a pair whose window is longer is not scored -- its result is
 * (INT32_MIN, -1, -1) -- and swb_sync() / swb_last_routing() fail saying how many there were.  max_q_len is a hint only
 * (0 = unknown): it picks the row count of the kernel instantiations (128 rows for reads <= 128 bp, the 161..320 bp list,
 * the long-pair kernel's band height); a read longer than the hint is still scored exactly, by a slower kernel. */
int  swb_score_batch_device(swb_ctx*, const uint8_t* d_q_bytes, const uint64_t* d_q_off, uint64_t q_total_bytes,
                            const uint8_t* d_r_bytes, const uint64_t* d_r_off, uint64_t r_total_bytes,
                            uint64_t n_pairs, uint32_t max_q_len, uint32_t max_r_len, swb_result* d_out);
int  swb_sync(swb_ctx*);

/* The two literal behaviours of the reference, for comparison:
 *   swb_ref_compat_align : what gpu_align returns today -- kernel smith_waterman_align
 *                          (smith_waterman.cl:11-71) under the host geometry aligner.rs:422-424,
 *                          result buffer zero-initialised;
 *   swb_last_row_max     : the reduction of the never-launched smith_waterman_detailed
 *                          (smith_waterman.cl:130-134): max over the last row only. */
int  swb_ref_compat_align(swb_ctx*, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2,
                          uint32_t dev_max_work_group, int32_t* out);
int  swb_last_row_max(swb_ctx*, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, int32_t* out);

/* 2-bit packing stage on its own (A,C,G,T -> 2 bits, 16 bases per 32-bit word; one flag bit per word
 * for "contains a byte outside ACGT").  Host buffers in, host buffers out; used by tests and bench. */
int  swb_pack2bit(swb_ctx*, const uint8_t* bytes, uint64_t n, uint32_t* packed_words /* ceil(n/16) */,
                  uint32_t* nonacgt_bitmap /* ceil(ceil(n/16)/32) */);

/* Synthetic workload generator on the device (counter RNG splitmix64, seeds 0xB200 /
 * 0xB201; distribution 0 = related reads, 1 = unrelated).  Fills ASCII bytes + offsets for pairs
 * [first_pair, first_pair + n_pairs) of shape read_len x window_len.  Device pointers. */
int  swb_synth_device(swb_ctx*, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len,
                      int distribution, uint8_t* d_q_bytes, uint64_t* d_q_off, uint8_t* d_r_bytes, uint64_t* d_r_off);
/* The same with windows CUT FROM A REFERENCE (device pointer, ref_len bases): window p starts at draw(0xB202, 0) mod
 * (ref_len - window_len + 1), the read is made from it by the same rule.  Also writes the window starts. */
int  swb_synth_device_ref(swb_ctx*, const uint8_t* d_ref, uint64_t ref_len, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len,
                          uint32_t window_len, int distribution, uint8_t* d_q_bytes, uint64_t* d_q_off, uint8_t* d_r_bytes,
                          uint64_t* d_r_off, uint64_t* d_win_start);

/* Per-stage device times of the last swb_score_batch* call on this context, CUDA events on the
 * context's stream (ms): [0] pack, [1] short-read kernel, [2] generic kernel, [3] total device span,
 * [4] h2d, [5] d2h.  Also the number of kernels launched by that call. */
int  swb_last_timings(swb_ctx*, float* ms /* 6 */, int* kernels_launched);
/* Pairs routed to each path by the last call: [0] short-read int16x2 kernel, [1] generic 32-bit byte-compare kernel
 * (any bytes, any length), [2] long-pair 32-bit banded wavefront kernel (ACGT-only pairs beyond the short limits). */
int  swb_last_routing(swb_ctx*, uint64_t* counts /* 3 */);
/* The same by kernel: [0] int16x2 stream kernel, reads <= 160 bp; [1] its 256- / 320-row instantiations, reads of 161..320 bp
 * (value scale 32; 256 rows when no read of the list is longer than 256 bp); [2] long-pair kernel on 2-bit codes; [3] long-pair kernel on raw bytes (a non-ACGT byte in the pair);
 * [4] generic kernel (beyond 2^20 rows or columns).  [0] + [1] is swb_last_routing's [0], [3] + [4] its [1]. */
int  swb_last_routing_ex(swb_ctx*, uint64_t* counts /* 5 */);
/* on = 0: reads of 161..320 bp take the 32-bit long-pair kernel instead of the 320-row int16x2 one (comparison, tests). */
int  swb_set_mid_path(swb_ctx*, int on);

/* Debugging aid (compute-sanitizer is not available on every box).  With SWB_GUARD=1 in the environment when the library
 * is loaded, every device arena is allocated exactly (no geometric over-allocation, the documented 64 B of read slack only)
 * between two 4 KiB zones of a known pattern; this call synchronises the device and returns how many zones of the
 * context's device have been written into (0 = none), -1 when guard mode is off.  *n_arenas = arenas checked. */
int  swb_debug_guard_check(swb_ctx*, char* report, uint64_t report_cap, uint64_t* n_arenas);

/* The packing stage alone on device-resident bytes (bench: HBM roofline of the packing kernel). */
int  swb_pack2bit_device(swb_ctx*, const uint8_t* d_bytes, uint64_t n, uint32_t* d_words, uint32_t* d_bitmap);

/* The alignment behind a result.  Like the end cell, it does not exist upstream (gpu_align returns
 * one i32, aligner.rs:410, 531); the rule is this repository's (the checker restates it on the CPU): walk back from the end
 * cell while H > 0 (recurrence smith_waterman.cl:114-125), at every cell the first predecessor that explains its value
 * in the order diagonal, up, left.  Operations are BAM-style words (length << 4 | op) in alignment order:
 * '=' 7 equal bytes, 'X' 8 different bytes, 'I' 1 a base of s1 (read) against a gap, 'D' 2 a base of s2 (window) against
 * a gap.  results[] are what swb_score_* returned for the same pairs.  cigar[] receives all operations, alignment k's
 * are cigar[out[k].cigar_off .. + cigar_len) (slices in no particular order); *cigar_used = operations of the whole
 * batch.  If that exceeds cigar_cap the call fails and *cigar_used says how much room a retry needs.
 * out[k].status: 0 ok (start = (-1,-1) and no operations when the score is 0), 1 results[k] is not an end cell of pair k
 * (outside the pair, or the recomputed value differs from the score; a gapless diagonal that adds up to the score is taken
 * at its word without recomputing the matrix),
 * 2 (only when the call fails for lack of room) this alignment's operations did not fit. */
typedef struct { int32_t start_i, start_j; uint32_t cigar_len; uint32_t status; uint64_t cigar_off; } swb_alignment;
int  swb_traceback_batch(swb_ctx*, const uint8_t* q_bytes, const uint64_t* q_off, const uint8_t* r_bytes, const uint64_t* r_off,
                         uint64_t n_pairs, const swb_result* results, swb_alignment* out,
                         uint32_t* cigar, uint64_t cigar_cap, uint64_t* cigar_used);

/* Which instantiation of the short-read kernel runs.  9 (default, the only one in the product library): the streaming
 * kernel, 16 lanes x 10 rows, two-step end-cell tracker, couples handed out dynamically.  The test build (make variants,
 * -DSWB_ALL_VARIANTS) also carries 4..8, 10, 11 (earlier stream kernels) and 0..3 (the one-couple-per-group kernel) as
 * independent implementations the parity tests cross-check.  All variants return identical results; asking for one the
 * build does not have fails. */
int  swb_set_short_variant(swb_ctx*, int variant);

/* Host batches (swb_score_batch, swb_score_batch_vs_reference) are cut into chunks of about chunk_bytes of ASCII
 * input (at least min_chunk_pairs pairs each) that are pipelined over three CUDA streams: H2D of one chunk, the
 * kernels of the previous one and D2H of the one before overlap (north_star: "streams it H2D on multiple CUDA
 * streams").  Defaults: 32 MiB (16 MiB when only reads travel), 16384 pairs; SWB_CHUNK_MB overrides the first.  Pass pinned host memory
 * (swb_malloc_pinned) for the copies to be asynchronous. */
int  swb_set_chunking(swb_ctx*, uint64_t chunk_bytes, uint64_t min_chunk_pairs);
/* A chunk whose reads (or whose windows) all have the same length uploads no offsets for that side: the device writes
 * k * length itself (SWB_UNIFORM_OFFSETS=0 uploads them regardless). */
/* A ramped batch (four or more chunks) starts with chunks of 1/8, 1/4, 1/2 of that size and ends with 1/2, 1/4 (never
 * below min_chunk_pairs): the first copy and the last chunk's kernels are the parts of the pipeline nothing overlaps
 * with.  ramp = 1 (default, or SWB_CHUNK_RAMP=1): only batches against the resident reference ramp (there the kernels
 * are the long leg; with reads and windows on the wire the copy engine is, and equal chunks keep it busiest);
 * 0 = equal chunks always, 2 = ramp always. */
int  swb_set_chunk_ramp(swb_ctx*, int ramp);

/* Raw device / pinned-host memory and copies for hosts without a CUDA runtime of their own (the CLI, the
 * ctypes tests, a Rust caller).  Replaces ocl::Buffer creation in gpu_align (aligner.rs:466-499);
 * USE_PINNED_MEMORY (aligner.rs:466-475) maps to swb_malloc_pinned. */
int  swb_malloc_device(swb_ctx*, uint64_t bytes, void** out);
int  swb_free_device(swb_ctx*, void* p);
int  swb_malloc_pinned(uint64_t bytes, void** out);
/* Makes the context's device the calling thread's current device.  A helper thread that allocates pinned memory for a
 * context (swb_malloc_pinned page-locks under the CURRENT device's context lock) should call this first: left on the
 * default device 0, the page-locking of every thread of a multi-GPU process stalls device 0's own launches. */
int  swb_bind_thread(swb_ctx*);
int  swb_free_pinned(void* p);
int  swb_memcpy_h2d(swb_ctx*, void* d_dst, const void* h_src, uint64_t bytes);
int  swb_memcpy_d2h(swb_ctx*, void* h_dst, const void* d_src, uint64_t bytes);

/* Host memory placement: make the calling thread PREFER the NUMA node the device hangs off for its next allocations
 * (pinned buffers the GPU reads over PCIe), and back to the default policy.  A preference only: no CPU affinity, nothing
 * fails when the node is not allowed.  Returns the node, or -1 when the platform does not say. */
int  swb_numa_prefer_device(int device_id);
void swb_numa_reset(void);

void*       swb_stream(swb_ctx*);            /* the cudaStream_t the context launches on */
const char* swb_last_error(void);
const char* swb_version(void);

#ifdef __cplusplus
}
#endif
#endif
