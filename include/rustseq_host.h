/*
 * rustseq_host.h -- C++ host mirror of the reference's Rust host functions for the alignment path
 * (the reference is compiled code and no Rust toolchain exists in this environment, so the host side above
 * the C ABI of swb200.h is C++; every function names the Rust item it mirrors).  Same names, argument meaning
 * and error behaviour; Result<T, String> becomes "return 0 / non-zero + rsm_last_error()".
 *
 * SCORE MODE.  The reference's gpu_align returns the value of its live kernel (2 or 0).  This
 * engine's contract is the true Smith-Waterman score of the same scoring function.  The mode is
 * chosen with the environment variable SWB_GPU_ALIGN_MODE:
 *     "sw"          (default)  full Smith-Waterman: best local score (+ end cell via rsm_gpu_align_ex)
 *     "ref_compat"             the literal value gpu_align returns today
 */
#ifndef RUSTSEQ_HOST_H
#define RUSTSEQ_HOST_H
#include <stdint.h>
#include <stddef.h>
#include "swb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* gpu.rs:17-23  pub struct GpuDevice { name, memory_gb, max_work_group_size } (+ the CUDA ordinal) */
typedef struct { char name[256]; float memory_gb; uint64_t max_work_group_size; int32_t ordinal; } rsm_gpu_device;
/* gpu.rs:26-30  pub struct GpuAlignmentResult { score, processing_time_ms, gpu_device } (+ totals the driver prints) */
typedef struct { int32_t score; int64_t score64; double processing_time_ms; char gpu_device[256];
                 uint64_t total_reads; uint64_t total_bases; } rsm_alignment_result;

/* gpu.rs:9-10 */
#define RSM_GPU_WORK_GROUP_SIZE 1024
#define RSM_GPU_MAX_WORK_GROUPS 1000000

int  rsm_is_gpu_available(void);                                   /* gpu.rs:33-45  is_gpu_available() */
int  rsm_get_gpu_devices(rsm_gpu_device* out, int cap);            /* gpu.rs:47-94  get_gpu_devices(); returns the count */

int  rsm_get_chunk_size_reads(uint64_t* out);                      /* aligner.rs:9-15; same two error strings */
int  rsm_get_chunk_size_bases(uint64_t* out);                      /* README.md:32 GPU_CHUNK_SIZE_BASES (0 = unset = no cap) */

/* aligner.rs:107-178  process_fastq_file_in_chunks(filepath, chunk_size_reads, processor).
 * The processor receives the chunk as CSR (what `&[String]` is, flattened): n_reads sequence lines,
 * read k = bases[offsets[k] .. offsets[k+1]).  Non-zero return aborts with that error (the `?`). */
typedef int (*rsm_chunk_fn)(void* user, const uint8_t* bases, const uint64_t* offsets, uint64_t n_reads);
int  rsm_process_fastq_file_in_chunks(const char* filepath, uint64_t chunk_size_reads, rsm_chunk_fn processor, void* user);

int  rsm_count_bases_in_fastq(const char* filepath, uint64_t* out);  /* aligner.rs:535-544 */
/* Test hook, needs no GPU: the BGZF readers of the --full-wgs driver (several pread() threads per file, ordered hand-off of
 * segments) against a consumer that only takes the segments in order.  *hash covers every block's compressed payload and
 * inflated size in stream order: the same for every reader count, segment size and pool size.  *status: 0 ok, 2 not BGZF. */
/* Test hook: a .gz file through the host gzip reader of the FASTQ path (csrc/host_gunzip.h; use_zlib = 1: zlib's gzread, the
 * behaviour it keeps) in read() calls of read_cap bytes.  *n = bytes delivered, *failed = 1 on corrupt data. */
int  rsm_debug_gunzip(const char* path, uint64_t read_cap, int use_zlib, uint8_t* out, uint64_t out_cap, uint64_t* n, int* failed);
/* Test hook: the same file through the parallel reader (csrc/host_pgunzip.h: `threads` decoder threads on chunks of chunk_bytes
 * compressed bytes, 0 = 1 MiB).  *parallel = 0 when the file went to the serial reader; *accepted / *serial_stretches = chunks
 * taken from the decoder threads / stretches the calling thread decoded itself. */
int  rsm_debug_pgunzip(const char* path, uint64_t read_cap, unsigned threads, uint64_t chunk_bytes, uint8_t* out, uint64_t out_cap, uint64_t* n,
                       int* failed, int* parallel, uint64_t* accepted, uint64_t* serial_stretches);
int  rsm_debug_bgzf_segments(const char* path, unsigned readers, uint64_t seg_bytes, unsigned pool_buffers, uint64_t* n_segments,
                             uint64_t* n_blocks, uint64_t* text_bytes, uint64_t* hash, int* status);

/* aligner.rs:410-532  gpu_align(seq1, seq2, device) -> Result<i32, String> */
int  rsm_gpu_align(const uint8_t* seq1, uint64_t n1, const uint8_t* seq2, uint64_t n2, const rsm_gpu_device* device, int32_t* score);
int  rsm_gpu_align_ex(const uint8_t* seq1, uint64_t n1, const uint8_t* seq2, uint64_t n2, const rsm_gpu_device* device, swb_result* out);
/* aligner.rs:365-373  gpu_align_chunk_self(chunk, device): < 1000 bytes -> Ok(0), else gpu_align(chunk, chunk) */
int  rsm_gpu_align_chunk_self(const uint8_t* chunk, uint64_t n, const rsm_gpu_device* device, int32_t* score);
/* aligner.rs:376-407  gpu_align_pair(file1, file2, device) */
int  rsm_gpu_align_pair(const char* file1, const char* file2, const rsm_gpu_device* device, rsm_alignment_result* out);
/* aligner.rs:183-362  process_full_wgs_dataset(device) -> Vec<GpuAlignmentResult> (one per file) */
int  rsm_process_full_wgs_dataset(const rsm_gpu_device* device, rsm_alignment_result* out, int cap, int* n_out);
/* the file list of aligner.rs:197-204: {WGS_DATA_DIR}/{WGS_SAMPLE_ID}_L{lane:03}_R{read}_001.fastq.gz */
int  rsm_wgs_file_list(char* buf, size_t cap, int* n_files);       /* newline-separated paths */

/* aligner.rs:23-104  FileCheckpoint / CheckpointState, serialised like serde_json::to_string_pretty does.  The reference
 * never finds its own checkpoints (load opens checkpoint_{run_id}.json with a fresh run id, save writes
 * checkpoint_run_{N}.json); here both sides use checkpoint_{run_id}.json in the working directory and the run id
 * is WGS_RUN_ID when set (resume), else wgs_{unix time} like aligner.rs:219.  score64 is an extra field (the i32 wraps). */
typedef struct { char file_path[1024]; uint64_t file_index; int32_t score; int64_t score64; double processing_time_ms;
                 uint64_t total_bases; uint64_t total_reads; int32_t completed; } rsm_file_checkpoint;
int  rsm_checkpoint_save(const char* path, const char* run_id, const rsm_file_checkpoint* files, int n_files, uint64_t total_files);
int  rsm_checkpoint_load(const char* path, char* run_id, size_t run_id_cap, rsm_file_checkpoint* files, int cap, int* n_files,
                         uint64_t* total_files);       /* returns 0 and *n_files = -1 when the file does not exist */

/* main.rs:48-192  the CLI (rustseq_mini).  Returns the process exit code. */
int  rsm_main(int argc, char** argv);

const char* rsm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
