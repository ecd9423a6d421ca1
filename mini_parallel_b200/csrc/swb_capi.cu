// swb_capi.cu -- the C ABI of include/swb200.h: context, device arenas, the batch
// pipeline (H2D -> pack -> classify -> short/generic kernels -> D2H) and the compat modes.
// No CPU fallback lives here: without a CUDA device swb_create() fails (main.rs:76-79).
#include "swb_kernels.cuh"
#include <string>
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <chrono>
#include <cctype>
#include <fstream>
#include <unistd.h>
#include <sys/syscall.h>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>

namespace swb {
int launch_generic_single(const BatchView& b, int32_t* last_row_out, cudaStream_t st);
}
#include <array>

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
  return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// SWB_GUARD=1 (debugging aid; compute-sanitizer is not available on every box): every arena is allocated with a 4 KiB
// zone of 0xA5 before and after it and WITHOUT the geometric over-allocation, and swb_debug_guard_check() reports every
// zone a kernel or copy wrote into.  A write past an arena then shows up as a damaged zone instead of as silent
// corruption of the neighbouring allocation; reads of the documented 64 B slack stay legal.
static const bool g_guard = [] { const char* v = std::getenv("SWB_GUARD"); return v && *v && *v != '0'; }();
constexpr size_t kGuardBytes = 4096;
struct GuardEntry { uint8_t* base; size_t cap; int device; };
static std::mutex g_guard_mu;
static std::vector<GuardEntry> g_guard_list;

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
    size_t want = g_guard ? ((bytes + 255) & ~(size_t)255) + 256 : bytes + bytes / 8 + 256;
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, want + (g_guard ? 2 * kGuardBytes : 0));
    if (e != cudaSuccess) { g_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return 1; }
    if (g_guard) {
      int dev = 0; cudaGetDevice(&dev);
      cudaMemset(base, 0xA5, kGuardBytes);
      cudaMemset(static_cast<uint8_t*>(base) + kGuardBytes + want, 0xA5, kGuardBytes);
      cudaDeviceSynchronize();
      std::lock_guard<std::mutex> lk(g_guard_mu);
      g_guard_list.push_back({static_cast<uint8_t*>(base), want, dev});
      base = static_cast<uint8_t*>(base) + kGuardBytes;
    }
    p = base; cap = want; return 0;
  }
  void release() {
    if (p && g_guard) {
      uint8_t* base = static_cast<uint8_t*>(p) - kGuardBytes;
      { std::lock_guard<std::mutex> lk(g_guard_mu);
        for (size_t k = 0; k < g_guard_list.size(); ++k) if (g_guard_list[k].base == base) { g_guard_list.erase(g_guard_list.begin() + k); break; } }
      cudaFree(base);
    } else if (p) cudaFree(p);
    p = nullptr; cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One pipeline lane = one CUDA stream + the arenas of one batch (or one chunk of a host batch) in flight.
struct Lane {
  cudaStream_t st = nullptr;
  cudaEvent_t ev[8] = {};
  DevBuf mid_desc;
  DevBuf q_bytes, r_bytes, q_off, r_off, q_pk, r_pk, q_bad, r_bad, short_list, short_desc, generic_list, long_list, bytes_list, long_scratch, bytes_scratch, counters, out, scratch, misc;
  DevBuf win_beg, win_end, win_len;
  void release_all() {
    for (DevBuf* b : {&q_bytes, &r_bytes, &q_off, &r_off, &q_pk, &r_pk, &q_bad, &r_bad, &short_list, &short_desc,
                      &generic_list, &long_list, &bytes_list, &long_scratch, &bytes_scratch, &counters, &out, &scratch, &misc, &win_beg, &win_end, &win_len, &mid_desc}) b->release();
  }
};

constexpr int kLanes = 6;                  // pipeline lanes a context owns; a host batch uses the first n_lanes of them
                                           // (chunks in flight: H2D / kernels / D2H overlap)

struct ChunkEvents { cudaEvent_t ev[8]; cudaEvent_t copied; };   // ev: stage boundaries (timed); copied: the chunk's host-to-device copies are done

// The packed text the windows of a batch are cut from: the resident reference (swb_set_reference) or the window buffer of
// one swb_score_batch_ranges call.  Window coordinates are relative to `base` (the first byte that was uploaded).
struct RefSrc {
  const uint8_t* bytes = nullptr; const uint32_t* pk = nullptr; const uint32_t* bad = nullptr;
  uint64_t len = 0;            // bytes the caller's coordinates may address
  uint64_t base = 0;           // caller coordinate of bytes[0]
  cudaEvent_t ready = nullptr; // recorded after upload + pack (lanes other than the uploading one wait for it)
};

struct swb_ctx : Lane {                    // lane 0 is the context itself (device-resident API, single-pair calls)
  int device = 0, sm_count = 0;
  Lane extra[kLanes - 1];
  Lane* lane(int i) { return i == 0 ? static_cast<Lane*>(this) : &extra[i - 1]; }
  DevBuf ref_bytes, ref_pk, ref_bad;       // device-resident reference (swb_set_reference)
  DevBuf call_ref_bytes, call_ref_pk, call_ref_bad;   // window buffer of one swb_score_batch_ranges call
  cudaEvent_t call_ref_ready = nullptr;
  swb::LaunchCfg lc;                       // launch state of this context's kernels (no process-wide statics)
  uint64_t last_ranges_uploaded = 0, last_ranges_window_bytes = 0;   // swb_last_ranges_info
  // swb_fastq_bgzf_score / _prefetch: two segment slots (compressed bytes, block table, text) so that the next segment can be
  // copied in and inflated on lane 1's stream while this one is indexed and scored on lane 0's
  struct FqSlot {
    DevBuf comp, blocks, out_off, text, fail;
    std::vector<uint64_t> out_off_host;
    const uint8_t* key = nullptr; uint64_t key_bytes = 0, n_blocks = 0, text_end = 0;
    cudaEvent_t done = nullptr; bool pending = false;
  } fq_slot[2];
  DevBuf fq_tile_count, fq_tile_prefix, fq_seq_beg, fq_seq_end, fq_scal;
  // longest read of the last segment of every file (by file_index): the next segment's bound on the read length, which picks
  // the 128-row instantiation for files of <= 128 bp reads.  A bound only: a longer read is still scored exactly (classify
  // routes it past the kernel), and the following segment runs on 160 rows again.
  std::vector<std::pair<uint64_t, uint32_t>> fq_read_hint;
  DevBuf tb_scratch, tb_res, tb_out, tb_cigar, tb_cursor, tb_handled;   // swb_traceback_batch
  uint64_t ref_len = 0;
  std::vector<ChunkEvents> chunk_ev;       // host path: one event set per chunk of the last call
  swb::Counters* h_counters = nullptr;     // pinned, one slot per chunk
  size_t h_counters_cap = 0;
  size_t last_chunks = 0;
  uint64_t chunk_bytes = 32ull << 20;      // ASCII bytes per chunk of a host batch (SWB_CHUNK_MB, swb_set_chunking); half of it
                                           // when only reads travel (windows of the resident reference): measured optima on B200
  uint64_t min_chunk_pairs = 16384;
  int n_lanes = 3;                         // SWB_LANES
  bool uniform_offsets = true;             // SWB_UNIFORM_OFFSETS=0: always upload the offsets
  int chunk_ramp = 1;                      // SWB_CHUNK_RAMP / swb_set_chunk_ramp: 0 equal chunks, 1 auto, 2 always ramp
  std::vector<uint64_t> chunk_bounds;      // pair index where every chunk of the last host batch starts (+ n_pairs)
  float last_ms[6] = {0, 0, 0, 0, 0, 0};
  int last_kernels = 0;
  uint64_t last_routing[3] = {0, 0, 0};
  bool timings_pending = false, host_path = false;
  int variant = 9;                         // sw_stream_kernel<16,10,4, two-step tracker + dynamic couple distribution>
  bool rows128 = true;                     // SWB_ROWS128=0: reads <= 128 bp also take the 160-row instantiation (comparison)
  bool mid_path = true;                    // SWB_MID_PATH=0: reads of 161..320 bp go to the 32-bit long-pair kernel (tests / comparison)
  uint64_t last_routing5[5] = {0, 0, 0, 0, 0};
  int force_bytes = 0;                     // SWB_FORCE_BYTES=1: every non-short pair through the byte-compare kernel (bench / tests)
};

extern "C" {

const char* swb_last_error(void) { return g_err.c_str(); }
const char* swb_version(void) { return "swb200 0.1 (sm_100a)"; }

int swb_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int swb_device_info(int device_id, char* name, size_t name_cap, double* memory_gb, int* max_wg)
{
  cudaDeviceProp p;
  CUDA_TRY(cudaGetDeviceProperties(&p, device_id));
  if (name && name_cap) { std::strncpy(name, p.name, name_cap - 1); name[name_cap - 1] = 0; }
  if (memory_gb) *memory_gb = (double)p.totalGlobalMem / (1024.0 * 1024.0 * 1024.0);
  if (max_wg) *max_wg = p.maxThreadsPerBlock;
  return 0;
}

int swb_memory_info(int device_id, uint64_t* free_bytes, uint64_t* total_bytes)
{
  CUDA_TRY(cudaSetDevice(device_id));
  size_t f = 0, t = 0;
  CUDA_TRY(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}

int swb_create(swb_ctx** out, int device_id, const swb_params* params)
{
  if (!out) return fail("swb_create: null out pointer");
  *out = nullptr;
  if (params && (params->match != swb::kMatch || params->mismatch != swb::kMismatch || params->gap != swb::kGap))
    return fail("swb_create: only the reference's constants {2,-1,-2} (smith_waterman.cl:5-7) are supported");
  int n = swb_device_count();
  if (n <= 0) return fail("error: gpu acceleration is required and no compatible gpu was found");   // main.rs:161
  if (device_id < 0 || device_id >= n) return fail("swb_create: no such device");
  CUDA_TRY(cudaSetDevice(device_id));
  // SWB_BLOCKING_SYNC=1: host threads sleep in cudaStreamSynchronize instead of spinning (a process that drives eight GPUs and
  // reads 16 files has more busy threads than a 32-core box has cores)
  if (const char* v = std::getenv("SWB_BLOCKING_SYNC")) if (std::atoi(v) != 0) { cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync); cudaGetLastError(); }
  swb_ctx* c = new swb_ctx();
  c->device = device_id;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device_id) != cudaSuccess) { delete c; return fail("cudaGetDeviceProperties failed"); }
  c->sm_count = p.multiProcessorCount;
  swb::launch_cfg_init(c->lc, c->sm_count);
  for (int i = 0; i < kLanes; ++i) {
    Lane* l = c->lane(i);
    if (cudaStreamCreateWithFlags(&l->st, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail("cudaStreamCreate failed"); }
    for (auto& e : l->ev) cudaEventCreate(&e);
  }
  if (const char* v = std::getenv("SWB_SHORT_VARIANT")) { const int w = std::atoi(v) & 15; if (swb::short_variant_available(w)) c->variant = w; }
  if (const char* v = std::getenv("SWB_FORCE_BYTES")) c->force_bytes = std::atoi(v) != 0;
  if (const char* v = std::getenv("SWB_MID_PATH")) c->mid_path = std::atoi(v) != 0;
  if (const char* v = std::getenv("SWB_ROWS128")) c->rows128 = std::atoi(v) != 0;
  if (const char* v = std::getenv("SWB_UNIFORM_OFFSETS")) c->uniform_offsets = std::atoi(v) != 0;
  if (const char* v = std::getenv("SWB_LANES")) c->n_lanes = std::min(kLanes, std::max(1, std::atoi(v)));
  if (const char* v = std::getenv("SWB_CHUNK_RAMP")) c->chunk_ramp = std::min(2, std::max(0, std::atoi(v)));
  if (const char* v = std::getenv("SWB_CHUNK_MB")) { const long mb = std::atol(v); if (mb > 0) c->chunk_bytes = (uint64_t)mb << 20; }
  *out = c;
  return 0;
}

void swb_destroy(swb_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < kLanes; ++i) {
    Lane* l = c->lane(i);
    l->release_all();
    for (auto& e : l->ev) cudaEventDestroy(e);
    cudaStreamDestroy(l->st);
  }
  if (c->call_ref_ready) cudaEventDestroy(c->call_ref_ready);
  for (DevBuf* b : {&c->call_ref_bytes, &c->call_ref_pk, &c->call_ref_bad}) b->release();
  for (DevBuf* b : {&c->ref_bytes, &c->ref_pk, &c->ref_bad, &c->fq_tile_count, &c->fq_tile_prefix, &c->fq_seq_beg, &c->fq_seq_end, &c->fq_scal,
                    &c->tb_scratch, &c->tb_res, &c->tb_out, &c->tb_cigar, &c->tb_cursor, &c->tb_handled}) b->release();
  for (auto& sl : c->fq_slot) {
    for (DevBuf* b : {&sl.comp, &sl.blocks, &sl.out_off, &sl.text, &sl.fail}) b->release();
    if (sl.done) cudaEventDestroy(sl.done);
  }
  for (auto& ce : c->chunk_ev) { for (auto& e : ce.ev) cudaEventDestroy(e); cudaEventDestroy(ce.copied); }
  if (c->h_counters) cudaFreeHost(c->h_counters);
  delete c;
}

// ---- host memory placement ----
// Pinned buffers a GPU reads over PCIe should live on the NUMA node the GPU hangs off: with eight ranks pulling ~50 GB/s
// each, buffers that all sit on one socket make the other socket's GPUs cross the inter-socket link.  This only sets the
// calling thread's allocation PREFERENCE (MPOL_PREFERRED; no CPU affinity, nothing fails if the node is not allowed):
// call it before allocating (swb_malloc_pinned, cudaHostAlloc, a pinned torch tensor), and swb_numa_reset() afterwards.
// Returns the node (>= 0), or -1 when the platform does not say (single node, container without the sysfs entry).
int swb_numa_prefer_device(int device_id)
{
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device_id) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char* p = bus; *p; ++p) *p = (char)std::tolower((unsigned char)*p);
  std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
  int node = -1;
  if (!(f >> node) || node < 0 || node >= 1024) return -1;
  unsigned long mask[16] = {0};
  mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
#ifdef SYS_set_mempolicy
  if (syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, 1024ul) != 0) return -1;
  return node;
#else
  return -1;
#endif
}

void swb_numa_reset(void)
{
#ifdef SYS_set_mempolicy
  syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
}

void* swb_stream(swb_ctx* c) { return c ? (void*)c->st : nullptr; }
int   swb_set_short_variant(swb_ctx* c, int v)
{
  if (!c) return fail("null ctx");
  if (!swb::short_variant_available(v)) return fail("swb_set_short_variant: variant " + std::to_string(v) + " is not in this build (the product library carries the default, 9; the others are in the test build, make variants)");
  c->variant = v;
  return 0;
}

int swb_set_chunking(swb_ctx* c, uint64_t chunk_bytes, uint64_t min_chunk_pairs)
{
  if (!c) return fail("null ctx");
  if (chunk_bytes == 0 || min_chunk_pairs == 0) return fail("swb_set_chunking: sizes must be positive");
  c->chunk_bytes = chunk_bytes; c->min_chunk_pairs = min_chunk_pairs;
  return 0;
}

int swb_set_chunk_ramp(swb_ctx* c, int ramp)
{
  if (!c) return fail("null ctx");
  if (ramp < 0 || ramp > 2) return fail("swb_set_chunk_ramp: 0 (equal chunks), 1 (auto) or 2 (always)");
  c->chunk_ramp = ramp;
  return 0;
}

int swb_sync(swb_ctx* c)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  for (int i = 0; i < kLanes; ++i) CUDA_TRY(cudaStreamSynchronize(c->lane(i)->st));
  if (c->timings_pending && !c->host_path && c->counters.p) {         // a device batch: was the caller's window bound right?
    uint32_t refused = 0;
    CUDA_TRY(cudaMemcpy(&refused, &c->counters.as<swb::Counters>()->n_overflow, 4, cudaMemcpyDeviceToHost));
    if (refused) return fail("swb_score_batch_device: " + std::to_string(refused) + " pairs have a window longer than max_r_len; they were not scored (score INT32_MIN)");
  }
  return 0;
}

// ---- the device pipeline: everything after the inputs are resident in HBM ----
// Runs on lane `l` (its stream, its arenas), events into `ev` (8 of them).  Windows are either CSR ranges of d_r
// (d_rend == d_rbeg + 1, packed here) or ranges of the resident, already packed reference (ref_windows).
static int run_device_pipeline(swb_ctx* c, Lane* l, cudaEvent_t* ev, const uint8_t* d_q, const uint64_t* d_qo, uint64_t q_total,
                               const uint8_t* d_r, const uint64_t* d_rbeg, const uint64_t* d_rend, uint64_t r_total,
                               const RefSrc* ref /* windows are ranges of this packed text; nullptr: CSR windows in d_r */, uint64_t n_pairs, uint32_t max_q_len, uint32_t max_r_len, swb_result* d_out, int* kernels,
                               const uint64_t* d_qend = nullptr /* reads as [beg,end) ranges instead of CSR */,
                               bool q_prepacked = false /* lane's q_pk / q_bad already hold the packed d_q */)
{
  if (n_pairs >= (1ull << 32)) return fail("swb: at most 2^32-1 pairs per batch");
  const bool ref_windows = ref != nullptr;
  const uint64_t qw = (q_total + 15) / 16, rw = ref_windows ? 0 : (r_total + 15) / 16;
  if (l->q_pk.reserve(qw * 4 + 64) || l->r_pk.reserve(rw * 4 + 64) ||
      l->q_bad.reserve((qw + 31) / 32 * 4 + 64) || l->r_bad.reserve((rw + 31) / 32 * 4 + 64) ||
      l->short_list.reserve(n_pairs * 4 + 64) || l->short_desc.reserve(n_pairs * sizeof(swb::ShortDesc) + 64) ||
      (max_q_len > swb::kShortMaxRead && l->mid_desc.reserve(n_pairs * sizeof(swb::ShortDesc) + 64)) ||
      l->generic_list.reserve(n_pairs * 4 + 64) || l->long_list.reserve(n_pairs * 4 + 64) || l->bytes_list.reserve(n_pairs * 4 + 64) ||
      l->counters.reserve(sizeof(swb::Counters))) return 1;
  // generic kernel: persistent grid, one boundary row of max_r_len ints per resident warp
  int ctas = c->sm_count * 4;
  const uint64_t stride = ((uint64_t)max_r_len + 32) & ~31ull;
  while (ctas > 1 && (uint64_t)ctas * 4 * stride * 4 > (2ull << 30)) ctas /= 2;
  if (l->scratch.reserve((uint64_t)ctas * 4 * stride * 4) || l->long_scratch.reserve((uint64_t)ctas * 4 * stride * 4) ||
      l->bytes_scratch.reserve((uint64_t)ctas * 4 * stride * 4)) return 1;

  swb::BatchView b;
  b.q_bytes = d_q; b.q_beg = d_qo; b.q_end = d_qend ? d_qend : d_qo + 1; b.r_bytes = d_r; b.r_beg = d_rbeg; b.r_end = d_rend;
  b.q_pk = l->q_pk.as<uint32_t>(); b.q_bad = l->q_bad.as<uint32_t>();
  b.r_pk = ref_windows ? ref->pk : l->r_pk.as<uint32_t>();
  b.r_bad = ref_windows ? ref->bad : l->r_bad.as<uint32_t>();
  b.n_pairs = n_pairs;
  b.short_list = l->short_list.as<uint32_t>(); b.short_desc = l->short_desc.as<swb::ShortDesc>();
  b.mid_desc = (max_q_len > swb::kShortMaxRead && c->mid_path) ? l->mid_desc.as<swb::ShortDesc>() : nullptr;   // no read beyond 160 bp: no mid list, no launch
  b.generic_list = l->generic_list.as<uint32_t>(); b.long_list = l->long_list.as<uint32_t>(); b.bytes_list = l->bytes_list.as<uint32_t>();
  b.force_bytes = c->force_bytes;
  // max_q_len is exact on the host paths and the caller's bound on the device path; a read longer than it is still scored
  // (classify_kernel routes it to a kernel without the row limit), a wrong bound only costs speed
  b.short_max_read = (max_q_len >= 1 && max_q_len <= 128 && c->variant == 9 && c->rows128) ? 128u : swb::kShortMaxRead;   // 0 = unknown
  b.counters = l->counters.as<swb::Counters>(); b.out = d_out;
  b.scratch = l->scratch.as<int32_t>(); b.scratch_stride = stride; b.max_window = max_r_len;

  int k = 0;
  cudaStream_t st = l->st;
  CUDA_TRY(cudaMemsetAsync(l->counters.p, 0, sizeof(swb::Counters), st));
  CUDA_TRY(cudaEventRecord(ev[0], st));
  if (!q_prepacked) k += swb::launch_pack2bit(d_q, q_total, l->q_pk.as<uint32_t>(), l->q_bad.as<uint32_t>(), st);
  if (!ref_windows) k += swb::launch_pack2bit(d_r, r_total, l->r_pk.as<uint32_t>(), l->r_bad.as<uint32_t>(), st);
  if (ref_windows && ref->ready) CUDA_TRY(cudaStreamWaitEvent(st, ref->ready, 0));
  k += swb::launch_classify(b, st);
  CUDA_TRY(cudaEventRecord(ev[1], st));
  const uint32_t wcap = std::min<uint32_t>(std::max<uint32_t>(max_r_len, 1), swb::kShortMaxWindow);
  k += swb::launch_short(b, wcap, c->variant, c->lc, st);
  k += swb::launch_mid(b, c->lc, st);
  CUDA_TRY(cudaEventRecord(ev[2], st));
  k += swb::launch_generic(b, ctas / 4 > 0 ? ctas / 4 : 1, 0, st);
  {
    swb::BatchView bl = b;
    bl.scratch = l->long_scratch.as<int32_t>();
    k += swb::launch_long(bl, ctas, max_q_len, c->lc, st);
    bl.scratch = l->bytes_scratch.as<int32_t>();
    k += swb::launch_long_bytes(bl, ctas, max_q_len, c->lc, st);
  }
  CUDA_TRY(cudaEventRecord(ev[3], st));
  CUDA_TRY(cudaGetLastError());
  *kernels += k;
  return 0;
}

int swb_score_batch_device(swb_ctx* c, const uint8_t* d_q, const uint64_t* d_qo, uint64_t q_total,
                           const uint8_t* d_r, const uint64_t* d_ro, uint64_t r_total,
                           uint64_t n_pairs, uint32_t max_q_len, uint32_t max_r_len, swb_result* d_out)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  c->host_path = false;
  c->last_kernels = 0; c->last_chunks = 0;
  if (n_pairs == 0) { c->timings_pending = false; return 0; }
  if (run_device_pipeline(c, c, c->ev, d_q, d_qo, q_total, d_r, d_ro, d_ro + 1, r_total, nullptr, n_pairs,
                          max_q_len ? max_q_len : 0xffffffffu, max_r_len, d_out, &c->last_kernels)) return 1;
  c->timings_pending = true;
  return 0;
}

// ---- host batches: chunks of pairs pipelined over kLanes streams ----
// Chunk k runs on lane k % n_lanes: H2D of its bytes and offsets, offset rebase, the device pipeline, D2H of its
// results.  While one lane computes, the next lane's H2D and the previous lane's D2H are in flight (one copy
// engine per direction), so a large batch costs max(PCIe, kernels) instead of their sum.
static int ensure_chunk_slots(swb_ctx* c, size_t n_chunks)
{
  while (c->chunk_ev.size() < n_chunks) {
    ChunkEvents ce;
    for (auto& e : ce.ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventCreateWithFlags(&ce.copied, cudaEventDisableTiming));
    c->chunk_ev.push_back(ce);
  }
  if (c->h_counters_cap < n_chunks) {
    if (c->h_counters) cudaFreeHost(c->h_counters);
    c->h_counters = nullptr; c->h_counters_cap = 0;
    CUDA_TRY(cudaHostAlloc((void**)&c->h_counters, sizeof(swb::Counters) * (n_chunks + 16), cudaHostAllocDefault));
    c->h_counters_cap = n_chunks + 16;
  }
  return 0;
}

static RefSrc resident_ref(swb_ctx* c)
{
  RefSrc r;
  r.bytes = c->ref_bytes.as<uint8_t>(); r.pk = c->ref_pk.as<uint32_t>(); r.bad = c->ref_bad.as<uint32_t>(); r.len = c->ref_len;
  return r;
}

// q / r are the bases of the caller's byte arrays and qo / ro hold ABSOLUTE offsets into them, so a slice of a larger batch
// (swb_multi_score_batch: pairs [lo, hi) of the whole) is scored by passing qo + lo, ro + lo, out + lo unchanged.
static int score_host_batch_enqueue(swb_ctx* c, const char* who, const uint8_t* q, const uint64_t* qo,
                                   const uint8_t* r, const uint64_t* ro,                     /* CSR windows, or */
                                   const uint64_t* win_start, const uint32_t* win_len,       /* windows of `ref` */
                                   const RefSrc* ref, uint64_t n_pairs, swb_result* out)
{
  const bool ref_windows = (r == nullptr && ro == nullptr);
  const uint64_t bytes_total = (qo[n_pairs] - qo[0]) + (ref_windows ? 0 : ro[n_pairs] - ro[0]);
  const uint64_t cb = ref_windows ? std::max<uint64_t>(c->chunk_bytes / 2, 1) : c->chunk_bytes;
  uint64_t n_uniform = std::max<uint64_t>(1, (bytes_total + cb - 1) / cb);
  uint64_t per = (n_pairs + n_uniform - 1) / n_uniform;
  per = std::max<uint64_t>(per, std::min<uint64_t>(n_pairs, c->min_chunk_pairs));
  n_uniform = (n_pairs + per - 1) / per;
  // Chunk schedule (pairs per chunk).  When only reads travel (windows of the resident reference) the kernels are the long
  // leg of the pipeline, and the first H2D and the last chunk's kernels + D2H are the parts nothing overlaps with: such a
  // batch ramps up from per/8 and down again to per/4, the steady chunks stay large enough for the persistent kernels.
  // With reads AND windows on the wire the copy engine is the long leg and equal chunks keep it busiest (measured:
  // 13.4 vs 13.6 ms per 1 M pairs; reference windows 10.9 -> 10.4 ms with the ramp).  chunk_ramp: 0 never, 1 auto, 2 always.
  std::vector<uint64_t>& bounds = c->chunk_bounds;
  bounds.clear(); bounds.push_back(0);
  if ((c->chunk_ramp == 2 || (c->chunk_ramp == 1 && ref_windows)) && n_uniform >= 4) {
    const uint64_t lo = std::max<uint64_t>(1, std::min<uint64_t>(per, c->min_chunk_pairs));
    const uint64_t head[3] = {std::max(per / 8, lo), std::max(per / 4, lo), std::max(per / 2, lo)};
    const uint64_t tail[2] = {std::max(per / 2, lo), std::max(per / 4, lo)};
    const uint64_t tail_sum = tail[0] + tail[1];
    uint64_t p = 0;
    for (int k = 0; k < 3 && p + head[k] + tail_sum < n_pairs; ++k) { p += head[k]; bounds.push_back(p); }
    while (p + per + tail_sum < n_pairs) { p += per; bounds.push_back(p); }
    const uint64_t rest = n_pairs - p;                     // 0 < rest <= per + tail_sum: three pieces 4 : 2 : 1
    const uint64_t a = rest * 4 / 7, b = rest * 2 / 7;
    if (a) { p += a; bounds.push_back(p); }
    if (b) { p += b; bounds.push_back(p); }
    if (p == n_pairs) bounds.pop_back();
    bounds.push_back(n_pairs);
  } else {
    for (uint64_t p = per; p < n_pairs; p += per) bounds.push_back(p);
    bounds.push_back(n_pairs);
  }
  const uint64_t n_chunks = bounds.size() - 1;
  if (ensure_chunk_slots(c, n_chunks)) return 1;
  c->last_kernels = 0;
  uint64_t win_bytes = 0;

  for (uint64_t ch = 0; ch < n_chunks; ++ch) {
    const uint64_t p0 = bounds[ch], p1 = bounds[ch + 1], n = p1 - p0;
    Lane* l = c->lane((int)(ch % (uint64_t)c->n_lanes));
    cudaEvent_t* ev = c->chunk_ev[ch].ev;
    cudaStream_t st = l->st;
    uint32_t max_r = 0, max_q = 0;
    // A side whose sequences all have the same length (150 bp reads, fixed windows: the usual FASTQ chunk) sends no
    // offsets: the device writes k * length itself -- two of the chunk's four copies and 16 bytes per pair stay home.
    const uint64_t q_len0 = qo[p0 + 1] - qo[p0], r_len0 = ref_windows ? 0 : ro[p0 + 1] - ro[p0];
    bool q_uniform = q_len0 != 0 && c->uniform_offsets, r_uniform = !ref_windows && r_len0 != 0 && c->uniform_offsets;
    const uint32_t w_len0 = ref_windows ? win_len[p0] : 0;
    bool w_uniform = ref_windows && w_len0 != 0 && c->uniform_offsets;
    for (uint64_t k = p0; k < p1; ++k) {                      // validation + longest read / window of this chunk
      if (qo[k + 1] < qo[k]) return fail(std::string(who) + ": offsets must be non-decreasing");
      if (qo[k + 1] - qo[k] > 0x7fffffffull) return fail("Sequence too large (more than 2^31-1 bytes)");
      max_q = std::max<uint32_t>(max_q, (uint32_t)(qo[k + 1] - qo[k]));
      q_uniform &= qo[k + 1] - qo[k] == q_len0;
      if (ref_windows) {
        if (win_start[k] > ref->len || win_len[k] > ref->len - win_start[k])
          return fail(std::string(who) + (ref->ready ? ": window outside the buffer" : ": window outside the reference"));
        max_r = std::max<uint32_t>(max_r, win_len[k]);
        w_uniform &= win_len[k] == w_len0;
        win_bytes += win_len[k];
      } else {
        if (ro[k + 1] < ro[k]) return fail(std::string(who) + ": offsets must be non-decreasing");
        if (ro[k + 1] - ro[k] > 0x7fffffffull) return fail("Sequence too large (more than 2^31-1 bytes)");
        max_r = std::max<uint32_t>(max_r, (uint32_t)(ro[k + 1] - ro[k]));
        r_uniform &= ro[k + 1] - ro[k] == r_len0;
      }
    }
    const uint64_t qb = qo[p1] - qo[p0], rb = ref_windows ? 0 : ro[p1] - ro[p0];
    if (l->q_bytes.reserve(qb + 64) || l->q_off.reserve((n + 1) * 8) || l->out.reserve(n * sizeof(swb_result))) return 1;
    if (ref_windows) { if (l->win_beg.reserve(n * 8) || l->win_end.reserve(n * 8) || l->win_len.reserve(n * 4)) return 1; }
    else             { if (l->r_bytes.reserve(rb + 64) || l->r_off.reserve((n + 1) * 8)) return 1; }

    // One chunk's copies at a time, in chunk order.  Copies of different streams otherwise share the link and finish
    // together, their kernels then start together, and the lanes march in step -- copy, compute, copy, compute, nothing
    // overlapping (measured: 12.0 instead of 10.5 ms per million pairs when the three lanes start on one event).
    if (ch > 0) CUDA_TRY(cudaStreamWaitEvent(st, c->chunk_ev[ch - 1].copied, 0));
    CUDA_TRY(cudaEventRecord(ev[4], st));
    if (qb) CUDA_TRY(cudaMemcpyAsync(l->q_bytes.p, q + qo[p0], qb, cudaMemcpyHostToDevice, st));
    if (!q_uniform) CUDA_TRY(cudaMemcpyAsync(l->q_off.p, qo + p0, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    int k = 0;
    if (ref_windows) {
      CUDA_TRY(cudaMemcpyAsync(l->win_beg.p, win_start + p0, n * 8, cudaMemcpyHostToDevice, st));
      if (!w_uniform) CUDA_TRY(cudaMemcpyAsync(l->win_len.p, win_len + p0, n * 4, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaEventRecord(c->chunk_ev[ch].copied, st));
      k += swb::launch_chunk_prepare(l->q_off.as<uint64_t>(), n + 1, qo[p0], q_uniform ? q_len0 : 0, nullptr, 0, 0, 0,
                                     l->win_beg.as<uint64_t>(), l->win_len.as<uint32_t>(), w_uniform ? w_len0 : 0, l->win_end.as<uint64_t>(), n, ref->base, st);
    } else {
      if (rb) CUDA_TRY(cudaMemcpyAsync(l->r_bytes.p, r + ro[p0], rb, cudaMemcpyHostToDevice, st));
      if (!r_uniform) CUDA_TRY(cudaMemcpyAsync(l->r_off.p, ro + p0, (n + 1) * 8, cudaMemcpyHostToDevice, st));
      CUDA_TRY(cudaEventRecord(c->chunk_ev[ch].copied, st));
      k += swb::launch_chunk_prepare(l->q_off.as<uint64_t>(), n + 1, qo[p0], q_uniform ? q_len0 : 0, l->r_off.as<uint64_t>(), n + 1, ro[p0],
                                     r_uniform ? r_len0 : 0, nullptr, nullptr, 0, nullptr, 0, 0, st);
    }
    CUDA_TRY(cudaEventRecord(ev[5], st));
    c->last_kernels += k;
    if (ref_windows) {
      if (run_device_pipeline(c, l, ev, l->q_bytes.as<uint8_t>(), l->q_off.as<uint64_t>(), qb, ref->bytes,
                              l->win_beg.as<uint64_t>(), l->win_end.as<uint64_t>(), 0, ref, n, max_q, max_r, l->out.as<swb_result>(),
                              &c->last_kernels)) return 1;
    } else {
      if (run_device_pipeline(c, l, ev, l->q_bytes.as<uint8_t>(), l->q_off.as<uint64_t>(), qb, l->r_bytes.as<uint8_t>(),
                              l->r_off.as<uint64_t>(), l->r_off.as<uint64_t>() + 1, rb, nullptr, n, max_q, max_r, l->out.as<swb_result>(),
                              &c->last_kernels)) return 1;
    }
    CUDA_TRY(cudaEventRecord(ev[6], st));
    CUDA_TRY(cudaMemcpyAsync(out + p0, l->out.p, n * sizeof(swb_result), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(c->h_counters + ch, l->counters.p, sizeof(swb::Counters), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(ev[7], st));
  }
  for (int i = 0; i < kLanes; ++i) CUDA_TRY(cudaStreamSynchronize(c->lane(i)->st));
  CUDA_TRY(cudaGetLastError());
  c->host_path = true; c->last_chunks = (size_t)n_chunks; c->timings_pending = true;
  c->last_ranges_window_bytes = win_bytes;
  return 0;
}

static int score_host_batch(swb_ctx* c, const char* who, const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                            const uint64_t* win_start, const uint32_t* win_len, const RefSrc* ref, uint64_t n_pairs, swb_result* out)
{
  const int rc = score_host_batch_enqueue(c, who, q, qo, r, ro, win_start, win_len, ref, n_pairs, out);
  if (rc != 0) {
    // an error after some chunks were enqueued: the caller's buffers are borrowed for the call only, so nothing may
    // still be reading or writing them when we return
    const std::string keep = g_err;
    for (int i = 0; i < kLanes; ++i) cudaStreamSynchronize(c->lane(i)->st);
    cudaGetLastError();
    c->timings_pending = false;
    g_err = keep;
  }
  return rc;
}

int swb_score_batch(swb_ctx* c, const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                    uint64_t n_pairs, swb_result* out)
{
  if (!c) return fail("null ctx");
  if (n_pairs == 0) return 0;
  if (!qo || !ro || !out) return fail("swb_score_batch: null pointer");
  if (qo[0] != 0 || ro[0] != 0) return fail("swb_score_batch: offsets must start at 0");
  CUDA_TRY(cudaSetDevice(c->device));
  static const uint8_t empty = 0;
  return score_host_batch(c, "swb_score_batch", q ? q : &empty, qo, r ? r : &empty, ro, nullptr, nullptr, nullptr, n_pairs, out);
}

// ---- reads against windows of a device-resident reference ----
int swb_set_reference(swb_ctx* c, const uint8_t* ref, uint64_t n)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t nw = (n + 15) / 16;
  if (c->ref_bytes.reserve(n + 64) || c->ref_pk.reserve(nw * 4 + 64) || c->ref_bad.reserve((nw + 31) / 32 * 4 + 64)) return 1;
  cudaStream_t st = c->st;
  if (n) CUDA_TRY(cudaMemcpyAsync(c->ref_bytes.p, ref, n, cudaMemcpyHostToDevice, st));
  swb::launch_pack2bit(c->ref_bytes.as<uint8_t>(), n, c->ref_pk.as<uint32_t>(), c->ref_bad.as<uint32_t>(), st);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  c->ref_len = n;
  return 0;
}

int swb_score_batch_vs_reference(swb_ctx* c, const uint8_t* q, const uint64_t* qo, uint64_t n_pairs,
                                 const uint64_t* win_start, const uint32_t* win_len, swb_result* out)
{
  if (!c) return fail("null ctx");
  if (n_pairs == 0) return 0;
  if (!qo || !win_start || !win_len || !out) return fail("swb_score_batch_vs_reference: null pointer");
  if (qo[0] != 0) return fail("swb_score_batch_vs_reference: offsets must start at 0");
  CUDA_TRY(cudaSetDevice(c->device));
  static const uint8_t empty = 0;
  const RefSrc ref = resident_ref(c);
  return score_host_batch(c, "swb_score_batch_vs_reference", q ? q : &empty, qo, nullptr, nullptr, win_start, win_len, &ref, n_pairs, out);
}

// ---- reads against windows that are RANGES of one host buffer (they may overlap, repeat, come in any order) ----
// What a read mapper hands over: candidate windows of a genome, many of them sharing bases.  The covered part of the buffer
// crosses PCIe once per call -- not once per window -- is packed once, and the reads are pipelined against it like against
// the resident reference.  1 M windows of 500 bp in a 16 Mbase buffer: 16 MB on the wire instead of 500 MB.
int swb_score_batch_ranges(swb_ctx* c, const uint8_t* q, const uint64_t* qo, uint64_t n_pairs,
                           const uint8_t* w_bytes, uint64_t w_total, const uint64_t* win_start, const uint32_t* win_len, swb_result* out)
{
  if (!c) return fail("null ctx");
  if (n_pairs == 0) return 0;
  if (!qo || !win_start || !win_len || !out || (w_total && !w_bytes)) return fail("swb_score_batch_ranges: null pointer");
  if (qo[0] != 0) return fail("swb_score_batch_ranges: offsets must start at 0");
  CUDA_TRY(cudaSetDevice(c->device));
  // Which part of the buffer do the windows touch?  Finding out is a pass over the coordinates on the host with the GPU
  // idle, ~1 ns per pair -- the time in which PCIe moves ~64 bytes -- so a buffer of fewer than 64 bytes per pair goes up
  // whole; the windows are still validated chunk by chunk while the GPU works.
  uint64_t lo = 0, hi = w_total, sum = 0;
  if (w_total / 64 > n_pairs) {
    lo = w_total; hi = 0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
      const uint64_t s = win_start[k], n = win_len[k];
      if (s > w_total || n > w_total - s) return fail("swb_score_batch_ranges: window outside the buffer");
      if (n) { lo = std::min(lo, s); hi = std::max(hi, s + n); sum += n; }
    }
    if (hi <= lo) { lo = 0; hi = 0; }
  }
  lo &= ~511ull;                                          // one bitmap word = 32 packed words = 512 bases
  const uint64_t n = hi - lo, nw = (n + 15) / 16;
  if (c->call_ref_bytes.reserve(n + 64) || c->call_ref_pk.reserve(nw * 4 + 64) || c->call_ref_bad.reserve((nw + 31) / 32 * 4 + 64)) return 1;
  if (!c->call_ref_ready) CUDA_TRY(cudaEventCreateWithFlags(&c->call_ref_ready, cudaEventDisableTiming));
  cudaStream_t st = c->st;                                // lane 0: the first chunk follows on the same stream
  if (n) CUDA_TRY(cudaMemcpyAsync(c->call_ref_bytes.p, w_bytes + lo, n, cudaMemcpyHostToDevice, st));
  swb::launch_pack2bit(c->call_ref_bytes.as<uint8_t>(), n, c->call_ref_pk.as<uint32_t>(), c->call_ref_bad.as<uint32_t>(), st);
  CUDA_TRY(cudaEventRecord(c->call_ref_ready, st));
  RefSrc ref;
  ref.bytes = c->call_ref_bytes.as<uint8_t>(); ref.pk = c->call_ref_pk.as<uint32_t>(); ref.bad = c->call_ref_bad.as<uint32_t>();
  ref.len = w_total; ref.base = lo; ref.ready = c->call_ref_ready;
  static const uint8_t empty = 0;
  const int rc = score_host_batch(c, "swb_score_batch_ranges", q ? q : &empty, qo, nullptr, nullptr, win_start, win_len, &ref, n_pairs, out);
  if (rc == 0) c->last_kernels += n ? 1 : 0;
  c->last_ranges_uploaded = n;                            // (the window bytes were summed while the chunks were validated)
  return rc;
}

int swb_last_ranges_info(swb_ctx* c, uint64_t* bytes_uploaded, uint64_t* window_bytes)
{
  if (!c) return fail("null ctx");
  if (bytes_uploaded) *bytes_uploaded = c->last_ranges_uploaded;
  if (window_bytes) *window_bytes = c->last_ranges_window_bytes;
  return 0;
}

// ---- FASTQ.gz (BGZF) -> scores entirely on the device ----
constexpr uint64_t kFqCarryRoom = 1ull << 20;            // text layout: [carry, right-aligned in 1 MiB][inflated blocks]

// copy one segment to the device and inflate it on `st` into slot `s` (asynchronous; s.done fires when the text is there)
static int fq_enqueue_inflate(swb_ctx* c, swb_ctx::FqSlot& s, cudaStream_t st, const uint8_t* comp, uint64_t comp_bytes,
                              const swb_bgzf_block* blocks, uint64_t n_blocks, int* kernels)
{
  s.out_off_host.resize(n_blocks);
  uint64_t text_end = kFqCarryRoom;
  for (uint64_t k = 0; k < n_blocks; ++k) {
    if (blocks[k].in_off + blocks[k].in_len > comp_bytes) return fail("swb_fastq_bgzf: block outside the compressed buffer");
    s.out_off_host[k] = text_end; text_end += blocks[k].out_len;
  }
  if (s.comp.reserve(comp_bytes + 64) || s.blocks.reserve(n_blocks * sizeof(swb_bgzf_block) + 64) || s.out_off.reserve(n_blocks * 8 + 64) ||
      s.text.reserve(text_end + 4096 + 64) || s.fail.reserve(64)) return 1;
  if (!s.done) CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  CUDA_TRY(cudaMemsetAsync(s.fail.p, 0, 64, st));
  if (comp_bytes) CUDA_TRY(cudaMemcpyAsync(s.comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, st));
  if (n_blocks) {
    CUDA_TRY(cudaMemcpyAsync(s.blocks.p, blocks, n_blocks * sizeof(swb_bgzf_block), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s.out_off.p, s.out_off_host.data(), n_blocks * 8, cudaMemcpyHostToDevice, st));
  }
  *kernels += swb::launch_inflate_bgzf(s.comp.as<uint8_t>(), s.blocks.as<swb_bgzf_block>(), n_blocks, s.out_off.as<uint64_t>(), s.text.as<uint8_t>(),
                                       s.fail.as<uint32_t>(), st);
  CUDA_TRY(cudaEventRecord(s.done, st));
  s.key = comp; s.key_bytes = comp_bytes; s.n_blocks = n_blocks; s.text_end = text_end;
  return 0;
}

int swb_fastq_bgzf_prefetch(swb_ctx* c, const uint8_t* comp, uint64_t comp_bytes, const swb_bgzf_block* blocks, uint64_t n_blocks)
{
  if (!c) return fail("null ctx");
  if (n_blocks && (!comp || !blocks)) return fail("swb_fastq_bgzf_prefetch: null input");
  CUDA_TRY(cudaSetDevice(c->device));
  for (auto& s : c->fq_slot) if (s.pending && s.key == comp && s.key_bytes == comp_bytes) return 0;      // already on its way
  for (auto& s : c->fq_slot) {
    if (s.pending) continue;
    int k = 0;
    if (fq_enqueue_inflate(c, s, c->lane(1)->st, comp, comp_bytes, blocks, n_blocks, &k)) return 1;
    s.pending = true;
    return 0;
  }
  return 0;                                               // both slots busy: the segment is inflated when it is scored
}

// a segment that was prefetched but will not be scored (the file fell back to the host reader, the run was aborted): wait
// for its copy and release the slot, so that the pinned buffer can be reused and no later segment at the same address is
// taken for this one
int swb_fastq_bgzf_cancel(swb_ctx* c, const uint8_t* comp)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  for (auto& s : c->fq_slot)
    if (s.pending && s.key == comp) {
      CUDA_TRY(cudaStreamSynchronize(c->lane(1)->st));
      s.pending = false; s.key = nullptr;
    }
  return 0;
}

// nothing may still be reading the caller's buffers when a call returns with an error
static int drain_after_error(swb_ctx* c, int rc)
{
  if (rc != 0 && c) {
    const std::string keep = g_err;
    for (int i = 0; i < kLanes; ++i) cudaStreamSynchronize(c->lane(i)->st);
    cudaGetLastError();
    for (auto& s : c->fq_slot) { s.pending = false; s.key = nullptr; }
    g_err = keep;
  }
  return rc;
}

static int fastq_bgzf_score_impl(swb_ctx* c, const uint8_t* comp, uint64_t comp_bytes, const swb_bgzf_block* blocks, uint64_t n_blocks,
                         const uint8_t* carry, uint64_t carry_len, int final_segment,
                         uint64_t file_index, uint64_t first_read, uint32_t window_len,
                         int64_t* score_sum, uint64_t* n_reads, uint64_t* n_bases, uint64_t* n_lines,
                         uint8_t* carry_out, uint64_t carry_cap, uint64_t* carry_out_len, int* status)
{
  if (!c || !score_sum || !n_reads || !n_bases || !n_lines || !carry_out_len || !status) return fail("swb_fastq_bgzf_score: null pointer");
  if ((n_blocks && (!comp || !blocks)) || (carry_len && !carry)) return fail("swb_fastq_bgzf_score: null input");
  if (c->ref_len == 0) return fail("swb_fastq_bgzf_score: no resident reference (swb_set_reference)");
  if (window_len == 0 || window_len > c->ref_len) return fail("swb_fastq_bgzf_score: window outside the reference");
  if (carry_len > kFqCarryRoom) return fail("swb_fastq_bgzf_score: carry larger than 1 MiB");
  CUDA_TRY(cudaSetDevice(c->device));
  *score_sum = 0; *n_reads = 0; *n_bases = 0; *n_lines = 0; *carry_out_len = 0; *status = 0;
  cudaStream_t st = c->st;
  const auto wall0 = std::chrono::steady_clock::now();
  int k = 0;

  // the segment is either already being inflated (swb_fastq_bgzf_prefetch) or goes into a free slot now
  swb_ctx::FqSlot* slot = nullptr;
  for (auto& s : c->fq_slot) if (s.pending && s.key == comp && s.key_bytes == comp_bytes && s.n_blocks == n_blocks) slot = &s;
  if (!slot) {
    for (auto& s : c->fq_slot) if (!s.pending) { slot = &s; break; }
    if (!slot) {                                          // two stale prefetches: wait for them and reuse the first slot
      CUDA_TRY(cudaStreamSynchronize(c->lane(1)->st));
      for (auto& s : c->fq_slot) s.pending = false;
      slot = &c->fq_slot[0];
    }
    if (fq_enqueue_inflate(c, *slot, st, comp, comp_bytes, blocks, n_blocks, &k)) return 1;
  }
  struct Release { swb_ctx::FqSlot* s; ~Release() { s->pending = false; s->key = nullptr; } } release{slot};
  CUDA_TRY(cudaStreamWaitEvent(st, slot->done, 0));
  const uint64_t begin = kFqCarryRoom - carry_len, end = slot->text_end;
  if (end == begin) { CUDA_TRY(cudaStreamSynchronize(st)); return 0; }
  const uint64_t n_tiles = swb::fq_tiles(begin, end);
  if (c->fq_tile_count.reserve(n_tiles * 4 + 64) || c->fq_tile_prefix.reserve(n_tiles * 8 + 64) || c->fq_scal.reserve(64)) return 1;
  uint8_t* d_text = slot->text.as<uint8_t>();
  uint64_t* d_scal = c->fq_scal.as<uint64_t>();          // [0] newlines [1] tail_start [2] score sum [3] bases [4] (u32) -, flags
  uint32_t* d_fail = reinterpret_cast<uint32_t*>(d_scal + 4);
  CUDA_TRY(cudaMemsetAsync(d_scal, 0, 64, st));
  if (carry_len) CUDA_TRY(cudaMemcpyAsync(d_text + begin, carry, carry_len, cudaMemcpyHostToDevice, st));
  const bool dbg = std::getenv("SWB_DEBUG") != nullptr;
  cudaEvent_t te[6] = {};
  if (dbg) for (auto& e : te) cudaEventCreate(&e);
  if (dbg) cudaEventRecord(te[0], st);
  if (dbg) cudaEventRecord(te[1], st);
  k += swb::launch_fq_index(d_text, begin, end, c->fq_tile_count.as<uint32_t>(), c->fq_tile_prefix.as<uint64_t>(), d_scal, d_fail + 1, final_segment, st);
  if (dbg) cudaEventRecord(te[2], st);
  uint64_t h_scal[8] = {};
  uint32_t h_fail[4] = {};
  uint8_t last = 0;
  CUDA_TRY(cudaMemcpyAsync(h_scal, d_scal, 64, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(h_fail, slot->fail.p, 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(&last, d_text + end - 1, 1, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));                    // the counts size what follows
  CUDA_TRY(cudaGetLastError());
  const uint32_t failed = h_fail[0], flags = (uint32_t)(h_scal[4] >> 32);
  h_scal[5] = (uint64_t)h_fail[2] | ((uint64_t)h_fail[3] << 32);
  if (std::getenv("SWB_DEBUG")) std::fprintf(stderr, "[fastq] blocks %llu text [%llu,%llu) newlines %llu failed %u (first status %u after %u bytes) flags %u last %d\n", (unsigned long long)n_blocks,
                                            (unsigned long long)begin, (unsigned long long)end, (unsigned long long)h_scal[0], failed,
                                            (uint32_t)(h_scal[5] & 0xffffffffu), (uint32_t)(h_scal[5] >> 32), flags, (int)last);
  if (failed || flags) { *status = 1; return 0; }
  const uint64_t newlines = h_scal[0];
  const uint64_t lines = newlines + ((final_segment && last != '\n') ? 1 : 0);
  const uint64_t R = final_segment ? (lines + 2) / 4 : newlines / 4;
  if (c->fq_seq_beg.reserve((R + 1) * 8 + 64) || c->fq_seq_end.reserve((R + 1) * 8 + 64)) return 1;
  uint64_t* d_beg = c->fq_seq_beg.as<uint64_t>(); uint64_t* d_end = c->fq_seq_end.as<uint64_t>();
  CUDA_TRY(cudaMemsetAsync(d_beg, 0, (R + 1) * 8, st));
  CUDA_TRY(cudaMemsetAsync(d_end, 0, (R + 1) * 8, st));
  const unsigned long long tail0 = begin;
  CUDA_TRY(cudaMemcpyAsync(d_scal + 1, &tail0, 8, cudaMemcpyHostToDevice, st));
  // one pass: read ranges, tail, masked text, and the 2-bit packed text + non-ACGT bitmap the batches below score in place
  const uint64_t pk_bytes = (swb::fq_tiles(0, end) + 1) * 4096;        // the kernel's tiles cover [begin & ~511, end)
  if (c->q_pk.reserve(pk_bytes / 4 + 64) || c->q_bad.reserve(pk_bytes / 128 + 64)) return 1;
  k += swb::launch_fq_extract(d_text, begin, end, c->fq_tile_count.as<uint32_t>(), c->fq_tile_prefix.as<uint64_t>(), d_scal, d_beg, d_end, R,
                              reinterpret_cast<unsigned long long*>(d_scal + 1), final_segment, c->q_pk.as<uint32_t>(), c->q_bad.as<uint32_t>(), st);
  if (const char* dump = std::getenv("SWB_FASTQ_DUMP")) {        // debug: text after masking and the read ranges
    std::vector<uint8_t> ht(end - begin); std::vector<uint64_t> hb(R), he(R);
    CUDA_TRY(cudaMemcpyAsync(ht.data(), d_text + begin, end - begin, cudaMemcpyDeviceToHost, st));
    if (R) { CUDA_TRY(cudaMemcpyAsync(hb.data(), d_beg, R * 8, cudaMemcpyDeviceToHost, st)); CUDA_TRY(cudaMemcpyAsync(he.data(), d_end, R * 8, cudaMemcpyDeviceToHost, st)); }
    CUDA_TRY(cudaStreamSynchronize(st));
    if (FILE* f = std::fopen(dump, "wb")) {
      const uint64_t hdr[4] = {begin, end, R, newlines};
      std::fwrite(hdr, 8, 4, f); std::fwrite(ht.data(), 1, ht.size(), f); std::fwrite(hb.data(), 8, R, f); std::fwrite(he.data(), 8, R, f);
      std::fclose(f);
    }
  }
  if (dbg) cudaEventRecord(te[3], st);
  const uint64_t batch = 4ull << 20;
  if (R) {
    const uint64_t nb = std::min(R, batch);
    if (c->win_beg.reserve(nb * 8) || c->win_end.reserve(nb * 8) || c->out.reserve(nb * sizeof(swb_result))) return 1;
  }
  for (uint64_t a = 0; a < R; a += batch) {
    const uint64_t n = std::min(batch, R - a);
    k += swb::launch_fq_windows(file_index, first_read + a, n, c->ref_len, window_len, d_beg + a, d_end + a, c->win_beg.as<uint64_t>(),
                                c->win_end.as<uint64_t>(), st);
    const RefSrc rref = resident_ref(c);
    uint32_t read_bound = 0xffffffffu;                            // unknown: 160 rows, mid and long lists
    for (const auto& h : c->fq_read_hint) if (h.first == file_index && h.second >= 1 && h.second <= 128) read_bound = 128;
    if (run_device_pipeline(c, c, c->ev, d_text, d_beg + a, end, c->ref_bytes.as<uint8_t>(), c->win_beg.as<uint64_t>(),
                            c->win_end.as<uint64_t>(), 0, &rref, n, read_bound, window_len, c->out.as<swb_result>(), &k, d_end + a, true)) return 1;
    k += swb::launch_fq_reduce(c->out.as<swb_result>(), d_beg + a, d_end + a, n, reinterpret_cast<unsigned long long*>(d_scal + 2), st);
  }
  if (dbg) cudaEventRecord(te[4], st);
  CUDA_TRY(cudaMemcpyAsync(h_scal, d_scal, 64, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  if (dbg) {
    float a, b2, c2, d2;
    cudaEventElapsedTime(&a, te[0], te[1]); cudaEventElapsedTime(&b2, te[1], te[2]); cudaEventElapsedTime(&c2, te[2], te[3]); cudaEventElapsedTime(&d2, te[3], te[4]);
    std::fprintf(stderr, "[fastq] %llu blocks, %.1f MB text, %llu reads: (inflate %s) index %.2f, extract+mask+pack %.2f, score %.2f ms; call wall %.2f ms\n",
                 (unsigned long long)n_blocks, (end - begin) / 1e6, (unsigned long long)R, a > 1e9 ? "?" : "on its own stream", b2, c2, d2,
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
    for (auto& e : te) cudaEventDestroy(e);
  }
  if (!final_segment) {
    const uint64_t tail_start = h_scal[1];
    const uint64_t tl = end - tail_start;
    if (tl > carry_cap || tl > kFqCarryRoom || (tl && !carry_out)) { *status = 1; return 0; }   // a single record larger than the carry buffer (or than
                                                                                              // the next call would take; fq_extract_kernel only looks that far back)
    if (tl) {
      CUDA_TRY(cudaMemcpyAsync(carry_out, d_text + tail_start, tl, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
    }
    *carry_out_len = tl;
  }
  if (R) {                                                       // this segment's longest read: the next one's bound
    const uint32_t longest = (uint32_t)std::min<uint64_t>(h_scal[6], 0xffffffffull);
    bool found = false;
    for (auto& h : c->fq_read_hint) if (h.first == file_index) { h.second = longest; found = true; }
    if (!found) { if (c->fq_read_hint.size() >= 64) c->fq_read_hint.erase(c->fq_read_hint.begin()); c->fq_read_hint.emplace_back(file_index, longest); }
  }
  *score_sum = (int64_t)h_scal[2]; *n_reads = R; *n_bases = h_scal[3];
  *n_lines = final_segment ? lines : 4 * R;               // lines of the records scored here (the carried ones count next time)
  c->last_kernels = k; c->host_path = false; c->timings_pending = false;
  return 0;
}

int swb_fastq_bgzf_score(swb_ctx* c, const uint8_t* comp, uint64_t comp_bytes, const swb_bgzf_block* blocks, uint64_t n_blocks,
                         const uint8_t* carry, uint64_t carry_len, int final_segment,
                         uint64_t file_index, uint64_t first_read, uint32_t window_len,
                         int64_t* score_sum, uint64_t* n_reads, uint64_t* n_bases, uint64_t* n_lines,
                         uint8_t* carry_out, uint64_t carry_cap, uint64_t* carry_out_len, int* status)
{
  return drain_after_error(c, fastq_bgzf_score_impl(c, comp, comp_bytes, blocks, n_blocks, carry, carry_len, final_segment, file_index, first_read,
                                                    window_len, score_sum, n_reads, n_bases, n_lines, carry_out, carry_cap, carry_out_len, status));
}

// ---- alignments behind the scores ----
static int traceback_batch_impl(swb_ctx* c, const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro, uint64_t n_pairs,
                        const swb_result* results, swb_alignment* out, uint32_t* cigar, uint64_t cigar_cap, uint64_t* cigar_used)
{
  if (!c) return fail("null ctx");
  if (cigar_used) *cigar_used = 0;
  if (n_pairs == 0) return 0;
  if (!qo || !ro || !results || !out || !cigar_used || (cigar_cap && !cigar)) return fail("swb_traceback_batch: null pointer");
  if (n_pairs >= (1ull << 32)) return fail("swb: at most 2^32-1 pairs per batch");
  if (qo[0] != 0 || ro[0] != 0) return fail("swb_traceback_batch: offsets must start at 0");
  CUDA_TRY(cudaSetDevice(c->device));
  // per-warp scratch: sized by the largest rectangle of the batch (rows 0..end_i, at most twice as many columns)
  uint64_t max_rows = 0, max_width = 0;
  for (uint64_t k = 0; k < n_pairs; ++k) {
    if (qo[k + 1] < qo[k] || ro[k + 1] < ro[k]) return fail("swb_traceback_batch: offsets must be non-decreasing");
    const swb_result& e = results[k];
    if (e.score <= 0 || e.end_i < 0 || e.end_j < 0 || (uint64_t)e.end_i >= qo[k + 1] - qo[k] || (uint64_t)e.end_j >= ro[k + 1] - ro[k]) continue;
    const uint64_t rows = (uint64_t)e.end_i + 1, width = std::min<uint64_t>((uint64_t)e.end_j + 1, 2 * rows);
    max_rows = std::max(max_rows, rows); max_width = std::max(max_width, width);
  }
  const uint32_t cpl = swb::tb_pick_cpl(max_width);
  const uint64_t rows_b = (swb::tb_rows_bytes(max_rows, max_width, cpl) + 255) & ~255ull;
  const uint64_t dirs_b = (swb::tb_dirs_bytes(max_rows, max_width, cpl) + 255) & ~255ull;
  const bool in_smem = rows_b * 4 <= 48 * 1024;                // the two H rows + the runs of four warps in one CTA's shared memory
  const uint64_t need = dirs_b + (in_smem ? 0 : rows_b) + 256;
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const uint64_t budget = std::max<uint64_t>(c->tb_scratch.cap, std::min<uint64_t>(4ull << 30, free_b / 2));
  if (need > budget) return fail("swb_traceback_batch: a pair needs " + std::to_string(need >> 20) + " MiB of direction bits, more than the device has room for");
  int warps = c->sm_count * 32;
  if (in_smem) warps = std::min<int>(warps, c->sm_count * 4 * (int)std::max<uint64_t>(1, (200 * 1024) / (rows_b * 4)));
  warps = (int)std::min<uint64_t>((uint64_t)warps, std::max<uint64_t>(1, budget / need));
  warps = std::max(4, warps / 4 * 4);                          // whole CTAs of four warps
  if ((uint64_t)warps * need > budget && warps > 4) warps -= 4;
  warps = (int)std::min<uint64_t>((uint64_t)warps, (n_pairs + 3) / 4 * 4);
  const uint64_t qb = qo[n_pairs], rb = ro[n_pairs];
  if (c->q_bytes.reserve(qb + 64) || c->r_bytes.reserve(rb + 64) || c->q_off.reserve((n_pairs + 1) * 8) || c->r_off.reserve((n_pairs + 1) * 8) ||
      c->tb_res.reserve(n_pairs * sizeof(swb_result)) || c->tb_out.reserve(n_pairs * sizeof(swb_alignment)) ||
      c->tb_cigar.reserve(cigar_cap * 4 + 64) || c->tb_cursor.reserve(64) || c->tb_handled.reserve(n_pairs + 64) ||
      c->tb_scratch.reserve((uint64_t)warps * need)) return 1;
  cudaStream_t st = c->st;
  if (qb) CUDA_TRY(cudaMemcpyAsync(c->q_bytes.p, q, qb, cudaMemcpyHostToDevice, st));
  if (rb) CUDA_TRY(cudaMemcpyAsync(c->r_bytes.p, r, rb, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->q_off.p, qo, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->r_off.p, ro, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->tb_res.p, results, n_pairs * sizeof(swb_result), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(c->tb_cursor.p, 0, 64, st));
  swb::TracebackArgs a;
  a.q = c->q_bytes.as<uint8_t>(); a.qo = c->q_off.as<uint64_t>(); a.r = c->r_bytes.as<uint8_t>(); a.ro = c->r_off.as<uint64_t>();
  a.res = c->tb_res.as<swb_result>(); a.out = c->tb_out.as<swb_alignment>(); a.cigar = c->tb_cigar.as<uint32_t>(); a.cigar_cap = cigar_cap;
  a.cursor = c->tb_cursor.as<unsigned long long>(); a.n_pairs = n_pairs;
  a.scratch = c->tb_scratch.as<uint8_t>(); a.scratch_per_warp = need; a.dirs_per_warp = dirs_b; a.rows_per_warp = rows_b; a.rows_in_smem = in_smem;
  a.handled = c->tb_handled.as<uint8_t>();
  CUDA_TRY(cudaEventRecord(c->ev[0], st));
  c->last_kernels = swb::launch_traceback(a, cpl, warps, st);
  CUDA_TRY(cudaEventRecord(c->ev[1], st));
  unsigned long long h_cursor[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(h_cursor, c->tb_cursor.p, 16, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(out, c->tb_out.p, n_pairs * sizeof(swb_alignment), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  for (float& v : c->last_ms) v = 0;
  cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[1]);          // swb_last_timings: device_ms = the traceback kernel
  c->host_path = false; c->timings_pending = false;
  *cigar_used = h_cursor[1];
  if (h_cursor[1] > cigar_cap)
    return fail("swb_traceback_batch: the batch has " + std::to_string(h_cursor[1]) + " operations, cigar_cap is " + std::to_string(cigar_cap));
  if (h_cursor[1]) CUDA_TRY(cudaMemcpy(cigar, c->tb_cigar.p, h_cursor[1] * 4, cudaMemcpyDeviceToHost));
  c->host_path = false; c->timings_pending = false;
  return 0;
}

int swb_traceback_batch(swb_ctx* c, const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro, uint64_t n_pairs,
                        const swb_result* results, swb_alignment* out, uint32_t* cigar, uint64_t cigar_cap, uint64_t* cigar_used)
{
  return drain_after_error(c, traceback_batch_impl(c, q, qo, r, ro, n_pairs, results, out, cigar, cigar_cap, cigar_used));
}

// SWB_GUARD: how many guard zones of this process's arenas on the context's device are damaged (0 = none; -1 = guard mode
// is off).  Synchronises the device.  `report` receives one line per damaged zone.
int swb_debug_guard_check(swb_ctx* c, char* report, uint64_t report_cap, uint64_t* n_arenas)
{
  if (report && report_cap) report[0] = 0;
  if (n_arenas) *n_arenas = 0;
  if (!g_guard) return -1;
  if (!c) return -1;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  std::vector<GuardEntry> list;
  { std::lock_guard<std::mutex> lk(g_guard_mu); list = g_guard_list; }
  std::vector<uint8_t> h(kGuardBytes);
  int bad = 0; std::string rep; uint64_t seen = 0;
  for (const GuardEntry& ge : list) {
    if (ge.device != c->device) continue;
    ++seen;
    for (int side = 0; side < 2; ++side) {
      const uint8_t* z = side ? ge.base + kGuardBytes + ge.cap : ge.base;
      if (cudaMemcpy(h.data(), z, kGuardBytes, cudaMemcpyDeviceToHost) != cudaSuccess) { ++bad; rep += "guard zone unreadable\n"; continue; }
      size_t first = kGuardBytes, n = 0;
      for (size_t k = 0; k < kGuardBytes; ++k) if (h[k] != 0xA5) { if (first == kGuardBytes) first = k; ++n; }
      if (n) {
        ++bad;
        rep += std::string(side ? "after" : "before") + " an arena of " + std::to_string(ge.cap) + " bytes: " + std::to_string(n) +
               " bytes damaged, first at zone offset " + std::to_string(first) + "\n";
      }
    }
  }
  if (n_arenas) *n_arenas = seen;
  if (report && report_cap) std::snprintf(report, report_cap, "%s", rep.c_str());
  return bad;
}

int swb_score_pair(swb_ctx* c, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, swb_result* out)
{
  if (!out) return fail("swb_score_pair: null out");
  const uint64_t qo[2] = {0, n1}, ro[2] = {0, n2};
  return swb_score_batch(c, s1, qo, s2, ro, 1, out);
}

int swb_last_timings(swb_ctx* c, float* ms, int* kernels)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  if (c->timings_pending) {
    for (int i = 0; i < kLanes; ++i) CUDA_TRY(cudaStreamSynchronize(c->lane(i)->st));
    for (float& v : c->last_ms) v = 0;
    c->last_routing[0] = c->last_routing[1] = c->last_routing[2] = 0;
    for (auto& v : c->last_routing5) v = 0;
    if (!c->host_path) {
      cudaEventElapsedTime(&c->last_ms[0], c->ev[0], c->ev[1]);
      cudaEventElapsedTime(&c->last_ms[1], c->ev[1], c->ev[2]);
      cudaEventElapsedTime(&c->last_ms[2], c->ev[2], c->ev[3]);
      cudaEventElapsedTime(&c->last_ms[3], c->ev[0], c->ev[3]);
      swb::Counters h;
      CUDA_TRY(cudaMemcpy(&h, c->counters.p, sizeof(h), cudaMemcpyDeviceToHost));
      c->last_routing[0] = h.n_short + h.n_mid; c->last_routing[1] = h.n_generic + h.n_bytes; c->last_routing[2] = h.n_long;
      c->last_routing5[0] = h.n_short; c->last_routing5[1] = h.n_mid; c->last_routing5[2] = h.n_long; c->last_routing5[3] = h.n_bytes; c->last_routing5[4] = h.n_generic;
    } else {
      // sums over the chunks of the call (chunks overlap in time, so [3] is the span first event -> last event)
      for (size_t ch = 0; ch < c->last_chunks; ++ch) {
        cudaEvent_t* ev = c->chunk_ev[ch].ev;
        float t;
        cudaEventElapsedTime(&t, ev[0], ev[1]); c->last_ms[0] += t;
        cudaEventElapsedTime(&t, ev[1], ev[2]); c->last_ms[1] += t;
        cudaEventElapsedTime(&t, ev[2], ev[3]); c->last_ms[2] += t;
        cudaEventElapsedTime(&t, ev[4], ev[5]); c->last_ms[4] += t;
        cudaEventElapsedTime(&t, ev[6], ev[7]); c->last_ms[5] += t;
        const swb::Counters& hc = c->h_counters[ch];
        c->last_routing[0] += hc.n_short + hc.n_mid; c->last_routing[1] += hc.n_generic + hc.n_bytes; c->last_routing[2] += hc.n_long;
        c->last_routing5[0] += hc.n_short; c->last_routing5[1] += hc.n_mid; c->last_routing5[2] += hc.n_long; c->last_routing5[3] += hc.n_bytes; c->last_routing5[4] += hc.n_generic;
      }
      if (std::getenv("SWB_DEBUG_TIMELINE")) {                 // per chunk, ms since the first event of the call
        for (size_t ch = 0; ch < c->last_chunks; ++ch) {
          float t[6]; const int idx[6] = {4, 5, 1, 2, 3, 7};
          for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], c->chunk_ev[0].ev[4], c->chunk_ev[ch].ev[idx[k]]);
          std::fprintf(stderr, "[timeline] chunk %2zu lane %zu pairs %8llu: h2d %7.3f -> %7.3f | pack+classify -> %7.3f | short -> %7.3f | other -> %7.3f | d2h -> %7.3f\n",
                       ch, ch % (size_t)c->n_lanes, (unsigned long long)(c->chunk_bounds[ch + 1] - c->chunk_bounds[ch]), t[0], t[1], t[2], t[3], t[4], t[5]);
        }
      }
      if (c->last_chunks) {
        float span = 0;
        for (size_t ch = 0; ch < c->last_chunks; ++ch) {      // lanes finish out of order: take the latest end
          float t; cudaEventElapsedTime(&t, c->chunk_ev[0].ev[4], c->chunk_ev[ch].ev[7]);
          span = std::max(span, t);
        }
        c->last_ms[3] = span;
      }
    }
    c->timings_pending = false;
  }
  if (ms) std::memcpy(ms, c->last_ms, sizeof(c->last_ms));
  if (kernels) *kernels = c->last_kernels;
  return 0;
}

int swb_last_routing(swb_ctx* c, uint64_t* counts)
{
  if (!c || !counts) return fail("null pointer");
  if (swb_last_timings(c, nullptr, nullptr)) return 1;
  counts[0] = c->last_routing[0]; counts[1] = c->last_routing[1]; counts[2] = c->last_routing[2];
  return 0;
}

int swb_last_routing_ex(swb_ctx* c, uint64_t* counts)
{
  if (!c || !counts) return fail("null pointer");
  if (swb_last_timings(c, nullptr, nullptr)) return 1;
  for (int k = 0; k < 5; ++k) counts[k] = c->last_routing5[k];
  return 0;
}

int swb_set_mid_path(swb_ctx* c, int on) { if (!c) return fail("null ctx"); c->mid_path = on != 0; return 0; }

int swb_ref_compat_align(swb_ctx* c, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2,
                         uint32_t dev_max_wg, int32_t* out)
{
  if (!c || !out) return fail("null pointer");
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t len = std::min(n1, n2);                           // aligner.rs:412
  if (len == 0) { *out = 0; return 0; }                            // aligner.rs:413-416
  if (dev_max_wg == 0) return fail("swb_ref_compat_align: work-group size must be positive");
  const uint32_t wgs = std::min<uint32_t>(dev_max_wg, 1024);       // aligner.rs:422, gpu.rs:9
  const uint64_t groups = std::min<uint64_t>((len + wgs - 1) / wgs, 1000000ull);   // aligner.rs:423-424
  if (len > 1000000ull * 1024ull)                                  // aligner.rs:436-456 (work-item limit)
    return fail("Sequence too large (" + std::to_string(len) + " bytes), max allowed: 1024000000 bytes");
  if (c->q_bytes.reserve(n1 + 64) || c->r_bytes.reserve(n2 + 64) || c->misc.reserve(64)) return 1;
  cudaStream_t st = c->st;
  CUDA_TRY(cudaMemcpyAsync(c->q_bytes.p, s1, n1, cudaMemcpyHostToDevice, st));   // full buffers, aligner.rs:478-492
  CUDA_TRY(cudaMemcpyAsync(c->r_bytes.p, s2, n2, cudaMemcpyHostToDevice, st));
  swb::launch_ref_compat(c->q_bytes.as<uint8_t>(), c->r_bytes.as<uint8_t>(), len, wgs, groups, c->misc.as<int32_t>(), st);
  CUDA_TRY(cudaMemcpyAsync(out, c->misc.p, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int swb_last_row_max(swb_ctx* c, const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, int32_t* out)
{
  if (!c || !out) return fail("null pointer");
  CUDA_TRY(cudaSetDevice(c->device));
  if (n1 == 0 || n2 == 0) { *out = 0; return 0; }
  if (n1 > 0x7fffffffull || n2 > 0x7fffffffull) return fail("Sequence too large (more than 2^31-1 bytes)");
  const uint64_t stride = (n2 + 32) & ~31ull;
  if (c->q_bytes.reserve(n1 + 64) || c->r_bytes.reserve(n2 + 64) || c->q_off.reserve(16) || c->r_off.reserve(16) ||
      c->out.reserve(sizeof(swb_result)) || c->generic_list.reserve(64) || c->counters.reserve(sizeof(swb::Counters)) ||
      c->scratch.reserve(stride * 4) || c->misc.reserve(64)) return 1;
  cudaStream_t st = c->st;
  const uint64_t qo[2] = {0, n1}, ro[2] = {0, n2};
  CUDA_TRY(cudaMemcpyAsync(c->q_bytes.p, s1, n1, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->r_bytes.p, s2, n2, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->q_off.p, qo, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(c->r_off.p, ro, 16, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaStreamSynchronize(st));          // qo/ro live on this stack frame
  swb::BatchView b{};
  b.q_bytes = c->q_bytes.as<uint8_t>(); b.q_beg = c->q_off.as<uint64_t>(); b.q_end = b.q_beg + 1;
  b.r_bytes = c->r_bytes.as<uint8_t>(); b.r_beg = c->r_off.as<uint64_t>(); b.r_end = b.r_beg + 1;
  b.n_pairs = 1; b.generic_list = c->generic_list.as<uint32_t>(); b.counters = c->counters.as<swb::Counters>();
  b.out = c->out.as<swb_result>(); b.scratch = c->scratch.as<int32_t>(); b.scratch_stride = stride;
  swb::launch_generic_single(b, c->misc.as<int32_t>(), st);
  CUDA_TRY(cudaMemcpyAsync(out, c->misc.p, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int swb_pack2bit(swb_ctx* c, const uint8_t* bytes, uint64_t n, uint32_t* packed_words, uint32_t* bitmap)
{
  if (!c) return fail("null ctx");
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t nw = (n + 15) / 16, nb = (nw + 31) / 32;
  if (c->q_bytes.reserve(n + 64) || c->q_pk.reserve(nw * 4 + 64) || c->q_bad.reserve(nb * 4 + 64)) return 1;
  cudaStream_t st = c->st;
  CUDA_TRY(cudaMemcpyAsync(c->q_bytes.p, bytes, n, cudaMemcpyHostToDevice, st));
  swb::launch_pack2bit(c->q_bytes.as<uint8_t>(), n, c->q_pk.as<uint32_t>(), c->q_bad.as<uint32_t>(), st);
  CUDA_TRY(cudaMemcpyAsync(packed_words, c->q_pk.p, nw * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(bitmap, c->q_bad.p, nb * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// pack stage alone on device-resident bytes (bench: HBM roofline of the packing kernel)
int swb_pack2bit_device(swb_ctx* c, const uint8_t* d_bytes, uint64_t n, uint32_t* d_words, uint32_t* d_bitmap)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  swb::launch_pack2bit(d_bytes, n, d_words, d_bitmap, c->st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int swb_synth_device(swb_ctx* c, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, uint32_t window_len,
                     int distribution, uint8_t* d_q, uint64_t* d_qo, uint8_t* d_r, uint64_t* d_ro)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  swb::launch_synth(first_pair, n_pairs, read_len, window_len, distribution, d_q, d_qo, d_r, d_ro, c->st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

int swb_synth_device_ref(swb_ctx* c, const uint8_t* d_ref, uint64_t ref_len, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len,
                         uint32_t window_len, int distribution, uint8_t* d_q, uint64_t* d_qo, uint8_t* d_r, uint64_t* d_ro, uint64_t* d_win_start)
{
  if (!c) return fail("null ctx");
  if (!d_ref || window_len == 0 || window_len > ref_len) return fail("swb_synth_device_ref: the window must fit the reference");
  CUDA_TRY(cudaSetDevice(c->device));
  swb::launch_synth_ref(d_ref, ref_len, first_pair, n_pairs, read_len, window_len, distribution, d_q, d_qo, d_r, d_ro, d_win_start, c->st);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// raw device / pinned memory for callers without a CUDA runtime of their own (CLI, ctypes tests)
int swb_malloc_device(swb_ctx* c, uint64_t bytes, void** out)
{
  if (!c || !out) return fail("null pointer");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
}
int swb_free_device(swb_ctx* c, void* p) { if (c) cudaSetDevice(c->device); cudaFree(p); return 0; }
int swb_bind_thread(swb_ctx* c)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  return 0;
}

int swb_malloc_pinned(uint64_t bytes, void** out)
{
  if (!out) return fail("null pointer");
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
}
int swb_free_pinned(void* p) { cudaFreeHost(p); return 0; }
int swb_memcpy_d2h(swb_ctx* c, void* dst, const void* src, uint64_t bytes)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  return 0;
}
int swb_memcpy_h2d(swb_ctx* c, void* dst, const void* src, uint64_t bytes)
{
  if (!c) return fail("null ctx");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->st));
  CUDA_TRY(cudaStreamSynchronize(c->st));
  return 0;
}

// =====================================================================================
// Several devices behind one handle (north_star: "each chunk is partitioned across the 8 GPUs of one box with per-GPU
// streams"; the reference drives devices[0] only, gpu.rs:117-131, main.rs:95).  One persistent host thread per device owns
// that device's swb_ctx (three pipeline streams each); a batch is cut into contiguous slices of pairs of about equal
// bytes, every slice is scored by its device into its own part of the caller's result array -- no inter-GPU traffic, no
// collective.  The contexts are created by their threads, i.e. concurrently.
// =====================================================================================
}  // extern "C"

struct swb_multi {
  struct Worker {
    int device = 0; swb_ctx* ctx = nullptr;
    std::thread th; std::mutex mu; std::condition_variable cv;
    std::function<int(swb_ctx*)> job; bool has_job = false, done = false, quit = false;
    int rc = 0; std::string err;
  };
  std::vector<Worker*> w;
  // run fn(k, ctx) on every worker concurrently; returns the first failure
  int run(const std::function<int(int, swb_ctx*)>& fn)
  {
    for (size_t k = 0; k < w.size(); ++k) {
      Worker* x = w[k];
      std::lock_guard<std::mutex> lk(x->mu);
      x->job = [fn, k](swb_ctx* c) { return fn((int)k, c); };
      x->has_job = true; x->done = false;
      x->cv.notify_all();
    }
    int rc = 0;
    for (Worker* x : w) {
      std::unique_lock<std::mutex> lk(x->mu);
      x->cv.wait(lk, [x] { return x->done; });
      if (x->rc && !rc) { rc = x->rc; g_err = "device " + std::to_string(x->device) + ": " + x->err; }
    }
    return rc;
  }
  static void loop(Worker* x)
  {
    for (;;) {
      std::function<int(swb_ctx*)> job;
      {
        std::unique_lock<std::mutex> lk(x->mu);
        x->cv.wait(lk, [x] { return x->has_job || x->quit; });
        if (x->quit) return;
        job = x->job; x->has_job = false;
      }
      const int rc = job(x->ctx);
      std::lock_guard<std::mutex> lk(x->mu);
      x->rc = rc; x->err = rc ? std::string(swb_last_error()) : std::string();
      x->done = true;
      x->cv.notify_all();
    }
  }
};

// pair index where the first `part` of `parts` equal shares of the batch's bytes ends (offsets are monotone)
static uint64_t split_point(const uint64_t* qo, const uint64_t* ro, uint64_t n_pairs, uint64_t part, uint64_t parts)
{
  if (part == 0) return 0;
  if (part >= parts) return n_pairs;
  auto bytes_before = [&](uint64_t k) { return (qo[k] - qo[0]) + (ro ? ro[k] - ro[0] : 0); };
  const uint64_t total = bytes_before(n_pairs);
  if (total == 0) return n_pairs * part / parts;
  const uint64_t want = (uint64_t)((unsigned __int128)total * part / parts);
  uint64_t lo = 0, hi = n_pairs;
  while (lo < hi) { const uint64_t mid = (lo + hi) / 2; if (bytes_before(mid) < want) lo = mid + 1; else hi = mid; }
  return lo;
}

extern "C" {

int swb_create_multi(swb_multi** out, const int* device_ids, int n_devices, const swb_params* params)
{
  if (!out) return fail("swb_create_multi: null out pointer");
  *out = nullptr;
  const int visible = swb_device_count();
  if (visible <= 0) return fail("error: gpu acceleration is required and no compatible gpu was found");   // main.rs:161
  if (n_devices <= 0 || !device_ids) { n_devices = visible; device_ids = nullptr; }                      // all visible devices
  if (n_devices > 64) return fail("swb_create_multi: at most 64 devices");
  for (int k = 0; k < n_devices; ++k) {
    const int d = device_ids ? device_ids[k] : k;
    if (d < 0 || d >= visible) return fail("swb_create_multi: no such device");
    for (int j = 0; j < k; ++j) if ((device_ids ? device_ids[j] : j) == d) return fail("swb_create_multi: a device is listed twice");
  }
  swb_multi* m = new swb_multi();
  const swb_params pcopy = params ? *params : swb_params{swb::kMatch, swb::kMismatch, swb::kGap};
  for (int k = 0; k < n_devices; ++k) {
    auto* x = new swb_multi::Worker();
    x->device = device_ids ? device_ids[k] : k;
    m->w.push_back(x);
    x->th = std::thread(swb_multi::loop, x);
  }
  // every worker creates its own context: CUDA context creation, streams and module load run concurrently per device
  const int rc = m->run([m, pcopy](int k, swb_ctx*) { return swb_create(&m->w[k]->ctx, m->w[k]->device, &pcopy); });
  if (rc) { const std::string keep = g_err; swb_destroy_multi(m); g_err = keep; return 1; }
  *out = m;
  return 0;
}

void swb_destroy_multi(swb_multi* m)
{
  if (!m) return;
  for (auto* x : m->w) {
    { std::lock_guard<std::mutex> lk(x->mu); x->quit = true; x->cv.notify_all(); }
    if (x->th.joinable()) x->th.join();
    if (x->ctx) swb_destroy(x->ctx);
    delete x;
  }
  delete m;
}

int swb_multi_device_count(swb_multi* m) { return m ? (int)m->w.size() : 0; }
swb_ctx* swb_multi_ctx(swb_multi* m, int k) { return (m && k >= 0 && k < (int)m->w.size()) ? m->w[k]->ctx : nullptr; }

int swb_multi_score_batch(swb_multi* m, const uint8_t* q, const uint64_t* qo, const uint8_t* r, const uint64_t* ro,
                          uint64_t n_pairs, swb_result* out)
{
  if (!m) return fail("null ctx");
  if (n_pairs == 0) return 0;
  if (!qo || !ro || !out) return fail("swb_multi_score_batch: null pointer");
  if (qo[0] != 0 || ro[0] != 0) return fail("swb_multi_score_batch: offsets must start at 0");
  const uint64_t parts = m->w.size();
  static const uint8_t empty = 0;
  const uint8_t* qq = q ? q : &empty; const uint8_t* rr = r ? r : &empty;
  return m->run([=](int k, swb_ctx* c) -> int {
    const uint64_t lo = split_point(qo, ro, n_pairs, (uint64_t)k, parts), hi = split_point(qo, ro, n_pairs, (uint64_t)k + 1, parts);
    if (hi <= lo) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return fail("cudaSetDevice failed");
    return score_host_batch(c, "swb_multi_score_batch", qq, qo + lo, rr, ro + lo, nullptr, nullptr, nullptr, hi - lo, out + lo);
  });
}

int swb_multi_set_reference(swb_multi* m, const uint8_t* ref, uint64_t n)
{
  if (!m) return fail("null ctx");
  return m->run([=](int, swb_ctx* c) { return swb_set_reference(c, ref, n); });
}

int swb_multi_score_batch_vs_reference(swb_multi* m, const uint8_t* q, const uint64_t* qo, uint64_t n_pairs,
                                       const uint64_t* win_start, const uint32_t* win_len, swb_result* out)
{
  if (!m) return fail("null ctx");
  if (n_pairs == 0) return 0;
  if (!qo || !win_start || !win_len || !out) return fail("swb_multi_score_batch_vs_reference: null pointer");
  if (qo[0] != 0) return fail("swb_multi_score_batch_vs_reference: offsets must start at 0");
  const uint64_t parts = m->w.size();
  static const uint8_t empty = 0;
  const uint8_t* qq = q ? q : &empty;
  return m->run([=](int k, swb_ctx* c) -> int {
    const uint64_t lo = split_point(qo, nullptr, n_pairs, (uint64_t)k, parts), hi = split_point(qo, nullptr, n_pairs, (uint64_t)k + 1, parts);
    if (hi <= lo) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess) return fail("cudaSetDevice failed");
    const RefSrc ref = resident_ref(c);
    return score_host_batch(c, "swb_multi_score_batch_vs_reference", qq, qo + lo, nullptr, nullptr, win_start + lo, win_len + lo, &ref, hi - lo, out + lo);
  });
}

}  // extern "C"
