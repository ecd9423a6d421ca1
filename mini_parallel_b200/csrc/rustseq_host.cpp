// rustseq_host.cpp -- C++ mirror of the reference's Rust host side for the alignment path
// (smith_waterman/src/aligner.rs, gpu.rs, main.rs).  See include/rustseq_host.h.  All scoring goes
// through the C ABI of swb200.h; there is no CPU scoring code in this file.
#include "../../include/rustseq_host.h"
#include "host_gunzip.h"
#include "host_pgunzip.h"
#include <zlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <filesystem>
#include <system_error>
#include <vector>

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }

// SWB_DEBUG: seconds since the library was loaded (~ process start), for the start-up / tear-down accounting of the CLI
const std::chrono::steady_clock::time_point g_loaded = std::chrono::steady_clock::now();
void dbg_stamp(const char* what)
{
  if (std::getenv("SWB_DEBUG") || std::getenv("SWB_STAMPS"))
    std::fprintf(stderr, "[main] %-36s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - g_loaded).count());
}

bool ref_compat_mode()
{
  const char* m = std::getenv("SWB_GPU_ALIGN_MODE");
  return m && std::string(m) == "ref_compat";
}

// ---- process-wide context per device: the OPENCL_CONTEXT singleton of gpu.rs:13-14, :97-115 ----
// One lock PER ORDINAL: the --full-wgs driver's per-GPU threads build their contexts (CUDA context, six streams, module
// load: ~1 s each) at the same time instead of one after another.  g_use_mu serialises the callers that SHARE a
// process-wide context (rsm_gpu_align*, like the reference's Mutex around its one queue).
std::mutex g_ctx_mu[64];
std::mutex g_use_mu[64];
swb_ctx* g_ctx[64] = {};

swb_ctx* context_for(int ordinal)
{
  if (ordinal < 0 || ordinal >= 64) { g_err = "Failed to get GPU context: bad device ordinal"; return nullptr; }
  std::lock_guard<std::mutex> lk(g_ctx_mu[ordinal]);
  if (!g_ctx[ordinal]) {
    swb_ctx* c = nullptr;
    if (swb_create(&c, ordinal, nullptr) != 0) { g_err = std::string("Failed to get GPU context: ") + swb_last_error(); return nullptr; }
    g_ctx[ordinal] = c;
  }
  return g_ctx[ordinal];
}

// ---- Rust's str::parse::<usize>() and its ParseIntError texts (aligner.rs:13-14) ----
int parse_usize(const std::string& s, uint64_t* out, std::string* why)
{
  size_t i = 0;
  if (s.empty()) { *why = "cannot parse integer from empty string"; return 1; }
  if (s[0] == '+') { i = 1; if (s.size() == 1) { *why = "invalid digit found in string"; return 1; } }
  uint64_t v = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') { *why = "invalid digit found in string"; return 1; }
    const uint64_t d = (uint64_t)(s[i] - '0');
    if (v > (UINT64_MAX - d) / 10) { *why = "number too large to fit in target type"; return 1; }
    v = v * 10 + d;
  }
  *out = v;
  return 0;
}

bool valid_utf8(const uint8_t* p, size_t n)
{
  size_t i = 0;
  while (i + 8 <= n) {                             // ASCII fast path, 8 bytes at a time (FASTQ lines are ASCII)
    uint64_t w; std::memcpy(&w, p + i, 8);
    if (w & 0x8080808080808080ull) break;
    i += 8;
  }
  while (i < n) {
    const uint8_t c = p[i];
    if (c < 0x80) { ++i; continue; }
    int extra; uint32_t cp;
    if ((c & 0xE0) == 0xC0) { extra = 1; cp = c & 0x1F; }
    else if ((c & 0xF0) == 0xE0) { extra = 2; cp = c & 0x0F; }
    else if ((c & 0xF8) == 0xF0) { extra = 3; cp = c & 0x07; }
    else return false;
    if (i + (size_t)extra >= n) return false;
    for (int k = 1; k <= extra; ++k) { if ((p[i + k] & 0xC0) != 0x80) return false; cp = (cp << 6) | (p[i + k] & 0x3F); }
    if ((extra == 1 && cp < 0x80) || (extra == 2 && cp < 0x800) || (extra == 3 && cp < 0x10000) || cp > 0x10FFFF ||
        (cp >= 0xD800 && cp <= 0xDFFF)) return false;
    i += extra + 1;
  }
  return true;
}

// Host memory the GPU can copy from asynchronously (USE_PINNED_MEMORY, aligner.rs:466-475): std::vector over
// swb_malloc_pinned.  The WGS pipeline keeps its chunk buffers in it so H2D overlaps inflate and scoring.
// (rsm_debug_bgzf_segments walks a file with the readers alone, no GPU in sight: its buffers are plain memory.)
std::atomic<int> g_plain_buffers{0};
template <class T> struct PinnedAlloc {
  using value_type = T;
  PinnedAlloc() = default;
  template <class U> PinnedAlloc(const PinnedAlloc<U>&) {}
  T* allocate(size_t n)
  {
    void* p = nullptr;
    if (g_plain_buffers.load()) { p = std::malloc(n * sizeof(T) + 16); if (!p) throw std::bad_alloc(); *(uint64_t*)p = 0x504C41494E; return (T*)((char*)p + 16); }
    if (swb_malloc_pinned(n * sizeof(T) + 16, &p) != 0 || !p) throw std::bad_alloc();
    *(uint64_t*)p = 0;
    return (T*)((char*)p + 16);
  }
  void deallocate(T* p, size_t)
  {
    void* base = (char*)p - 16;
    if (*(uint64_t*)base == 0x504C41494E) std::free(base); else swb_free_pinned(base);
  }
  template <class U> bool operator==(const PinnedAlloc<U>&) const { return true; }
  template <class U> bool operator!=(const PinnedAlloc<U>&) const { return false; }
};
template <class T> using pinned_vector = std::vector<T, PinnedAlloc<T>>;

// Beyond this many decoder threads a file's single parsing thread is the limit (it splits lines at ~2 GB/s of text).
constexpr unsigned kMaxInflateThreads = 12;

// ---- streaming FASTQ reader: the body of process_fastq_file_in_chunks (aligner.rs:107-178), pull style ----
class FastqReader {
 public:
  ~FastqReader() { close(); }
  // spare_threads: threads this file may use beside the caller's, which parses (plain .gz only): 0 = inflate here, 1-2 = one
  // thread inflates ahead (hgz::AsyncGunzip), >= 3 = that many decode chunks of the file side by side (hgz::ParallelGunzip).
  // SWB_INFLATE_THREADS=n overrides the caller's figure, SWB_ASYNC_INFLATE=0/1 is the older switch between 0 and 1.
  int open(const std::string& path, unsigned spare_threads = 0)
  {
    path_ = path;
    if (const char* v = std::getenv("SWB_ASYNC_INFLATE")) spare_threads = std::atoi(v) != 0 ? 1 : 0;
    if (const char* v = std::getenv("SWB_INFLATE_THREADS")) { const int n = std::atoi(v); spare_threads = n < 0 ? 0u : (unsigned)std::min(n, 64); }
    const bool gz = path.size() >= 3 && path.compare(path.size() - 3, 3, ".gz") == 0;   // aligner.rs:109
    const char* how = std::getenv("SWB_HOST_INFLATE");
    if (gz && !(how && std::string(how) == "zlib")) {
      // in-process inflate instead of a `zcat` child (aligner.rs:111-120): this repository's decoder (host_gunzip.h), 1.6-1.7x
      // zlib on FASTQ text and the same behaviour at the edges; SWB_HOST_INFLATE=zlib selects gzread
      if (spare_threads >= 3) {
        pz_.set_threads(spare_threads);
        if (!pz_.open(path.c_str())) return fail("Failed to open file " + path + ": " + std::strerror(errno));
        use_pz_ = true;
      } else if (spare_threads) {
        if (!az_.open(path.c_str())) return fail("Failed to open file " + path + ": " + std::strerror(errno));
        use_az_ = true;
      } else {
        if (!hz_.open(path.c_str())) return fail("Failed to open file " + path + ": " + std::strerror(errno));
        use_hz_ = true;
      }
    } else if (gz) {
      gz_ = gzopen(path.c_str(), "rb");
      if (!gz_) return fail("Failed to open file " + path + ": " + std::strerror(errno));
      gzbuffer(gz_, 1 << 20);
    } else {
      fp_ = std::fopen(path.c_str(), "rb");        // aligner.rs:123-125
      if (!fp_) return fail("Failed to open file " + path + ": " + std::strerror(errno));
    }
    buf_.resize(4 << 20);
    return 0;
  }
  void close() { if (gz_) gzclose(gz_); if (fp_) std::fclose(fp_); gz_ = nullptr; fp_ = nullptr; hz_.close(); use_hz_ = false; az_.close(); use_az_ = false; pz_.close(); use_pz_ = false; }

  // Appends up to max_reads sequence lines (and at most max_bases bases, 0 = no cap) to bases/offs.
  // Returns 0 ok, 1 error; *eof set when the input is exhausted.
  template <class VB, class VO>
  int next_chunk(uint64_t max_reads, uint64_t max_bases, VB& bases, VO& offs, bool* eof)
  {
    bases.clear(); offs.clear(); offs.push_back(0);
    *eof = false;
    for (;;) {
      // consume complete lines in the buffer
      while (pos_ < len_) {
        const uint8_t* start = buf_.data() + pos_;
        const uint8_t* nl = (const uint8_t*)std::memchr(start, '\n', len_ - pos_);
        if (!nl) break;
        size_t n = (size_t)(nl - start);
        pos_ += n + 1;
        if (handle_line(start, n, bases, offs)) return 1;
        if (chunk_full(max_reads, max_bases, bases, offs)) return 0;
      }
      // refill, keeping the partial line
      if (pos_ > 0) { std::memmove(buf_.data(), buf_.data() + pos_, len_ - pos_); len_ -= pos_; pos_ = 0; }
      if (len_ == buf_.size()) buf_.resize(buf_.size() * 2);
      long got = 0;
      if (!at_eof_) {
        got = use_pz_ ? pz_.read(buf_.data() + len_, buf_.size() - len_)
            : use_az_ ? az_.read(buf_.data() + len_, buf_.size() - len_)
            : use_hz_ ? hz_.read(buf_.data() + len_, buf_.size() - len_)
            : gz_ ? gzread(gz_, buf_.data() + len_, (unsigned)std::min<size_t>(buf_.size() - len_, 1u << 30))
                  : (long)std::fread(buf_.data() + len_, 1, buf_.size() - len_, fp_);
        if (got < 0) return fail("Failed to read " + path_ + ": gzip stream error");
        if (got == 0) at_eof_ = true;
        // the kept partial line came out of the previous buffer: ASCII if that one was, as a whole
        buf_ascii_ = (len_ == 0 || buf_ascii_) && all_ascii(buf_.data() + len_, (size_t)got);
        len_ += (size_t)got;
      }
      if (at_eof_ && got == 0) {
        if (len_ > pos_) {                         // final line without a newline: BufRead::lines() yields it
          const size_t n = len_ - pos_;
          std::vector<uint8_t> last(buf_.begin() + pos_, buf_.begin() + pos_ + n);
          pos_ = len_ = 0;
          if (handle_line(last.data(), n, bases, offs, false)) return 1;
          if (chunk_full(max_reads, max_bases, bases, offs)) return 0;
        }
        *eof = true;
        return 0;
      }
    }
  }
  uint64_t line_count = 0, total_reads = 0, error_count = 0;

 private:
  template <class VB, class VO>
  bool chunk_full(uint64_t max_reads, uint64_t max_bases, const VB& bases, const VO& offs) const
  {
    const uint64_t n = offs.size() - 1;
    if (n == 0) return false;
    return n >= max_reads || (max_bases && bases.size() >= max_bases);   // aligner.rs:143 (+ GPU_CHUNK_SIZE_BASES)
  }
  template <class VB, class VO>
  int handle_line(const uint8_t* p, size_t n, VB& bases, VO& offs, bool terminated = true)
  {
    if (terminated && n && p[n - 1] == '\r') --n;  // lines() strips "\n" and "\r\n"; a last line without '\n' keeps its '\r'
    if (!buf_ascii_ && !valid_utf8(p, n)) {        // lines() yields Err for invalid UTF-8 (aligner.rs:155-163); a buffer that is
                                                   // ASCII as a whole (one wide pass when it was read) needs no check per line
      ++error_count;
      if (error_count <= 5) std::printf("    Warning: Error reading line %llu: stream did not contain valid UTF-8\n", (unsigned long long)line_count);
      if (error_count > 10) return fail("Too many read errors (>10), stopping at line " + std::to_string(line_count));
      return 0;
    }
    ++line_count;                                  // aligner.rs:136
    if (line_count % 4 == 2) {                     // aligner.rs:138: the sequence line of a 4-line record
      bases.insert(bases.end(), p, p + n);
      offs.push_back(bases.size());
      ++total_reads;
    }
    if (line_count % 1000000 == 0)                 // aligner.rs:151-153
      std::printf("    Debug: Read %llu lines, found %llu reads, current chunk size: %llu\n", (unsigned long long)line_count,
                  (unsigned long long)total_reads, (unsigned long long)(offs.size() - 1));
    return 0;
  }
  std::string path_;
  gzFile gz_ = nullptr; FILE* fp_ = nullptr;
  static bool all_ascii(const uint8_t* p, size_t n)
  {
    uint64_t acc = 0; size_t i = 0;
    for (; i + 32 <= n; i += 32) { uint64_t w[4]; std::memcpy(w, p + i, 32); acc |= w[0] | w[1] | w[2] | w[3]; }
    for (; i < n; ++i) acc |= p[i];
    return (acc & 0x8080808080808080ull) == 0;
  }
  bool buf_ascii_ = false;
  hgz::GunzipStream hz_; bool use_hz_ = false;
  hgz::AsyncGunzip az_; bool use_az_ = false;
  hgz::ParallelGunzip pz_; bool use_pz_ = false;
  std::vector<uint8_t> buf_;
  size_t pos_ = 0, len_ = 0;
  bool at_eof_ = false;
};

uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

std::string env_or(const char* k, const char* dflt) { const char* v = std::getenv(k); return v ? v : dflt; }

std::vector<std::string> wgs_files()
{
  // aligner.rs:184-204
  const std::string dir = env_or("WGS_DATA_DIR", "/path/to/wgs/data"), sample = env_or("WGS_SAMPLE_ID", "SAMPLE_ID");
  uint64_t lanes = 8, rpl = 2; std::string why;
  if (parse_usize(env_or("WGS_LANES", "8"), &lanes, &why)) lanes = 8;
  if (parse_usize(env_or("WGS_READS_PER_LANE", "2"), &rpl, &why)) rpl = 2;
  std::vector<std::string> files;
  for (uint64_t lane = 1; lane <= lanes; ++lane)
    for (uint64_t read = 1; read <= rpl; ++read) {
      char name[64];
      std::snprintf(name, sizeof name, "_L%03llu_R%llu_001.fastq.gz", (unsigned long long)lane, (unsigned long long)read);
      files.push_back(dir + "/" + sample + name);
    }
  return files;
}

std::string base_name(const std::string& p) { const size_t k = p.rfind('/'); return k == std::string::npos ? p : p.substr(k + 1); }

// ---- the reference genome the --full-wgs reads are scored against (this engine's pairing rule) ----
int load_reference(std::vector<uint8_t>& ref)
{
  const char* path = std::getenv("WGS_REFERENCE");
  if (path && *path) {
    gzFile g = gzopen(path, "rb");
    if (!g) return fail(std::string("Failed to open file ") + path + ": " + std::strerror(errno));
    std::vector<uint8_t> buf(4 << 20);
    bool header = false, line_start = true;
    for (;;) {
      const int got = gzread(g, buf.data(), (unsigned)buf.size());
      if (got < 0) { gzclose(g); return fail(std::string("Failed to read ") + path); }
      if (got == 0) break;
      for (int k = 0; k < got; ++k) {
        const uint8_t c = buf[k];
        if (line_start && c == '>') header = true;
        line_start = (c == '\n');
        if (c == '\n') { header = false; continue; }
        if (!header && c != '\r') ref.push_back(c);
      }
    }
    gzclose(g);
    if (ref.empty()) return fail(std::string("WGS_REFERENCE ") + path + " holds no bases");
    return 0;
  }
  uint64_t n = 16000000; std::string why;
  if (parse_usize(env_or("WGS_SYNTH_REFERENCE_BASES", "16000000"), &n, &why) || n == 0) n = 16000000;
  ref.resize(n);
  for (uint64_t k = 0; k < n; k += 32) {
    uint64_t x = splitmix64(0xB2F0ull + (k >> 5));
    for (uint64_t j = k; j < std::min(n, k + 32); ++j, x >>= 2) ref[j] = (uint8_t)"ACGT"[x & 3];
  }
  return 0;
}

struct FileOutcome { int rc = 0; std::string err; rsm_alignment_result res{}; };

// The run's checkpoint (aligner.rs:23-104): one entry per finished file, rewritten after every file.
struct CheckpointBook {
  std::mutex mu; std::string path, run_id; uint64_t total_files = 0;
  std::vector<rsm_file_checkpoint> files;
  void add(const rsm_file_checkpoint& fc)        // CheckpointState::add_file_result (aligner.rs:86-99)
  {
    std::lock_guard<std::mutex> lk(mu);
    files.erase(std::remove_if(files.begin(), files.end(), [&](const rsm_file_checkpoint& f) { return f.file_index == fc.file_index; }), files.end());
    files.push_back(fc);
    if (rsm_checkpoint_save(path.c_str(), run_id.c_str(), files.data(), (int)files.size(), total_files))
      std::printf("Warning: Failed to save checkpoint: %s\n", g_err.c_str());                    // aligner.rs:309
  }
  const rsm_file_checkpoint* completed(size_t index)                                               // is_file_completed (aligner.rs:101-103)
  {
    for (const auto& f : files) if (f.file_index == index && f.completed) return &f;
    return nullptr;
  }
};
CheckpointBook* g_book = nullptr;
std::atomic<uint64_t> g_gpu_call_us{0};              // wall time the scoring threads spent inside GPU calls (benchmark report)

void record_checkpoint(size_t index, const std::string& path, const FileOutcome& o)
{
  if (!g_book) return;
  rsm_file_checkpoint fc; std::memset(&fc, 0, sizeof fc);
  std::snprintf(fc.file_path, sizeof fc.file_path, "%s", path.c_str());
  fc.file_index = index; fc.score = o.res.score; fc.score64 = o.res.score64; fc.processing_time_ms = o.res.processing_time_ms;
  fc.total_bases = o.res.total_bases; fc.total_reads = o.res.total_reads; fc.completed = o.rc == 0;
  g_book->add(fc);
}

// One file of the --full-wgs loop (aligner.rs:261-339) in ref_compat mode: sequential, concat + self-align.
void process_one_file_compat(size_t index, size_t total, const std::string& file, uint64_t chunk_reads, uint64_t chunk_bases,
                             const rsm_gpu_device* dev, FileOutcome* out)
{
  std::printf("Processing file %zu/%zu: %s\n", index + 1, total, base_name(file).c_str());
  std::printf("    Using chunk size: %llu reads \n", (unsigned long long)chunk_reads);
  const auto t0 = std::chrono::steady_clock::now();
  int64_t total_score = 0; uint64_t processed_chunks = 0, total_bases = 0, total_reads = 0;
  FastqReader rd;
  int rc = rd.open(file);
  std::vector<uint8_t> bases; std::vector<uint64_t> offs;
  bool eof = false;
  while (rc == 0 && !eof) {
    rc = rd.next_chunk(chunk_reads ? chunk_reads : 1, chunk_bases, bases, offs, &eof);
    const uint64_t n = offs.size() - 1;
    if (rc != 0 || n == 0) break;
    int32_t sc = 0;
    const int crc = rsm_gpu_align_chunk_self(bases.data(), bases.size(), dev, &sc);     // aligner.rs:270-276
    total_bases += bases.size(); total_reads += n;
    if (crc == 0) {
      total_score += sc; ++processed_chunks;
      if (processed_chunks % 10 == 0)                               // aligner.rs:278-282
        std::printf("    Processed %llu chunks (%llu reads), current score: %lld\n", (unsigned long long)processed_chunks,
                    (unsigned long long)n, (long long)total_score);
    } else {
      std::printf("    Warning: Failed to align chunk %llu: %s\n", (unsigned long long)processed_chunks, g_err.c_str());   // aligner.rs:284-286
    }
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == 0) {
    const uint64_t cr = chunk_reads ? chunk_reads : 1;
    std::printf("    Processed %llu total reads in %llu chunks\n", (unsigned long long)rd.total_reads, (unsigned long long)((rd.total_reads + cr - 1) / cr));
    std::printf("    Total lines read: %llu\n", (unsigned long long)rd.line_count);
    std::printf("  File %zu complete: Score=%lld, Bases=%llu, Time: %.2f s \n", index + 1, (long long)total_score, (unsigned long long)total_bases, secs);
  } else {
    std::printf("  File %zu failed: %s\n", index + 1, g_err.c_str());
    out->rc = 1; out->err = "File " + std::to_string(index + 1) + " failed: " + g_err;
  }
  out->res.score64 = total_score; out->res.score = (int32_t)total_score;
  out->res.processing_time_ms = std::floor(secs * 1000.0);
  std::snprintf(out->res.gpu_device, sizeof out->res.gpu_device, "%s", dev->name);
  out->res.total_reads = total_reads; out->res.total_bases = total_bases;
  record_checkpoint(index, file, *out);                                                           // aligner.rs:298-310, :322-334
}

// ---- the --full-wgs pipeline in Smith-Waterman mode ----
// The reference inflates with one `zcat` child, parses on the main thread and blocks on every chunk
// (aligner.rs:111-148, :269-289, :527).  Here every file has its own inflate+parse thread that fills pinned chunk
// buffers (reads, offsets, and the window each read is scored against); every GPU has one consumer thread that takes
// whichever chunk of its files is ready and hands it to swb_score_batch_vs_reference, whose three stream lanes overlap
// the chunk's H2D, kernels and D2H.  Inflate of all files, PCIe and the SMs run concurrently.
struct WgsChunk {
  pinned_vector<uint8_t> bases; pinned_vector<uint64_t> offs, wstart; pinned_vector<uint32_t> wlen;
  pinned_vector<swb_result> res;
  bool ends_chunk = true;                      // last piece of a GPU_CHUNK_SIZE_READS/BASES chunk (progress accounting)
  // BGZF files: a segment of whole compressed blocks for swb_fastq_bgzf_score (the GPU inflates and parses)
  bool bgzf = false, final_segment = false;
  pinned_vector<uint8_t>* comp = nullptr; uint64_t comp_len = 0;     // a buffer of the consumer's CompPool while the segment is in flight
  uint64_t comp_off = 0;                                              // the segment's first block starts here in *comp
  const uint8_t* comp_data() const { return comp->data() + comp_off; }
  std::vector<swb_bgzf_block> blocks;
};

// Compressed bytes per BGZF segment.  The inflate kernel decodes one block per warp (32 warps/SM x 148 SMs = 4736 blocks
// at once), a segment should carry a few thousand blocks so the device is full: 112 MiB of the synthetic FASTQ is ~9470
// blocks, ~600 MB of text, ~1.9 M reads of 150 bp.  (Cutting segments to whole multiples of 4736 blocks was tried for files
// that compress differently and changes nothing measurable: the inflate of the next segment runs beside the scoring of
// the current one, so a half-empty last wave leaves no SM idle.)  Smaller files use segments of their own size (pinned
// memory is slow to allocate).
constexpr uint64_t kBgzfSegmentMax = 112ull << 20;
constexpr uint64_t kBgzfSegmentMulti = 32ull << 20;      // when the process drives more than one GPU (see wgs_device_pipeline)

// Is this a blocked-gzip file (first member carries a 'BC' extra field)?  SWB_GPU_INFLATE=0 keeps every file on the host.
bool file_is_bgzf(const std::string& path)
{
  if (const char* v = std::getenv("SWB_GPU_INFLATE")) if (std::string(v) == "0") return false;
  if (path.size() < 3 || path.compare(path.size() - 3, 3, ".gz") != 0) return false;
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) return false;
  uint8_t h[18];
  const size_t n = std::fread(h, 1, sizeof h, fp);
  std::fclose(fp);
  return n == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 && h[12] == 'B' && h[13] == 'C' &&
         h[14] == 2 && h[15] == 0;
}

// Whole BGZF blocks in buf[0, n): payload ranges + inflated sizes; returns the bytes consumed, or (uint64_t)-1 on a
// member that is not BGZF (the caller falls back to the host path).
uint64_t walk_bgzf(const uint8_t* buf, uint64_t n, std::vector<swb_bgzf_block>& blocks)
{
  uint64_t pos = 0;
  while (pos + 18 <= n) {
    const uint8_t* h = buf + pos;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return ~0ull;
    const uint64_t xlen = h[10] | ((uint64_t)h[11] << 8);
    if (pos + 12 + xlen > n) break;
    uint64_t bsize = ~0ull;
    for (uint64_t q = 12; q + 4 <= 12 + xlen;) {
      const uint64_t slen = h[q + 2] | ((uint64_t)h[q + 3] << 8);
      if (h[q] == 'B' && h[q + 1] == 'C' && slen == 2 && q + 6 <= 12 + xlen) bsize = h[q + 4] | ((uint64_t)h[q + 5] << 8);
      q += 4 + slen;
    }
    if (bsize == ~0ull || bsize + 1 < 12 + xlen + 8) return ~0ull;
    const uint64_t total = bsize + 1;
    if (pos + total > n) break;
    swb_bgzf_block b;
    b.in_off = pos + 12 + xlen; b.in_len = (uint32_t)(total - 12 - xlen - 8);
    b.out_len = (uint32_t)h[total - 4] | ((uint32_t)h[total - 3] << 8) | ((uint32_t)h[total - 2] << 16) | ((uint32_t)h[total - 1] << 24);
    if (b.out_len > 65536) return ~0ull;
    blocks.push_back(b);
    pos += total;
  }
  return pos;
}

// A reference-sized chunk (GPU_CHUNK_SIZE_READS can be millions of reads) moves through the pipeline in pieces of at most
// this many reads: pinned memory stays small (page-locking is slow) and inflate, PCIe and the SMs overlap inside a chunk.
constexpr uint64_t kWgsPieceReads = 16384;

struct DeviceGate { std::mutex mu; std::condition_variable cv; };   // wakes the consumer of one GPU

// Pinned staging buffers for compressed BGZF segments, shared by all files of one consumer: the consumer works on one
// segment at a time, so three buffers keep it fed however many files it serves (page-locking 112 MiB per file and per
// pipeline slot would cost seconds).
std::atomic<uint64_t> g_pin_us{0}, g_pin_bytes{0};     // SWB_STAMPS: time spent page-locking segment buffers (summed over threads)
struct CompPool {
  std::mutex mu; std::condition_variable cv;
  std::vector<std::unique_ptr<pinned_vector<uint8_t>>> all;
  std::vector<pinned_vector<uint8_t>*> free;
  size_t max_buffers = 3; uint64_t bytes = 0;
  pinned_vector<uint8_t>* acquire()
  {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      if (!free.empty()) { auto* b = free.back(); free.pop_back(); return b; }
      if (all.size() < max_buffers) {
        all.push_back(std::make_unique<pinned_vector<uint8_t>>());
        auto* b = all.back().get();
        lk.unlock();
        const auto t0 = std::chrono::steady_clock::now();
        b->resize(bytes);                           // the slow part, outside the lock
        g_pin_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        g_pin_bytes += bytes;
        return b;
      }
      cv.wait(lk);
    }
  }
  void release(pinned_vector<uint8_t>* b) { { std::lock_guard<std::mutex> lk(mu); free.push_back(b); } cv.notify_one(); }
};

struct WgsFile {
  size_t index = 0; std::string path;
  DeviceGate* gate = nullptr;
  std::deque<WgsChunk*> ready, spare;          // guarded by gate->mu
  bool closed = false; int rc = 0; std::string err;
  uint64_t total_reads = 0, line_count = 0;    // filled by the reader before it closes
  std::vector<std::unique_ptr<WgsChunk>> pool;
  std::chrono::steady_clock::time_point t0;
  // consumer side
  int64_t score = 0; uint64_t chunks = 0, bases = 0, reads = 0, chunk_reads_seen = 0;
  bool bgzf = false, gpu_path_failed = false;  // BGZF: inflated and parsed on the GPU; failed -> the file is redone on the host
  std::vector<uint8_t> carry; uint64_t lines = 0;
  uint64_t seg_bytes = 0; CompPool* comp_pool = nullptr;
  // several readers per file (wgs_bgzf_reader_thread): segment k may take its buffers once k-1 has, and is walked after k-1
  unsigned spare_threads = 0;                  // plain .gz: threads the box has for this file beside its parsing thread (FastqReader::open)
  swb_ctx* ctx = nullptr;                      // the consumer's context: readers make its device current before they page-lock buffers
  int fd = -1; uint64_t file_bytes = 0, n_segments = 0;
  unsigned n_readers = 1, live_readers = 0;
  uint64_t next_acquire = 0, next_walk = 0;    // guarded by gate->mu
  std::vector<uint8_t> left;                   // the partial block at the end of the last walked segment (owned by the walker in turn)
  int reader_rc = 0; std::string reader_err;   // first failure among the readers, guarded by gate->mu
};

// BGZF files: the readers only move compressed bytes and walk block headers; inflate + parse happen on the GPU.
// A file has n_readers of these threads (reader r takes segments r, r + n_readers, ...): the bytes of segment k are the
// fixed file range [k * seg_bytes, (k+1) * seg_bytes), so several pread()s of one file run at once -- one thread copying out
// of the page cache delivers 1-3 GB/s, a B200 inflates and scores ~3.7 GB/s of compressed FASTQ, and a box has more cores
// than files.  Only the cheap part is ordered: segment k takes its chunk and its pinned buffer after segment k-1 has taken
// its own (so the segment the consumer needs next is never the one left without a buffer), and its block headers are
// walked after k-1's, because the first block of k starts where the last whole block of k-1 ended; the bytes in between
// (`left`, less than one block) are copied in front of the range (kBgzfFront bytes of room).
constexpr uint64_t kBgzfFront = 128u << 10;

void wgs_bgzf_reader_thread(WgsFile* f, unsigned r)
{
  int rc = 0; std::string err;
  // Page-locking runs under the current device's context lock, and a new thread's current device is 0: without this the
  // readers of ALL GPUs pin under device 0's lock and its consumer's launches wait behind them (8 GPUs, 32 MiB segments:
  // 22 ms per call on device 0 against 9.5 ms on the other seven, profiles/wgs_8gpu_consumer_accounting_r02.txt).
  if (f->ctx) swb_bind_thread(f->ctx);
  const uint64_t S = f->seg_bytes;
  for (uint64_t k = r; k < f->n_segments && rc == 0; k += f->n_readers) {
    WgsChunk* c = nullptr;
    {
      std::unique_lock<std::mutex> lk(f->gate->mu);
      f->gate->cv.wait(lk, [&] { return (f->next_acquire == k && !f->spare.empty()) || f->gpu_path_failed || f->reader_rc; });
      if (f->gpu_path_failed || f->reader_rc) break;
      c = f->spare.front(); f->spare.pop_front();
    }
    bool push = false;
    try {
      c->bgzf = true; c->blocks.clear();
      c->comp = f->comp_pool->acquire();
      { std::lock_guard<std::mutex> lk(f->gate->mu); ++f->next_acquire; }
      f->gate->cv.notify_all();
      uint8_t* buf = c->comp->data() + kBgzfFront;
      uint64_t got = 0;
      while (got < S) {
        const ssize_t n = ::pread(f->fd, buf + got, (size_t)std::min<uint64_t>(S - got, 1u << 30), (off_t)(k * S + got));
        if (n < 0) { if (errno == EINTR) continue; rc = 1; err = "Failed to read file " + f->path + ": " + std::strerror(errno); break; }
        if (n == 0) break;
        got += (uint64_t)n;
      }
      {
        std::unique_lock<std::mutex> lk(f->gate->mu);
        f->gate->cv.wait(lk, [&] { return f->next_walk == k || f->gpu_path_failed || f->reader_rc; });
        if (f->gpu_path_failed || f->reader_rc) rc = rc ? rc : 3;
      }
      if (rc == 0) {                                            // this thread's turn: nobody else touches f->left
        const bool last = k + 1 == f->n_segments;
        const uint64_t nl = f->left.size();
        uint8_t* start = buf - nl;
        std::memcpy(start, f->left.data(), nl);
        const uint64_t have = nl + got;
        const uint64_t used = walk_bgzf(start, have, c->blocks);
        if (used == ~0ull || (last && used != have)) { rc = 2; err = "not a BGZF stream"; }      // -> host path
        else {
          f->left.assign(start + used, start + have);
          if (f->left.size() > kBgzfFront) { rc = 2; err = "not a BGZF stream"; }
          c->comp_off = kBgzfFront - nl; c->comp_len = used; c->final_segment = last;
          push = rc == 0;
        }
      }
    } catch (const std::bad_alloc&) { rc = 1; err = "out of pinned host memory"; }
    if (!push && c->comp) { f->comp_pool->release(c->comp); c->comp = nullptr; }
    std::lock_guard<std::mutex> lk(f->gate->mu);
    if (push) { f->ready.push_back(c); ++f->next_walk; } else f->spare.push_back(c);
    if (rc == 2) f->gpu_path_failed = true;
    else if (rc == 1 && !f->reader_rc) { f->reader_rc = 1; f->reader_err = err; }
    f->gate->cv.notify_all();
  }
  std::lock_guard<std::mutex> lk(f->gate->mu);
  if (--f->live_readers == 0) {
    if (f->fd >= 0) { ::close(f->fd); f->fd = -1; }
    f->rc = f->reader_rc; f->err = f->reader_err; f->closed = true;
  }
  f->gate->cv.notify_all();
}

void wgs_reader_thread(WgsFile* f, uint64_t chunk_reads, uint64_t chunk_bases, uint64_t ref_len, uint32_t window_len)
{
  if (f->ctx) swb_bind_thread(f->ctx);          // a chunk buffer that grows is page-locked by this thread: under its own GPU's context lock
  FastqReader rd;
  int rc = rd.open(f->path, f->spare_threads);
  std::string err = rc ? g_err : "";
  bool eof = false; uint64_t first = 0, in_chunk_reads = 0, in_chunk_bases = 0;
  while (rc == 0 && !eof) {
    WgsChunk* c = nullptr;
    {
      std::unique_lock<std::mutex> lk(f->gate->mu);
      f->gate->cv.wait(lk, [&] { return !f->spare.empty(); });
      c = f->spare.front(); f->spare.pop_front();
    }
    try {
      const uint64_t cr = chunk_reads ? chunk_reads : 1;
      const uint64_t want = std::min<uint64_t>(kWgsPieceReads, cr - in_chunk_reads);
      const uint64_t bcap = chunk_bases ? (chunk_bases > in_chunk_bases ? chunk_bases - in_chunk_bases : 1) : 0;
      rc = rd.next_chunk(want, bcap, c->bases, c->offs, &eof);
      if (rc) err = g_err;
      const uint64_t n = c->offs.size() - 1;
      in_chunk_reads += n; in_chunk_bases += c->bases.size();
      c->ends_chunk = eof || in_chunk_reads >= cr || (chunk_bases && in_chunk_bases >= chunk_bases);   // aligner.rs:143 (+ BASES)
      if (c->ends_chunk) { in_chunk_reads = 0; in_chunk_bases = 0; }
      c->wstart.resize(n); c->wlen.resize(n); c->res.resize(n);
      const uint32_t w = (uint32_t)std::min<uint64_t>(window_len, ref_len);
      for (uint64_t k = 0; k < n; ++k) {       // this engine's pairing rule: a deterministic window per read
        const uint64_t g = ((uint64_t)f->index << 40) + first + k;
        c->wstart[k] = splitmix64(g ^ 0xB202ull) % (ref_len - w + 1);
        c->wlen[k] = w;
      }
      first += n;
    } catch (const std::bad_alloc&) { rc = 1; err = "out of pinned host memory"; }
    std::lock_guard<std::mutex> lk(f->gate->mu);
    if (rc == 0 && c->offs.size() > 1) f->ready.push_back(c); else f->spare.push_back(c);
    f->gate->cv.notify_all();
  }
  std::lock_guard<std::mutex> lk(f->gate->mu);
  f->total_reads = rd.total_reads; f->line_count = rd.line_count;
  f->rc = rc; f->err = err; f->closed = true;
  f->gate->cv.notify_all();
}

void wgs_finish_file(WgsFile* f, size_t total, uint64_t chunk_reads, const rsm_gpu_device* dev, FileOutcome* out)
{
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - f->t0).count();
  if (f->rc == 0) {
    const uint64_t cr = chunk_reads ? chunk_reads : 1;
    if (f->bgzf) { f->total_reads = f->reads; f->line_count = f->lines; }
    std::printf("    Processed %llu total reads in %llu chunks\n", (unsigned long long)f->total_reads, (unsigned long long)((f->total_reads + cr - 1) / cr));
    std::printf("    Total lines read: %llu\n", (unsigned long long)f->line_count);
    std::printf("  File %zu complete: Score=%lld, Bases=%llu, Time: %.2f s \n", f->index + 1, (long long)f->score, (unsigned long long)f->bases, secs);
  } else {
    std::printf("  File %zu failed: %s\n", f->index + 1, f->err.c_str());
    out->rc = 1; out->err = "File " + std::to_string(f->index + 1) + " failed: " + f->err;
  }
  (void)total;
  out->res.score64 = f->score; out->res.score = (int32_t)f->score;
  out->res.processing_time_ms = std::floor(secs * 1000.0);            // as_millis() as f64
  std::snprintf(out->res.gpu_device, sizeof out->res.gpu_device, "%s", dev->name);
  out->res.total_reads = f->reads; out->res.total_bases = f->bases;
  record_checkpoint(f->index, f->path, *out);                                                     // aligner.rs:298-310, :322-334
}

// All files of one GPU: readers in parallel, one consumer (this thread) scoring whatever is ready.
void wgs_device_pipeline(const std::vector<size_t>& mine, const std::vector<std::string>& files, size_t total, uint64_t chunk_reads,
                         uint64_t chunk_bases, const rsm_gpu_device* dev, swb_ctx* ctx, uint64_t ref_len, uint32_t window_len,
                         std::vector<FileOutcome>* outcomes, unsigned readers_per_file, size_t n_devices, unsigned spare_threads)
{
  DeviceGate gate;
  CompPool pool;
  size_t bgzf_files = 0;
  if (const char* v = std::getenv("SWB_BGZF_POOL")) { const long n = std::atol(v); if (n >= 2 && n <= 16) pool.max_buffers = (size_t)n; }
  std::vector<std::unique_ptr<WgsFile>> fs;
  const size_t depth = 3;
  for (size_t i : mine) {
    auto f = std::make_unique<WgsFile>();
    f->index = i; f->path = files[i]; f->gate = &gate; f->t0 = std::chrono::steady_clock::now(); f->ctx = ctx;
    f->spare_threads = spare_threads;
    f->bgzf = file_is_bgzf(files[i]);
    if (f->bgzf) {
      f->comp_pool = &pool;
      // Page-locking the segment buffers is the expensive part of start-up, and a process that drives several GPUs pays
      // more per byte (the pages are mapped for every device, all threads pin under one mm lock: 2.8 GB took 12 thread-seconds
      // on eight GPUs, 0.35 GB 0.3 s on one): smaller segments there (8 GPUs, 128 M reads: 76 -> 133-154 M reads/s in the
      // pipeline; one GPU alone loses 2 % with them and keeps the large ones).
      const uint64_t seg_max = n_devices > 1 ? kBgzfSegmentMulti : kBgzfSegmentMax;
      f->seg_bytes = seg_max;
      if (FILE* fp = std::fopen(files[i].c_str(), "rb")) {
        if (std::fseek(fp, 0, SEEK_END) == 0) {
          const long sz = std::ftell(fp);
          if (sz > 0) f->seg_bytes = std::min<uint64_t>(seg_max, (((uint64_t)sz + 1) / 2 + (1u << 16) + 4095) & ~4095ull);   // at least two segments
        }
        std::fclose(fp);
      }
      if (const char* v = std::getenv("SWB_BGZF_SEGMENT_MB")) { const long mb = std::atol(v); if (mb > 0) f->seg_bytes = (uint64_t)mb << 20; }
      if (const char* v = std::getenv("SWB_BGZF_SEGMENT_KB")) { const long kb = std::atol(v); if (kb >= 128) f->seg_bytes = (uint64_t)kb << 10; }   // tests: many segments of a small file
      pool.bytes = std::max<uint64_t>(pool.bytes, kBgzfFront + f->seg_bytes + 64);
      f->fd = ::open(files[i].c_str(), O_RDONLY);
      if (f->fd < 0) f->err = "Failed to open file " + f->path + ": " + std::strerror(errno);      // (errno is this call's only here)
      struct stat sb;
      if (f->fd >= 0 && ::fstat(f->fd, &sb) == 0) f->file_bytes = (uint64_t)sb.st_size;
      f->n_segments = std::max<uint64_t>(1, (f->file_bytes + f->seg_bytes - 1) / f->seg_bytes);
      f->n_readers = (unsigned)std::min<uint64_t>(std::max(1u, readers_per_file), f->n_segments);
      ++bgzf_files;
    }
    for (size_t d = 0; d < (f->bgzf ? 1 + f->n_readers : depth); ++d) {
      f->pool.push_back(std::make_unique<WgsChunk>());
      WgsChunk* c = f->pool.back().get();
      if (!f->bgzf) {
        // one pinned allocation per buffer instead of a doubling series (page-locking is slow and serialised in the driver)
        const uint64_t nr = std::min<uint64_t>(chunk_reads ? chunk_reads : 1, kWgsPieceReads);
        const uint64_t nb = chunk_bases ? std::min<uint64_t>(chunk_bases + 1024, nr * 152) : nr * 152;
        try {
          c->bases.reserve(std::min<uint64_t>(nb, 600ull << 20)); c->offs.reserve(nr + 1); c->wstart.reserve(nr); c->wlen.reserve(nr); c->res.reserve(nr);
        } catch (const std::bad_alloc&) {}
      }
      f->spare.push_back(c);
    }
    std::printf("Processing file %zu/%zu: %s\n", i + 1, total, base_name(files[i]).c_str());
    std::printf("    Using chunk size: %llu reads \n", (unsigned long long)chunk_reads);
    if (f->bgzf) std::printf("    Blocked gzip (BGZF): inflate + FASTQ parsing on the GPU\n");
    fs.push_back(std::move(f));
  }
  // pinned segment buffers of this GPU: the one being scored, the one being prefetched, and one per reader
  if (!std::getenv("SWB_BGZF_POOL") && readers_per_file > 1) pool.max_buffers = std::min<size_t>(16, 2 + bgzf_files * readers_per_file);
  std::vector<std::thread> readers;
  for (auto& f : fs) {
    if (f->bgzf) {
      if (f->fd < 0) {                                       // unreadable: reported like the host reader would
        std::lock_guard<std::mutex> lk(gate.mu);
        f->rc = 1; f->closed = true;
        continue;
      }
      f->live_readers = f->n_readers;
      for (unsigned r = 0; r < f->n_readers; ++r) readers.emplace_back(wgs_bgzf_reader_thread, f.get(), r);
    }
    else readers.emplace_back(wgs_reader_thread, f.get(), chunk_reads, chunk_bases, ref_len, window_len);
  }
  std::vector<uint8_t> carry_out(1 << 20);
  std::vector<WgsFile*> active;
  for (auto& f : fs) active.push_back(f.get());
  bool abort_run = false;
  size_t rr = 0;
  uint64_t st_wait_us = 0, st_call_us = 0, st_segments = 0, st_prefetched = 0;      // SWB_STAMPS: where this consumer's time went
  const auto st_t0 = std::chrono::steady_clock::now();
  WgsFile* pre_f = nullptr; WgsChunk* pre_c = nullptr;             // the BGZF segment already being inflated (prefetch)
  while (!active.empty()) {
    WgsFile* f = nullptr; WgsChunk* c = nullptr; bool finished = false;
    WgsFile* nf = nullptr; WgsChunk* nc = nullptr;
    {
      std::unique_lock<std::mutex> lk(gate.mu);
      if (pre_c) { f = pre_f; c = pre_c; f->ready.pop_front(); pre_f = nullptr; pre_c = nullptr; }      // it is the front of its file's queue
      for (; !f;) {
        for (size_t k = 0; k < active.size() && !f; ++k) {
          WgsFile* g = active[(rr + k) % active.size()];
          if (!g->ready.empty()) { f = g; c = g->ready.front(); g->ready.pop_front(); }
          else if (g->closed) { f = g; finished = true; }
        }
        if (f) break;
        const auto tw0 = std::chrono::steady_clock::now();
        gate.cv.wait(lk);
        st_wait_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - tw0).count();
      }
      ++rr;
      if (!finished && c->bgzf && !f->gpu_path_failed)            // the next BGZF segment, whichever file has one ready
        for (size_t k = 0; k < active.size() && !nc; ++k) {
          WgsFile* g = active[(rr + k) % active.size()];
          if (g->bgzf && !g->gpu_path_failed && !g->ready.empty() && g->ready.front()->bgzf) { nf = g; nc = g->ready.front(); }
        }
    }
    if (nc && !abort_run && swb_fastq_bgzf_prefetch(ctx, nc->comp_data(), nc->comp_len, nc->blocks.data(), nc->blocks.size()) == 0) {
      pre_f = nf; pre_c = nc;                                      // copied in and inflated on a second stream while `c` is scored
      ++st_prefetched;
    }
    if (finished && f->bgzf && f->gpu_path_failed && f->rc == 0 && !abort_run) {
      // the GPU path declined the file (not really BGZF, a corrupt block, non-ASCII text, a giant record): redo it with
      // zlib and the line reader, in this thread
      std::printf("    GPU FASTQ path not applicable to %s: falling back to the host reader\n", base_name(f->path).c_str());
      f->score = 0; f->reads = f->bases = f->chunks = f->lines = 0; f->bgzf = false;
      FastqReader rd;
      int rc = rd.open(f->path);
      std::vector<uint8_t> hb; std::vector<uint64_t> ho, hs; std::vector<uint32_t> hl; std::vector<swb_result> hr;
      bool eof = false;
      while (rc == 0 && !eof) {
        rc = rd.next_chunk(std::min<uint64_t>(chunk_reads ? chunk_reads : 1, 1u << 20), chunk_bases, hb, ho, &eof);
        const uint64_t n = ho.size() - 1;
        if (rc || n == 0) break;
        hs.resize(n); hl.resize(n); hr.resize(n);
        const uint32_t w = (uint32_t)std::min<uint64_t>(window_len, ref_len);
        for (uint64_t k = 0; k < n; ++k) { hs[k] = splitmix64((((uint64_t)f->index << 40) + f->reads + k) ^ 0xB202ull) % (ref_len - w + 1); hl[k] = w; }
        if (swb_score_batch_vs_reference(ctx, hb.data(), ho.data(), n, hs.data(), hl.data(), hr.data())) { rc = 1; g_err = swb_last_error(); break; }
        for (uint64_t k = 0; k < n; ++k) f->score += hr[k].score;
        f->reads += n; f->bases += hb.size(); ++f->chunks;
      }
      f->total_reads = rd.total_reads; f->line_count = rd.line_count;
      if (rc) { f->rc = 1; f->err = g_err; }
    }
    if (finished) {
      wgs_finish_file(f, total, chunk_reads, dev, &(*outcomes)[f->index]);
      if (f->rc) abort_run = true;                                   // aligner.rs:336: a failed file aborts the run
      active.erase(std::find(active.begin(), active.end(), f));
      continue;
    }
    if (c->bgzf) {                                                   // one segment of compressed blocks: everything on the GPU
      if (!f->gpu_path_failed && !abort_run) {
        int64_t ssum = 0; uint64_t nr = 0, nb = 0, nl = 0, ncarry = 0; int status = 0;
        const uint32_t w = (uint32_t)std::min<uint64_t>(window_len, ref_len);
        const auto tc0 = std::chrono::steady_clock::now();
        const int crc = swb_fastq_bgzf_score(ctx, c->comp_data(), c->comp_len, c->blocks.data(), c->blocks.size(), f->carry.data(), f->carry.size(),
                                             c->final_segment ? 1 : 0, f->index, f->reads, w, &ssum, &nr, &nb, &nl, carry_out.data(),
                                             carry_out.size(), &ncarry, &status);
        const uint64_t call_us = (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - tc0).count();
        g_gpu_call_us += call_us; st_call_us += call_us; ++st_segments;
        if (crc != 0 || status != 0) {
          if (crc) std::printf("    Warning: GPU FASTQ path failed on %s: %s\n", base_name(f->path).c_str(), swb_last_error());
          std::lock_guard<std::mutex> lk(gate.mu);
          f->gpu_path_failed = true;
        } else {
          const uint64_t cr = chunk_reads ? chunk_reads : 1;
          const uint64_t before = f->reads / cr;
          f->score += ssum; f->reads += nr; f->bases += nb; f->lines += nl;
          f->carry.assign(carry_out.begin(), carry_out.begin() + ncarry);
          f->chunks = f->reads / cr;
          if (f->chunks / 10 > before / 10)                            // aligner.rs:278-282, in units of GPU_CHUNK_SIZE_READS
            std::printf("    Processed %llu chunks (%llu reads), current score: %lld\n", (unsigned long long)(f->chunks / 10 * 10),
                        (unsigned long long)cr, (long long)f->score);
        }
      }
      if (c->comp) {
        if (f->gpu_path_failed || abort_run) swb_fastq_bgzf_cancel(ctx, c->comp_data());   // dropped, perhaps after it was prefetched
        pool.release(c->comp); c->comp = nullptr;
      }
      std::lock_guard<std::mutex> lk(gate.mu);
      f->spare.push_back(c);
      gate.cv.notify_all();
      continue;
    }
    const uint64_t n = c->offs.size() - 1;
    const auto th0 = std::chrono::steady_clock::now();
    int crc = abort_run ? 0 : swb_score_batch_vs_reference(ctx, c->bases.data(), c->offs.data(), n, c->wstart.data(), c->wlen.data(), c->res.data());
    g_gpu_call_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - th0).count();
    if (crc == 0 && !abort_run) {
      int64_t cs = 0;
      for (uint64_t k = 0; k < n; ++k) cs += c->res[k].score;
      f->score += cs; f->chunk_reads_seen += n;
      if (c->ends_chunk) {
        ++f->chunks;
        if (f->chunks % 10 == 0)                                     // aligner.rs:278-282
          std::printf("    Processed %llu chunks (%llu reads), current score: %lld\n", (unsigned long long)f->chunks,
                      (unsigned long long)f->chunk_reads_seen, (long long)f->score);
        f->chunk_reads_seen = 0;
      }
    } else if (crc) {
      std::printf("    Warning: Failed to align chunk %llu: %s\n", (unsigned long long)f->chunks, swb_last_error());   // aligner.rs:284-286
    }
    f->bases += c->bases.size(); f->reads += n;
    std::lock_guard<std::mutex> lk(gate.mu);
    f->spare.push_back(c);
    gate.cv.notify_all();
  }
  for (auto& t : readers) t.join();
  if (std::getenv("SWB_STAMPS") && st_segments)
    std::fprintf(stderr, "[wgs] device consumer: %.3f s in all, %.3f s waiting for a segment, %.3f s inside %llu GPU calls (%.2f ms each), %llu prefetched\n",
                 std::chrono::duration<double>(std::chrono::steady_clock::now() - st_t0).count(), st_wait_us / 1e6, st_call_us / 1e6,
                 (unsigned long long)st_segments, st_call_us / 1e3 / (double)st_segments, (unsigned long long)st_prefetched);
}

void load_dotenv()
{
  // dotenv::dotenv().ok() (main.rs:50): KEY=VALUE lines of ./.env, existing variables win
  FILE* f = std::fopen(".env", "r");
  if (!f) return;
  char line[4096];
  while (std::fgets(line, sizeof line, f)) {
    std::string s(line);
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r' || s.back() == ' ')) s.pop_back();
    size_t a = 0; while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) ++a;
    if (a >= s.size() || s[a] == '#') continue;
    if (s.compare(a, 7, "export ") == 0) a += 7;
    const size_t eq = s.find('=', a);
    if (eq == std::string::npos) continue;
    std::string k = s.substr(a, eq - a), v = s.substr(eq + 1);
    while (!k.empty() && k.back() == ' ') k.pop_back();
    if (v.size() >= 2 && ((v.front() == '"' && v.back() == '"') || (v.front() == '\'' && v.back() == '\''))) v = v.substr(1, v.size() - 2);
    setenv(k.c_str(), v.c_str(), 0);
  }
  std::fclose(f);
}

}  // namespace

extern "C" {

const char* rsm_last_error(void) { return g_err.c_str(); }

int rsm_is_gpu_available(void) { return swb_device_count() > 0 ? 1 : 0; }

int rsm_get_gpu_devices(rsm_gpu_device* out, int cap)
{
  const int n = swb_device_count();
  for (int i = 0; i < n && i < cap; ++i) {
    double gb = 0; int wg = 1024;
    std::memset(&out[i], 0, sizeof out[i]);
    if (swb_device_info(i, out[i].name, sizeof out[i].name, &gb, &wg) != 0) std::snprintf(out[i].name, sizeof out[i].name, "Unknown");
    out[i].memory_gb = (float)gb; out[i].max_work_group_size = (uint64_t)wg; out[i].ordinal = i;
  }
  return n;
}

int rsm_get_chunk_size_reads(uint64_t* out)
{
  const char* v = std::getenv("GPU_CHUNK_SIZE_READS");
  if (!v) return fail("GPU_CHUNK_SIZE_READS not set in .env file");                                  // aligner.rs:11
  std::string why;
  if (parse_usize(v, out, &why)) return fail(std::string("Invalid GPU_CHUNK_SIZE_READS value '") + v + "': " + why);   // aligner.rs:14
  return 0;
}

int rsm_get_chunk_size_bases(uint64_t* out)
{
  *out = 0;
  const char* v = std::getenv("GPU_CHUNK_SIZE_BASES");
  if (!v) return 0;
  std::string why;
  if (parse_usize(v, out, &why)) return fail(std::string("Invalid GPU_CHUNK_SIZE_BASES value '") + v + "': " + why);
  return 0;
}

int rsm_process_fastq_file_in_chunks(const char* filepath, uint64_t chunk_size_reads, rsm_chunk_fn processor, void* user)
{
  uint64_t max_bases = 0;
  if (rsm_get_chunk_size_bases(&max_bases)) return 1;
  FastqReader rd;
  // one file: every other core may inflate while this thread parses
  if (rd.open(filepath, std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()) - 1, kMaxInflateThreads))) return 1;
  std::vector<uint8_t> bases; std::vector<uint64_t> offs;
  bool eof = false;
  while (!eof) {
    if (rd.next_chunk(chunk_size_reads ? chunk_size_reads : 1, max_bases, bases, offs, &eof)) return 1;
    if (offs.size() > 1 && processor(user, bases.data(), offs.data(), offs.size() - 1) != 0) return 1;   // processor(&chunk)?
  }
  const uint64_t cr = chunk_size_reads ? chunk_size_reads : 1;
  std::printf("    Processed %llu total reads in %llu chunks\n", (unsigned long long)rd.total_reads, (unsigned long long)((rd.total_reads + cr - 1) / cr));
  std::printf("    Total lines read: %llu\n", (unsigned long long)rd.line_count);
  if (rd.error_count) std::printf("    Total read errors: %llu\n", (unsigned long long)rd.error_count);
  return 0;
}

static int count_cb(void* user, const uint8_t*, const uint64_t* offs, uint64_t n) { *(uint64_t*)user += offs[n]; return 0; }

// Test hook: a whole .gz file through the host gzip reader (hgz::GunzipStream, or gzread with use_zlib) in read() calls of
// read_cap bytes.  *n = bytes delivered (the first out_cap of them are in out), *failed = 1 when the reader reported corrupt data.
int rsm_debug_gunzip(const char* path, uint64_t read_cap, int use_zlib, uint8_t* out, uint64_t out_cap, uint64_t* n, int* failed)
{
  if (!path || !n || !failed || read_cap == 0) return fail("rsm_debug_gunzip: bad argument");
  *n = 0; *failed = 0;
  read_cap = std::min<uint64_t>(read_cap, 1ull << 28);               // a test hook: no gigabyte read buffers
  std::vector<uint8_t> buf(read_cap);
  auto deliver = [&](long got) {
    if (*n < out_cap && out) std::memcpy(out + *n, buf.data(), (size_t)std::min<uint64_t>((uint64_t)got, out_cap - *n));
    *n += (uint64_t)got;
  };
  if (use_zlib) {
    gzFile g = gzopen(path, "rb");
    if (!g) return fail(std::string("Failed to open file ") + path);
    gzbuffer(g, 1 << 20);
    for (;;) {
      const long got = gzread(g, buf.data(), (unsigned)std::min<uint64_t>(read_cap, 1u << 30));
      if (got < 0) { *failed = 1; break; }
      if (got == 0) break;
      deliver(got);
    }
    gzclose(g);
    return 0;
  }
  hgz::GunzipStream gs;
  if (!gs.open(path)) return fail(std::string("Failed to open file ") + path);
  for (;;) {
    const long got = gs.read(buf.data(), (size_t)read_cap);
    if (got < 0) { *failed = 1; g_err = gs.error(); break; }
    if (got == 0) break;
    deliver(got);
  }
  return 0;
}

// Test hook: the same through hgz::ParallelGunzip (host_pgunzip.h) with `threads` decoder threads and chunks of chunk_bytes
// compressed bytes (0 = the default, 1 MiB).  *parallel = 0 when the file went to the serial reader (too small, not gzip, not a
// regular file), *accepted / *serial_stretches = chunks taken from the workers / stretches the caller's thread decoded itself.
int rsm_debug_pgunzip(const char* path, uint64_t read_cap, unsigned threads, uint64_t chunk_bytes, uint8_t* out, uint64_t out_cap, uint64_t* n,
                      int* failed, int* parallel, uint64_t* accepted, uint64_t* serial_stretches)
{
  if (!path || !n || !failed || read_cap == 0) return fail("rsm_debug_pgunzip: bad argument");
  *n = 0; *failed = 0;
  try {
    read_cap = std::min<uint64_t>(read_cap, 1ull << 28);
    std::vector<uint8_t> buf(read_cap);
    hgz::ParallelGunzip pz;
    pz.set_threads(threads);
    if (chunk_bytes) pz.set_chunk_bytes(chunk_bytes);
    if (!pz.open(path)) return fail(std::string("Failed to open file ") + path);
    for (;;) {
      const long got = pz.read(buf.data(), (size_t)read_cap);
      if (got < 0) { *failed = 1; g_err = pz.error(); break; }
      if (got == 0) break;
      if (*n < out_cap && out) std::memcpy(out + *n, buf.data(), (size_t)std::min<uint64_t>((uint64_t)got, out_cap - *n));
      *n += (uint64_t)got;
    }
    if (parallel) *parallel = pz.parallel() ? 1 : 0;
    if (accepted) *accepted = pz.chunks_accepted();
    if (serial_stretches) *serial_stretches = pz.serial_stretches();
  } catch (const std::exception& e) { return fail(std::string("rsm_debug_pgunzip: ") + e.what()); }
  return 0;
}

// Test hook (no GPU): runs the BGZF readers of the --full-wgs driver on one file -- `readers` threads, segments of seg_bytes,
// `pool_buffers` segment buffers -- against a consumer that only takes the segments in order, and reports what it saw: the
// blocks, their inflated bytes, and an FNV-1a hash over every block's compressed payload and inflated size in stream order
// (identical for every reader count / segment size / pool size, or the hand-off protocol is broken).  status: 0 ok, 2 the
// file is not BGZF (the driver would fall back to the host reader), 1 error (rsm_last_error()).
int rsm_debug_bgzf_segments(const char* path, unsigned readers, uint64_t seg_bytes, unsigned pool_buffers, uint64_t* n_segments,
                            uint64_t* n_blocks, uint64_t* text_bytes, uint64_t* hash, int* status)
{
  if (!path || !n_segments || !n_blocks || !text_bytes || !hash || !status) return fail("rsm_debug_bgzf_segments: null pointer");
  *n_segments = *n_blocks = *text_bytes = 0; *hash = 1469598103934665603ull; *status = 0;
  if (seg_bytes < (128u << 10)) return fail("rsm_debug_bgzf_segments: segments of at least 128 KiB");
  struct PlainGuard { PlainGuard() { ++g_plain_buffers; } ~PlainGuard() { --g_plain_buffers; } } plain;
  DeviceGate gate; CompPool pool;
  pool.max_buffers = std::max(2u, pool_buffers); pool.bytes = kBgzfFront + seg_bytes + 64;
  WgsFile f;
  f.path = path; f.gate = &gate; f.bgzf = true; f.comp_pool = &pool; f.seg_bytes = seg_bytes;
  f.fd = ::open(path, O_RDONLY);
  if (f.fd < 0) return fail(std::string("Failed to open file ") + path + ": " + std::strerror(errno));
  struct stat sb;
  if (::fstat(f.fd, &sb) == 0) f.file_bytes = (uint64_t)sb.st_size;
  f.n_segments = std::max<uint64_t>(1, (f.file_bytes + seg_bytes - 1) / seg_bytes);
  f.n_readers = (unsigned)std::min<uint64_t>(std::max(1u, readers), f.n_segments);
  for (unsigned d = 0; d < 1 + f.n_readers; ++d) { f.pool.push_back(std::make_unique<WgsChunk>()); f.spare.push_back(f.pool.back().get()); }
  f.live_readers = f.n_readers;                                        // all of them before the first one runs: the last to leave closes the file
  std::vector<std::thread> th;
  try {
    for (unsigned r = 0; r < f.n_readers; ++r) th.emplace_back(wgs_bgzf_reader_thread, &f, r);
  } catch (const std::exception& e) {                                  // no exception leaves through the C ABI
    {
      std::lock_guard<std::mutex> lk(gate.mu);
      f.reader_rc = 1; f.reader_err = std::string("cannot start a reader thread: ") + e.what();      // the ones that run stop at their next step
      f.live_readers -= f.n_readers - (unsigned)th.size();
      if (th.empty()) { f.rc = 1; f.err = f.reader_err; f.closed = true; ::close(f.fd); f.fd = -1; }
    }
    gate.cv.notify_all();
  }
  for (;;) {
    WgsChunk* c = nullptr;
    {
      std::unique_lock<std::mutex> lk(gate.mu);
      gate.cv.wait(lk, [&] { return !f.ready.empty() || f.closed; });
      if (f.ready.empty()) break;
      c = f.ready.front(); f.ready.pop_front();
    }
    ++*n_segments;
    const uint8_t* base = c->comp_data();
    for (const swb_bgzf_block& b : c->blocks) {
      ++*n_blocks; *text_bytes += b.out_len;
      uint64_t h = *hash;
      for (uint32_t k = 0; k < b.in_len; ++k) { h ^= base[b.in_off + k]; h *= 1099511628211ull; }
      h ^= b.out_len; h *= 1099511628211ull;
      *hash = h;
    }
    pool.release(c->comp); c->comp = nullptr;
    std::lock_guard<std::mutex> lk(gate.mu);
    f.spare.push_back(c);
    gate.cv.notify_all();
  }
  for (auto& t : th) t.join();
  if (f.gpu_path_failed) *status = 2;
  if (f.rc) { *status = 1; return fail(f.err); }
  return 0;
}

int rsm_count_bases_in_fastq(const char* filepath, uint64_t* out)
{
  uint64_t chunk = 0;
  if (rsm_get_chunk_size_reads(&chunk)) return 1;                    // aligner.rs:538
  *out = 0;
  return rsm_process_fastq_file_in_chunks(filepath, chunk, count_cb, out);
}

int rsm_gpu_align_ex(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, const rsm_gpu_device* dev, swb_result* out)
{
  out->score = 0; out->end_i = -1; out->end_j = -1;
  const uint64_t len = std::min(n1, n2);
  if (len == 0) return 0;                                            // aligner.rs:413-416
  swb_ctx* c = context_for(dev ? dev->ordinal : 0);
  if (!c) return 1;
  std::lock_guard<std::mutex> use(g_use_mu[dev ? dev->ordinal : 0]);   // gpu.rs:13-14: callers of the shared context take turns
  if (ref_compat_mode()) {
    int32_t v = 0;
    const uint32_t wg = dev ? (uint32_t)std::min<uint64_t>(dev->max_work_group_size, 0xffffffffull) : 1024u;
    if (swb_ref_compat_align(c, s1, n1, s2, n2, wg, &v)) return fail(swb_last_error());
    out->score = v;
    return 0;
  }
  // the size guard of aligner.rs:436-456, restated for a quadratic kernel: bound the DP cells of one call
  uint64_t max_cells = 400000000000ull; std::string why;
  if (const char* v = std::getenv("SWB_MAX_CELLS")) parse_usize(v, &max_cells, &why);
  if ((long double)n1 * (long double)n2 > (long double)max_cells)
    return fail("Sequence too large (" + std::to_string(len) + " bytes): " + std::to_string(n1) + " x " + std::to_string(n2) +
                " DP cells exceed SWB_MAX_CELLS=" + std::to_string(max_cells));
  if (swb_score_pair(c, s1, n1, s2, n2, out)) return fail(swb_last_error());
  return 0;
}

int rsm_gpu_align(const uint8_t* s1, uint64_t n1, const uint8_t* s2, uint64_t n2, const rsm_gpu_device* dev, int32_t* score)
{
  swb_result r;
  if (rsm_gpu_align_ex(s1, n1, s2, n2, dev, &r)) return 1;
  *score = r.score;
  return 0;
}

int rsm_gpu_align_chunk_self(const uint8_t* chunk, uint64_t n, const rsm_gpu_device* dev, int32_t* score)
{
  if (n < 1000) { *score = 0; return 0; }                            // aligner.rs:366-368
  return rsm_gpu_align(chunk, n, chunk, n, dev, score);              // aligner.rs:372
}

int rsm_gpu_align_pair(const char* file1, const char* file2, const rsm_gpu_device* dev, rsm_alignment_result* out)
{
  uint64_t bases1 = 0, bases2 = 0, chunk = 0, chunk_bases = 0;
  if (rsm_count_bases_in_fastq(file1, &bases1) || rsm_count_bases_in_fastq(file2, &bases2)) return 1;   // aligner.rs:378-379
  std::printf("Loaded %llu bases from %s\n", (unsigned long long)bases1, file1);
  std::printf("Loaded %llu bases from %s\n", (unsigned long long)bases2, file2);
  if (rsm_get_chunk_size_reads(&chunk) || rsm_get_chunk_size_bases(&chunk_bases)) return 1;
  const auto t0 = std::chrono::steady_clock::now();
  int64_t total = 0; uint64_t reads = 0;
  std::vector<uint8_t> b1, b2; std::vector<uint64_t> o1, o2;
  if (ref_compat_mode()) {
    // aligner.rs:390-398 literally: every chunk of file1 against every chunk of file2 (file2 re-read each time)
    FastqReader r1; if (r1.open(file1)) return 1;
    bool eof1 = false;
    while (!eof1) {
      if (r1.next_chunk(chunk ? chunk : 1, chunk_bases, b1, o1, &eof1)) return 1;
      if (o1.size() <= 1) continue;
      FastqReader r2; if (r2.open(file2)) return 1;
      bool eof2 = false;
      while (!eof2) {
        if (r2.next_chunk(chunk ? chunk : 1, chunk_bases, b2, o2, &eof2)) return 1;
        if (o2.size() <= 1) continue;
        int32_t s = 0;
        if (rsm_gpu_align(b1.data(), b1.size(), b2.data(), b2.size(), dev, &s)) return 1;
        total += s;
      }
      reads += o1.size() - 1;
    }
  } else {
    // Smith-Waterman mode: read k of file1 against read k of file2 (mates), one batch per chunk
    swb_ctx* c = context_for(dev ? dev->ordinal : 0);
    if (!c) return 1;
    std::lock_guard<std::mutex> use(g_use_mu[dev ? dev->ordinal : 0]);
    const unsigned spare = std::min<unsigned>((std::max(2u, std::thread::hardware_concurrency()) - 2) / 2, kMaxInflateThreads);   // two files, this thread
    FastqReader r1, r2; if (r1.open(file1, spare) || r2.open(file2, spare)) return 1;
    bool eof1 = false, eof2 = false; std::vector<swb_result> res;
    while (!eof1 && !eof2) {
      if (r1.next_chunk(chunk ? chunk : 1, 0, b1, o1, &eof1) || r2.next_chunk(chunk ? chunk : 1, 0, b2, o2, &eof2)) return 1;
      const uint64_t n = std::min(o1.size(), o2.size()) - 1;
      if (n == 0) break;
      res.resize(n);
      if (swb_score_batch(c, b1.data(), o1.data(), b2.data(), o2.data(), n, res.data())) return fail(swb_last_error());
      for (uint64_t k = 0; k < n; ++k) total += res[k].score;
      reads += n;
    }
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::memset(out, 0, sizeof *out);
  out->score64 = total; out->score = (int32_t)total; out->processing_time_ms = std::floor(secs * 1000.0);
  std::snprintf(out->gpu_device, sizeof out->gpu_device, "%s", dev ? dev->name : "");
  out->total_reads = reads; out->total_bases = bases1 + bases2;
  return 0;
}

// ---- checkpoint / resume (aligner.rs:23-104) ----
static std::string json_escape(const std::string& s)
{
  std::string o;
  for (unsigned char ch : s) {
    if (ch == '"' || ch == '\\') { o += '\\'; o += (char)ch; }
    else if (ch == '\n') o += "\\n";
    else if (ch == '\t') o += "\\t";
    else if (ch < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
    else o += (char)ch;
  }
  return o;
}

int rsm_checkpoint_save(const char* path, const char* run_id, const rsm_file_checkpoint* files, int n_files, uint64_t total_files)
{
  // serde_json::to_string_pretty layout (aligner.rs:58-59): two-space indent, fields in declaration order
  std::string j = "{\n  \"run_id\": \"" + json_escape(run_id ? run_id : "") + "\",\n  \"files\": [";
  uint64_t completed = 0;
  for (int k = 0; k < n_files; ++k) {
    const rsm_file_checkpoint& f = files[k];
    char num[64];
    j += k ? ",\n    {\n" : "\n    {\n";
    j += "      \"file_path\": \"" + json_escape(f.file_path) + "\",\n";
    j += "      \"file_index\": " + std::to_string(f.file_index) + ",\n";
    j += "      \"score\": " + std::to_string(f.score) + ",\n";
    std::snprintf(num, sizeof num, "%.1f", f.processing_time_ms);
    j += std::string("      \"processing_time_ms\": ") + num + ",\n";
    j += "      \"total_bases\": " + std::to_string(f.total_bases) + ",\n";
    j += "      \"total_reads\": " + std::to_string(f.total_reads) + ",\n";
    j += std::string("      \"completed\": ") + (f.completed ? "true" : "false") + ",\n";
    j += "      \"score64\": " + std::to_string(f.score64) + "\n    }";
    completed += f.completed ? 1 : 0;
  }
  j += n_files ? "\n  ],\n" : "],\n";
  j += "  \"total_files\": " + std::to_string(total_files) + ",\n  \"completed_files\": " + std::to_string(completed) + "\n}";
  const std::string tmp = std::string(path) + ".tmp";
  FILE* fp = std::fopen(tmp.c_str(), "wb");
  if (!fp) return fail(std::string("Failed to create checkpoint file: ") + std::strerror(errno));      // aligner.rs:66
  const bool ok = std::fwrite(j.data(), 1, j.size(), fp) == j.size();
  if (std::fclose(fp) != 0 || !ok) return fail("Failed to write checkpoint: short write");             // aligner.rs:69
  if (std::rename(tmp.c_str(), path) != 0) return fail(std::string("Failed to write checkpoint: ") + std::strerror(errno));
  return 0;
}

// A reader for exactly the JSON the function above (and serde) writes: objects with string / number / bool members.
namespace {
struct JsonCursor {
  const std::string& s; size_t p = 0;
  void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\r' || s[p] == '\t')) ++p; }
  bool eat(char c) { ws(); if (p < s.size() && s[p] == c) { ++p; return true; } return false; }
  bool str(std::string* out)
  {
    ws();
    if (p >= s.size() || s[p] != '"') return false;
    ++p; out->clear();
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\' && p + 1 < s.size()) {
        const char e = s[p + 1];
        if (e == 'n') *out += '\n'; else if (e == 't') *out += '\t';
        else if (e == 'u' && p + 5 < s.size()) { *out += (char)std::strtol(s.substr(p + 2, 4).c_str(), nullptr, 16); p += 4; }
        else *out += e;
        p += 2;
      } else *out += s[p++];
    }
    if (p >= s.size()) return false;
    ++p;
    return true;
  }
  bool scalar(std::string* out)       // number / true / false / null as text
  {
    ws(); out->clear();
    while (p < s.size() && s[p] != ',' && s[p] != '}' && s[p] != ']' && s[p] != ' ' && s[p] != '\n') *out += s[p++];
    return !out->empty();
  }
};
}  // namespace

int rsm_checkpoint_load(const char* path, char* run_id, size_t run_id_cap, rsm_file_checkpoint* files, int cap, int* n_files, uint64_t* total_files)
{
  if (n_files) *n_files = -1;
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return 0;                                            // aligner.rs:81: no checkpoint file exists
  std::string text; char buf[65536]; size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, got);
  std::fclose(fp);
  JsonCursor c{text};
  auto bad = [&]() { return fail("Failed to parse checkpoint: malformed JSON in " + std::string(path)); };   // aligner.rs:77
  if (!c.eat('{')) return bad();
  int n = 0; uint64_t tot = 0; std::string rid;
  for (bool first = true;; first = false) {
    if (c.eat('}')) break;
    if (!first && !c.eat(',')) return bad();
    std::string key, val;
    if (!c.str(&key) || !c.eat(':')) return bad();
    if (key == "files") {
      if (!c.eat('[')) return bad();
      for (bool f0 = true;; f0 = false) {
        if (c.eat(']')) break;
        if (!f0 && !c.eat(',')) return bad();
        if (!c.eat('{')) return bad();
        rsm_file_checkpoint fc; std::memset(&fc, 0, sizeof fc);
        bool has64 = false;
        for (bool m0 = true;; m0 = false) {
          if (c.eat('}')) break;
          if (!m0 && !c.eat(',')) return bad();
          std::string k2, v2;
          if (!c.str(&k2) || !c.eat(':')) return bad();
          if (k2 == "file_path") { if (!c.str(&v2)) return bad(); std::snprintf(fc.file_path, sizeof fc.file_path, "%s", v2.c_str()); continue; }
          if (!c.scalar(&v2)) return bad();
          if (k2 == "file_index") fc.file_index = std::strtoull(v2.c_str(), nullptr, 10);
          else if (k2 == "score") fc.score = (int32_t)std::strtoll(v2.c_str(), nullptr, 10);
          else if (k2 == "score64") { fc.score64 = std::strtoll(v2.c_str(), nullptr, 10); has64 = true; }
          else if (k2 == "processing_time_ms") fc.processing_time_ms = std::strtod(v2.c_str(), nullptr);
          else if (k2 == "total_bases") fc.total_bases = std::strtoull(v2.c_str(), nullptr, 10);
          else if (k2 == "total_reads") fc.total_reads = std::strtoull(v2.c_str(), nullptr, 10);
          else if (k2 == "completed") fc.completed = v2 == "true";
        }
        if (!has64) fc.score64 = fc.score;
        if (files && n < cap) files[n] = fc;
        ++n;
      }
    } else if (key == "run_id") { if (!c.str(&rid)) return bad(); }
    else { if (!c.scalar(&val)) return bad(); if (key == "total_files") tot = std::strtoull(val.c_str(), nullptr, 10); }
  }
  if (run_id && run_id_cap) std::snprintf(run_id, run_id_cap, "%s", rid.c_str());
  if (n_files) *n_files = files ? std::min(n, cap) : n;         // entries written (files == NULL: entries in the file)
  if (total_files) *total_files = tot;
  return 0;
}

int rsm_wgs_file_list(char* buf, size_t cap, int* n_files)
{
  const auto files = wgs_files();
  std::string all;
  for (const auto& f : files) { all += f; all += '\n'; }
  if (n_files) *n_files = (int)files.size();
  if (buf && cap) { std::snprintf(buf, cap, "%s", all.c_str()); }
  return all.size() + 1 > cap ? 1 : 0;
}

int rsm_process_full_wgs_dataset(const rsm_gpu_device* device, rsm_alignment_result* out, int cap, int* n_out)
{
  const auto files = wgs_files();
  const size_t total = files.size();
  uint64_t chunk = 0, chunk_bases = 0;
  if (rsm_get_chunk_size_reads(&chunk) || rsm_get_chunk_size_bases(&chunk_bases)) return 1;        // aligner.rs:207
  const bool compat = ref_compat_mode();
  std::printf("==========================================\n");
  std::printf("GPU PROCESSING STARTING\n");
  std::printf("==========================================\n");
  std::printf("GPU_CHUNK_SIZE_READS: %llu (from .env)\n", (unsigned long long)chunk);
  std::printf("Score mode: %s\n", compat ? "ref_compat (reference's live kernel)" : "sw (Smith-Waterman, reads vs reference windows)");
  std::printf("Processing %zu files (your complete genome)...\n", total);
  std::printf("==========================================\n");

  // devices: the one handed in first, then every other visible GPU (files shard over GPUs)
  rsm_gpu_device devs[64];
  int nd = rsm_get_gpu_devices(devs, 64);
  if (nd <= 0) return fail("No GPU devices found");
  uint64_t want = (uint64_t)nd; std::string why;
  if (const char* v = std::getenv("SWB_NUM_DEVICES")) if (!parse_usize(v, &want, &why) && want >= 1 && want < (uint64_t)nd) nd = (int)want;
  if (compat) nd = 1;
  std::vector<int> order;
  const int first = device ? device->ordinal : 0;
  order.push_back(first);
  for (int d = 0; d < nd; ++d) if (d != first && (int)order.size() < nd) order.push_back(d);

  const auto t_start = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (std::getenv("SWB_DEBUG") || std::getenv("SWB_STAMPS")) std::fprintf(stderr, "[wgs] %-28s %.3f s\n", what, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
  };
  std::vector<uint8_t> ref; uint32_t window_len = 500;
  if (!compat) {
    if (load_reference(ref)) return 1;
    stamp("reference in host memory");
    uint64_t w = 500; if (!parse_usize(env_or("WGS_WINDOW_LEN", "500"), &w, &why) && w >= 1) window_len = (uint32_t)std::min<uint64_t>(w, ref.size());
    std::printf("Reference: %zu bases, window %u bp, %zu GPU(s)\n", ref.size(), window_len, order.size());
  }
  std::vector<FileOutcome> outcomes(total);
  // ---- checkpoint / resume (aligner.rs:218-259) ----
  CheckpointBook book;
  const char* rid = std::getenv("WGS_RUN_ID");
  book.run_id = (rid && *rid) ? rid : "wgs_" + std::to_string((long long)std::time(nullptr));      // aligner.rs:219
  book.path = "checkpoint_" + book.run_id + ".json";                                                // aligner.rs:74 (and, unlike :55, also where it is saved)
  if (const char* d = std::getenv("WGS_CHECKPOINT_DIR")) if (*d) book.path = std::string(d) + "/" + book.path;   // default: the working directory, like the reference
  book.total_files = total;
  {
    std::vector<rsm_file_checkpoint> prior(total + 16);
    char rbuf[256]; int np = -1; uint64_t ptotal = 0;
    // a run without WGS_RUN_ID is a fresh run by definition (its id is a new time stamp, aligner.rs:219): nothing to find
    if (rid && *rid && rsm_checkpoint_load(book.path.c_str(), rbuf, sizeof rbuf, prior.data(), (int)prior.size(), &np, &ptotal)) return 1;
    if (np >= 0) {
      for (int k = 0; k < np && k < (int)prior.size(); ++k) if (prior[k].file_index < total) book.files.push_back(prior[k]);
      size_t done = 0;
      for (const auto& f : book.files) done += f.completed ? 1 : 0;
      std::printf("Found existing checkpoint: %zu files completed\n", done);                       // aligner.rs:224
    } else {
      std::printf("No existing checkpoint found, starting fresh run \n");                          // aligner.rs:228
    }
  }
  std::printf("Checkpoint file: %s \n", book.path.c_str());                                        // aligner.rs:240
  std::vector<size_t> todo;
  for (size_t i = 0; i < total; ++i) {
    if (const rsm_file_checkpoint* fc = book.completed(i)) {                                        // aligner.rs:248-259
      std::printf("Skipping file %zu/%zu (already completed): %s\n", i + 1, total, base_name(files[i]).c_str());
      outcomes[i].res.score = fc->score; outcomes[i].res.score64 = fc->score64; outcomes[i].res.processing_time_ms = fc->processing_time_ms;
      outcomes[i].res.total_reads = fc->total_reads; outcomes[i].res.total_bases = fc->total_bases;
      std::snprintf(outcomes[i].res.gpu_device, sizeof outcomes[i].res.gpu_device, "%s", devs[first].name);
    } else todo.push_back(i);
  }
  g_book = &book;
  g_gpu_call_us = 0;
  struct BookGuard { ~BookGuard() { g_book = nullptr; } } book_guard;
  std::vector<std::thread> workers;
  // Smith-Waterman mode: SWB_CONSUMERS_PER_GPU scoring threads per device, each with its own context (streams, arenas, a
  // copy of the packed reference).  Default 1: on one B200 a second consumer did not pay for its context and arenas
  // (128 M reads from BGZF: 43.1 M reads/s with one, 39.8 M with two).
  uint64_t per_gpu = 1;
  if (const char* v = std::getenv("SWB_CONSUMERS_PER_GPU")) if (parse_usize(v, &per_gpu, &why) || per_gpu < 1 || per_gpu > 8) per_gpu = 1;
  if (compat) per_gpu = 1;
  const size_t n_workers = std::min<size_t>(order.size() * per_gpu, todo.size());      // nothing left to do: no worker, no context
  // BGZF files are read by several threads each when the box has more cores than files (SWB_READERS_PER_FILE overrides)
  // Default: one reader per file, two when a GPU serves a single file.  Every reader costs a page-locked segment buffer, and
  // page-locking is what a short run spends its time on (8 GPUs x 2 files, 256 M reads: 259 M reads/s with one reader per
  // file, 174-187 with two; profiles/wgs_8gpu_consumer_accounting_r02.txt).
  uint64_t rpf = (!todo.empty() && todo.size() < 2 * n_workers && std::thread::hardware_concurrency() >= 2 * todo.size()) ? 2 : 1;
  if (const char* v = std::getenv("SWB_READERS_PER_FILE")) if (parse_usize(v, &rpf, &why) || rpf < 1 || rpf > 16) rpf = 1;
  const unsigned readers_per_file = (unsigned)rpf;
  std::vector<std::string> werr(n_workers);
  std::vector<swb_ctx*> own_ctx(n_workers, nullptr);
  for (size_t w = 0; w < n_workers; ++w) {
    workers.emplace_back([&, w]() {
      const int ord = order[w % order.size()];
      if (compat) {
        for (size_t t = w; t < todo.size(); t += n_workers) {
          const size_t i = todo[t];
          process_one_file_compat(i, total, files[i], chunk, chunk_bases, &devs[ord], &outcomes[i]);
          if (outcomes[i].rc) return;                             // aligner.rs:336: a failed file aborts the run
        }
        return;
      }
      swb_ctx* ctx = nullptr;
      if (w < order.size()) ctx = context_for(ord);               // the first consumer of a device uses the process-wide context
      else if (swb_create(&ctx, ord, nullptr) == 0) own_ctx[w] = ctx;
      else { g_err = std::string("Failed to get GPU context: ") + swb_last_error(); ctx = nullptr; }
      if (!ctx) { werr[w] = g_err; return; }
      stamp("context created");
      if (swb_set_reference(ctx, ref.data(), ref.size())) { werr[w] = swb_last_error(); return; }
      stamp("context + reference on device");
      std::vector<size_t> mine;
      for (size_t t = w; t < todo.size(); t += n_workers) mine.push_back(todo[t]);
      if (mine.empty()) return;
      // plain .gz files: the cores left over once every file has its parsing thread and every GPU its consumer, dealt evenly
      const size_t hc = std::thread::hardware_concurrency(), taken = todo.size() + n_workers;
      const unsigned spare = hc > taken ? (unsigned)std::min<size_t>((hc - taken) / todo.size(), kMaxInflateThreads) : 0u;
      wgs_device_pipeline(mine, files, total, chunk, chunk_bases, &devs[ord], ctx, ref.size(), window_len, &outcomes, readers_per_file, order.size(), spare);
      stamp("a device's files done");
    });
  }
  stamp("reference + workers started");
  for (auto& t : workers) t.join();
  double mem_used_mb = 0;
  { uint64_t fr = 0, to = 0; if (swb_memory_info(first, &fr, &to) == 0) mem_used_mb = (double)(to - fr) / 1048576.0; }
  stamp("all files done");
  if (std::getenv("SWB_DEBUG") || std::getenv("SWB_STAMPS"))
    std::fprintf(stderr, "[wgs] pinned segment buffers: %.1f MB page-locked in %.3f thread-seconds\n", g_pin_bytes.load() / 1e6, g_pin_us.load() / 1e6);
  for (swb_ctx* x : own_ctx) if (x) swb_destroy(x);
  for (const auto& e : werr) if (!e.empty()) return fail(e);
  int n = 0;
  for (size_t i = 0; i < total; ++i) {
    if (outcomes[i].rc) return fail(outcomes[i].err);
    if (n < cap) out[n] = outcomes[i].res;
    ++n;
  }
  if (n_out) *n_out = n;

  // ---- benchmark report (tools/benchmark.rs:17-42, :165-208), with measured numbers instead of the constants of :159-163 ----
  {
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    uint64_t reads = 0, bases = 0; int64_t score = 0;
    for (size_t i = 0; i < total; ++i) { reads += outcomes[i].res.total_reads; bases += outcomes[i].res.total_bases; score += outcomes[i].res.score64; }
    const time_t now = std::time(nullptr);
    char stamp_s[64]; std::strftime(stamp_s, sizeof stamp_s, "%Y-%m-%dT%H:%M:%SZ", std::gmtime(&now));
    double ram_gb = 0;
    if (FILE* mi = std::fopen("/proc/meminfo", "r")) { unsigned long long kb = 0; if (std::fscanf(mi, "MemTotal: %llu kB", &kb) == 1) ram_gb = kb / 1048576.0; std::fclose(mi); }
    const double busy = n_workers ? std::min(1.0, (double)g_gpu_call_us.load() / 1e6 / secs / (double)n_workers) : 0.0;
    const uint64_t w = compat ? 0 : window_len;
    std::string dir = "benchmark_results";
    if (const char* d = std::getenv("WGS_CHECKPOINT_DIR")) if (*d) dir = std::string(d) + "/benchmark_results";
    std::error_code mk_ec;
    std::filesystem::create_directories(dir, mk_ec);            // no shell: the path comes from the environment
    if (!mk_ec) {
      const std::string path = dir + "/run_" + book.run_id + "_benchmark_results.json";
      if (FILE* fp = std::fopen(path.c_str(), "w")) {
        std::fprintf(fp, "{\n  \"timestamp\": \"%s\",\n  \"run_id\": \"%s\",\n  \"mode\": \"full_wgs\",\n  \"files_processed\": %zu,\n"
                         "  \"total_reads\": %llu,\n  \"total_bases\": %llu,\n  \"total_score\": %d,\n  \"total_time_seconds\": %.6f,\n"
                         "  \"throughput_reads_per_second\": %.3f,\n  \"throughput_bases_per_second\": %.3f,\n  \"chunk_size\": %llu,\n"
                         "  \"gpu_utilization_avg\": %.2f,\n  \"gpu_memory_used_mb\": %.1f,\n  \"cpu_cores_used\": %u,\n  \"parallel_files\": true,\n"
                         "  \"system_info\": {\n    \"gpu_name\": \"%s\",\n    \"gpu_memory_gb\": %.1f,\n    \"cpu_cores\": %u,\n    \"total_ram_gb\": %.1f\n  },\n"
                         "  \"total_score64\": %lld,\n  \"n_gpus\": %zu,\n  \"window_len\": %llu,\n  \"gcups_end_to_end\": %.3f\n}\n",
                     stamp_s, book.run_id.c_str(), total, (unsigned long long)reads, (unsigned long long)bases, (int)score, secs,
                     reads / secs, bases / secs, (unsigned long long)chunk, 100.0 * busy, mem_used_mb, std::thread::hardware_concurrency(),
                     devs[first].name, (double)devs[first].memory_gb, std::thread::hardware_concurrency(), ram_gb,
                     (long long)score, order.size(), (unsigned long long)w, (double)bases * (double)w / secs / 1e9);
        std::fclose(fp);
        std::printf("Benchmark results saved to: %s\n", path.c_str());                            // benchmark.rs:187
      }
    }
  }
  return 0;
}

static void usage()
{
  std::printf("High-performance sequence alignment for genome-scale data\n\nUsage: rustseq_mini [OPTIONS]\n\nOptions:\n"
              "  -1, --seq1 <SEQ1>              first sequence or file path\n"
              "  -2, --seq2 <SEQ2>              second sequence or file path\n"
              "  -f, --files                    treat inputs as file paths instead of direct sequences\n"
              "  -c, --chunk-size <CHUNK_SIZE>  chunk size in MB for large sequences [default: 1]\n"
              "  -g, --gpu                      use GPU acceleration if available\n"
              "  -n, --num-files <NUM_FILES>    number of files to process (for multi-file mode)\n"
              "  -t, --test-wgs                 test mode: read WGS files from USB drive\n"
              "      --full-wgs                 process full WGS dataset from all 16 files\n"
              "  -h, --help                     Print help\n");
}

int rsm_main(int argc, char** argv)
{
  dbg_stamp("main entered");
  std::atexit([] { dbg_stamp("atexit (before static destructors)"); });
  load_dotenv();                                                     // main.rs:50
  // SWB_NUM_DEVICES=n (use the first n GPUs) is known before the driver is: hide the others from it, so that a run on one
  // GPU of an eight-GPU box does not pay for initialising eight
  if (const char* v = std::getenv("SWB_NUM_DEVICES")) {
    const long n = std::atol(v);
    if (n >= 1 && n < 64 && !std::getenv("CUDA_VISIBLE_DEVICES")) {
      std::string list;
      for (long d = 0; d < n; ++d) list += (d ? "," : "") + std::to_string(d);
      setenv("CUDA_VISIBLE_DEVICES", list.c_str(), 0);
    }
  }
  std::string seq1, seq2; bool has1 = false, has2 = false, files = false, gpu = false, test_wgs = false, full_wgs = false;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto value = [&](const char* name, std::string* dst) -> int {
      const size_t eq = a.find('=');
      if (a.rfind("--", 0) == 0 && eq != std::string::npos) { *dst = a.substr(eq + 1); return 0; }
      if (i + 1 >= argc) { std::fprintf(stderr, "error: a value is required for '%s' but none was supplied\n", name); return 2; }
      *dst = argv[++i]; return 0;
    };
    std::string tmp;
    if (a == "-1" || a == "--seq1" || a.rfind("--seq1=", 0) == 0) { if (value("--seq1 <SEQ1>", &seq1)) return 2; has1 = true; }
    else if (a == "-2" || a == "--seq2" || a.rfind("--seq2=", 0) == 0) { if (value("--seq2 <SEQ2>", &seq2)) return 2; has2 = true; }
    else if (a == "-f" || a == "--files") files = true;
    else if (a == "-g" || a == "--gpu") gpu = true;
    else if (a == "-t" || a == "--test-wgs") test_wgs = true;
    else if (a == "--full-wgs") full_wgs = true;
    else if (a == "-c" || a == "--chunk-size" || a.rfind("--chunk-size=", 0) == 0) { if (value("--chunk-size <CHUNK_SIZE>", &tmp)) return 2; }   // parsed, unused (main.rs:27-29)
    else if (a == "-n" || a == "--num-files" || a.rfind("--num-files=", 0) == 0) { if (value("--num-files <NUM_FILES>", &tmp)) return 2; }      // parsed, unused (main.rs:35-37)
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else { std::fprintf(stderr, "error: unexpected argument '%s' found\n\nUsage: rustseq_mini [OPTIONS]\n", a.c_str()); return 2; }
  }

  if (full_wgs) {                                                    // main.rs:72-125
    std::printf("Processing FULL WGS dataset from all 16 files...\n");
    if (!gpu || !rsm_is_gpu_available()) {
      std::fprintf(stderr, "error: gpu acceleration is required for full WGS processing\n");   // main.rs:77
      return 1;
    }
    dbg_stamp("gpu available");
    rsm_gpu_device devs[64];
    const int nd = rsm_get_gpu_devices(devs, 64);
    dbg_stamp("devices listed");
    std::printf("GPU acceleration enabled\n");
    for (int d = 0; d < nd && d < 64; ++d) std::printf("  Found GPU: %s (%g GB)\n", devs[d].name, devs[d].memory_gb);
    std::vector<rsm_alignment_result> res(4096);
    int n = 0;
    const int wgs_rc = rsm_process_full_wgs_dataset(&devs[0], res.data(), (int)res.size(), &n);
    dbg_stamp("full wgs returned");
    if (wgs_rc) {
      std::fprintf(stderr, "Full WGS processing error: %s\n", rsm_last_error());               // main.rs:115
      return 1;
    }
    std::printf("\nFULL WGS PROCESSING COMPLETE!\n==========================================\n");
    std::printf("Total files processed: %d\n", n);
    double ms = 0; uint64_t reads = 0, bases = 0;
    for (int i = 0; i < n; ++i) { ms += res[i].processing_time_ms; reads += res[i].total_reads; bases += res[i].total_bases; }
    std::printf("Total reads processed: %llu\nTotal base pairs: %llu\n", (unsigned long long)reads, (unsigned long long)bases);
    std::printf("Total processing time: %.2f seconds\n", ms / 1000.0);
    for (int i = 0; i < n; ++i) std::printf("File %d: Score=%lld, Time=%.2fs\n", i + 1, (long long)res[i].score64, res[i].processing_time_ms / 1000.0);
    dbg_stamp("main returns");
    return 0;
  }

  if (test_wgs) {                                                    // main.rs:127-153: counts bases, needs no GPU
    std::printf("Testing WGS file reading from configured directory...\n");
    const std::string dir = env_or("WGS_DATA_DIR", "/path/to/wgs/data"), sample = env_or("WGS_SAMPLE_ID", "SAMPLE_ID");
    for (const char* suffix : {"_L001_R1_001.fastq.gz", "_L001_R2_001.fastq.gz"}) {
      const std::string file = sample + suffix, full = dir + "/" + file;
      std::printf("Testing: %s\n", full.c_str());
      uint64_t bases = 0;
      if (rsm_count_bases_in_fastq(full.c_str(), &bases) == 0) std::printf("Successfully counted %llu bases in %s\n", (unsigned long long)bases, file.c_str());
      else std::printf("Error counting bases in %s: %s\n", file.c_str(), rsm_last_error());
    }
    return 0;
  }

  if (!has1) { std::fprintf(stderr, "--seq1 is required when not in test mode\n"); return 101; }   // .expect() panic, main.rs:156
  if (!has2) { std::fprintf(stderr, "--seq2 is required when not in test mode\n"); return 101; }   // main.rs:157
  if (!gpu || !rsm_is_gpu_available()) {
    std::fprintf(stderr, "error: gpu acceleration is required and no compatible gpu was found\n");   // main.rs:161
    return 1;
  }
  std::printf("GPU acceleration enabled\n");
  rsm_gpu_device devs[64];
  const int nd = rsm_get_gpu_devices(devs, 64);
  for (int d = 0; d < nd && d < 64; ++d) std::printf("  Found GPU: %s (%g GB)\n", devs[d].name, devs[d].memory_gb);
  if (files) {                                                       // main.rs:170-182
    rsm_alignment_result r;
    if (rsm_gpu_align_pair(seq1.c_str(), seq2.c_str(), &devs[0], &r)) { std::fprintf(stderr, "GPU alignment error: %s\n", rsm_last_error()); return 1; }
    std::printf("GPU Alignment Result:\n  Score: %lld\n  Processing time: %.2f ms\n  GPU device: %s\n", (long long)r.score64, r.processing_time_ms, r.gpu_device);
  } else {                                                           // main.rs:183-190
    swb_result r;
    if (rsm_gpu_align_ex((const uint8_t*)seq1.data(), seq1.size(), (const uint8_t*)seq2.data(), seq2.size(), &devs[0], &r)) {
      std::fprintf(stderr, "GPU alignment error: %s\n", rsm_last_error());
      return 1;
    }
    std::printf("GPU Alignment score: %d\n", r.score);
    if (!ref_compat_mode()) std::printf("  End cell: (%d, %d)\n", r.end_i, r.end_j);
  }
  return 0;
}

}  // extern "C"
