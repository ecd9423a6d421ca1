// gunzip_bench.cpp -- host-side throughput of csrc/host_gunzip.h on a gzip file, beside zlib's gzread (no GPU):
//   g++ -O2 -std=c++17 -pthread -o build/gunzip_bench tools/gunzip_bench.cpp -lz
//   build/gunzip_bench file.gz [reps] [threads]      threads > 1: hgz::ParallelGunzip with that many decoder threads
// Prints MB/s of inflated text per decoder (best of `reps`) and checks that all of them deliver the same bytes (length + CRC-32).
#include "../mini_parallel_b200/csrc/host_pgunzip.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class R> static bool run(R& r, const char* path, uint64_t* bytes, uint32_t* crc, double* secs)
{
  std::vector<uint8_t> buf(4 << 20);
  const double t0 = now();
  if (!r.open(path)) { std::perror(path); return false; }
  uint64_t n = 0; uint32_t c = 0;
  for (;;) {
    const long got = r.read(buf.data(), buf.size());
    if (got < 0) { std::fprintf(stderr, "corrupt: %s\n", r.error().c_str()); return false; }
    if (got == 0) break;
    c = hgz::crc32_update(c, buf.data(), (size_t)got); n += (uint64_t)got;
  }
  r.close();
  *secs = now() - t0; *bytes = n; *crc = c;
  return true;
}

int main(int argc, char** argv)
{
  if (argc < 2) { std::fprintf(stderr, "usage: gunzip_bench file.gz [reps] [threads]\n"); return 2; }
  const int reps = argc > 2 ? std::atoi(argv[2]) : 3;
  const int threads = argc > 3 ? std::atoi(argv[3]) : 1;
  uint64_t n0 = 0; uint32_t c0 = 0; double best = 1e30;
  {                                                     // zlib
    std::vector<uint8_t> buf(4 << 20);
    for (int k = 0; k < reps; ++k) {
      const double t0 = now();
      gzFile g = gzopen(argv[1], "rb");
      if (!g) { std::perror(argv[1]); return 1; }
      gzbuffer(g, 1 << 20);
      uint64_t n = 0; uint32_t c = 0;
      for (;;) { const int got = gzread(g, buf.data(), (unsigned)buf.size()); if (got <= 0) break; c = hgz::crc32_update(c, buf.data(), (size_t)got); n += (uint64_t)got; }
      gzclose(g);
      best = std::min(best, now() - t0); n0 = n; c0 = c;
    }
    std::printf("zlib gzread        %8.1f MB/s  (%llu bytes, crc %08x)\n", n0 / best / 1e6, (unsigned long long)n0, c0);
  }
  {
    double b = 1e30; uint64_t n = 0; uint32_t c = 0;
    for (int k = 0; k < reps; ++k) { hgz::GunzipStream r; double s; if (!run(r, argv[1], &n, &c, &s)) return 1; b = std::min(b, s); }
    std::printf("hgz::GunzipStream  %8.1f MB/s  %s\n", n / b / 1e6, (n == n0 && c == c0) ? "same bytes" : "DIFFERENT BYTES");
    if (n != n0 || c != c0) return 1;
  }
#ifdef HGZ_HAVE_PARALLEL
  if (threads > 1) {
    double b = 1e30; uint64_t n = 0; uint32_t c = 0;
    unsigned long long acc = 0, ser = 0;
    for (int k = 0; k < reps; ++k) {
      hgz::ParallelGunzip r; r.set_threads((unsigned)threads); double s;
      if (!run(r, argv[1], &n, &c, &s)) return 1;
      b = std::min(b, s); acc = r.chunks_accepted(); ser = r.serial_stretches();
    }
    std::printf("hgz::ParallelGunzip %7.1f MB/s  %d threads  %s  (%llu chunks from the decoder threads, %llu serial stretches)\n", n / b / 1e6, threads,
                (n == n0 && c == c0) ? "same bytes" : "DIFFERENT BYTES", acc, ser);
    if (n != n0 || c != c0) return 1;
  }
#else
  (void)threads;
#endif
  return 0;
}
