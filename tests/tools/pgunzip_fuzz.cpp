// Mutation fuzz of csrc/host_pgunzip.h (hgz::ParallelGunzip) against the serial reader it must be indistinguishable from
// (hgz::GunzipStream): truncations, bit flips, overwritten and deleted runs, random chunk sizes / thread counts / read sizes:
// the same error flag and the same bytes, to the byte, also in front of an error.
//   g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread -o /tmp/pgunzip_fuzz tests/tools/pgunzip_fuzz.cpp -lz
//   /tmp/pgunzip_fuzz valid.gz /tmp/scratch.gz <trials> <seed>          (-fsanitize=thread for the race check)
#include "../../mini_parallel_b200/csrc/host_pgunzip.h"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv)
{
  if (argc < 5) { fprintf(stderr, "usage: pgunzip_fuzz valid.gz scratch.gz trials seed\n"); return 2; }
  FILE* f = fopen(argv[1], "rb"); if (!f) { perror(argv[1]); return 2; }
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> blob(n); if (fread(blob.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
  const int trials = atoi(argv[3]);
  std::mt19937_64 rng(atoi(argv[4]));
  std::vector<uint8_t> a(8 << 20), b(8 << 20);
  long n_err = 0, n_same = 0, n_par = 0; uint64_t acc = 0, ser = 0;
  for (int t = 0; t < trials; ++t) {
    std::vector<uint8_t> m = blob;
    const int kind = t == 0 ? 4 : rng() % 5;
    if (kind == 0) m.resize(rng() % (n + 1));                                  // truncate
    else if (kind == 1) for (int k = 0; k < 1 + (int)(rng() % 3); ++k) m[rng() % n] ^= (uint8_t)(1u << (rng() % 8));
    else if (kind == 2) { size_t p = rng() % n, l = std::min<size_t>(n - p, 1 + rng() % 64); for (size_t k = 0; k < l; ++k) m[p + k] = (uint8_t)rng(); }
    else if (kind == 3) { size_t p = rng() % n; m.erase(m.begin() + p, m.begin() + std::min<size_t>(n, p + 1 + rng() % 9)); }
    // kind 4: the valid file as it is (other chunk sizes and thread counts)
    f = fopen(argv[2], "wb"); fwrite(m.data(), 1, m.size(), f); fclose(f);
    size_t na = 0, nb = 0; bool fa = false, fb = false;
    const unsigned threads = 2 + rng() % 5; const uint64_t chunk = 64 + rng() % (rng() % 2 ? 2000 : 200000);
    {
      hgz::ParallelGunzip g; g.set_threads(threads); g.set_chunk_bytes(chunk);
      if (!g.open(argv[2])) { perror("open"); return 2; }
      const size_t cap = 1 + rng() % (1 << 18);
      for (;;) { if (na + cap > a.size()) a.resize(2 * a.size() + cap); long got = g.read(a.data() + na, cap); if (got < 0) { fa = true; break; } if (got == 0) break; na += got; }
      n_par += g.parallel(); acc += g.chunks_accepted(); ser += g.serial_stretches();
      if (rng() % 8 == 0) { g.close(); }                                        // (else the destructor)
    }
    {
      hgz::GunzipStream g; g.open(argv[2]); const size_t cap = 1 + rng() % (1 << 18);
      for (;;) { if (nb + cap > b.size()) b.resize(2 * b.size() + cap); long got = g.read(b.data() + nb, cap); if (got < 0) { fb = true; break; } if (got == 0) break; nb += got; }
    }
    if (fa != fb) { printf("trial %d kind %d threads %u chunk %llu: failed flags differ: parallel %d serial %d (na %zu nb %zu)\n", t, kind, threads, (unsigned long long)chunk, fa, fb, na, nb); return 1; }
    if (na != nb || memcmp(a.data(), b.data(), na)) { printf("trial %d kind %d threads %u chunk %llu: DATA differs (failed %d) na %zu nb %zu\n", t, kind, threads, (unsigned long long)chunk, fa, na, nb); return 1; }
    if (fa) ++n_err; else ++n_same;
    if (rng() % 16 == 0) {                                                      // a reader abandoned half way: its threads are stopped and joined
      hgz::ParallelGunzip g; g.set_threads(threads); g.set_chunk_bytes(chunk);
      if (g.open(argv[2])) { (void)g.read(a.data(), 1 + rng() % 100000); }
    }
  }
  printf("ok: %d trials, %ld errors agreed, %ld clean agreed; %ld ran on the parallel reader: %llu chunks accepted, %llu serial stretches\n", trials, n_err, n_same, n_par,
         (unsigned long long)acc, (unsigned long long)ser);
  return 0;
}
