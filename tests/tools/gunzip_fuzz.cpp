// Mutation fuzz of csrc/host_gunzip.h against gzread (truncations, bit flips, overwritten and deleted runs): same error flag, same
// bytes.  g++ -O1 -g -fsanitize=address,undefined -std=c++17 -o /tmp/gunzip_fuzz tests/tools/gunzip_fuzz.cpp -lz;
// /tmp/gunzip_fuzz valid.gz /tmp/scratch.gz <trials> <seed>
#include "../../mini_parallel_b200/csrc/host_gunzip.h"
#include <cstdio>
#include <cstdlib>
#include <random>
int main(int argc, char** argv)
{
  // argv[1]: a valid .gz file; writes mutated copies to argv[2] and reads them with GunzipStream and gzread
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> blob(n); fread(blob.data(), 1, n, f); fclose(f);
  const int trials = atoi(argv[3]);
  std::mt19937_64 rng(atoi(argv[4]));
  std::vector<uint8_t> a(8 << 20), b(8 << 20);
  long n_err = 0, n_same = 0;
  for (int t = 0; t < trials; ++t) {
    std::vector<uint8_t> m = blob;
    const int kind = rng() % 4;
    if (kind == 0) m.resize(rng() % (n + 1));                                  // truncate
    else if (kind == 1) for (int k = 0; k < 1 + (int)(rng() % 3); ++k) m[rng() % n] ^= (uint8_t)(1u << (rng() % 8));
    else if (kind == 2) { size_t p = rng() % n, l = std::min<size_t>(n - p, 1 + rng() % 64); for (size_t k = 0; k < l; ++k) m[p + k] = (uint8_t)rng(); }
    else { size_t p = rng() % n; m.erase(m.begin() + p, m.begin() + std::min<size_t>(n, p + 1 + rng() % 9)); }
    f = fopen(argv[2], "wb"); fwrite(m.data(), 1, m.size(), f); fclose(f);
    size_t na = 0, nb = 0; bool fa = false, fb = false;
    { hgz::GunzipStream g; g.open(argv[2]); const size_t cap = 1 + rng() % (1 << 18);
      for (;;) { if (na + cap > a.size()) a.resize(2 * a.size() + cap); long got = g.read(a.data() + na, cap); if (got < 0) { fa = true; break; } if (got == 0) break; na += got; } }
    { gzFile g = gzopen(argv[2], "rb"); gzbuffer(g, 1 << 16); const unsigned cap = 1 + rng() % (1 << 18);
      for (;;) { if (nb + cap > b.size()) b.resize(2 * b.size() + cap); int got = gzread(g, b.data() + nb, cap); if (got < 0) { fb = true; break; } if (got == 0) break; nb += got; } gzclose(g); }
    if (fa != fb) { printf("trial %d kind %d: failed flags differ: ours %d zlib %d (na %zu nb %zu)\n", t, kind, fa, fb, na, nb); return 1; }
    if (!fa) { if (na != nb || memcmp(a.data(), b.data(), na)) { printf("trial %d kind %d: DATA differs without error na %zu nb %zu\n", t, kind, na, nb); return 1; } ++n_same; }
    else { ++n_err; size_t c = std::min(na, nb); c = c > 70000 ? c - 70000 : 0; if (memcmp(a.data(), b.data(), c)) { printf("trial %d: prefix differs\n", t); return 1; } }
  }
  printf("ok: %d trials, %ld errors agreed, %ld clean agreed\n", trials, n_err, n_same);
  return 0;
}
