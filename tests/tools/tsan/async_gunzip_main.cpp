#include "../../../mini_parallel_b200/csrc/host_gunzip.h"
#include <chrono>
#include <cstdio>
int main(int argc, char** argv)
{
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
    std::vector<uint8_t> buf(4 << 20);
    auto t0 = std::chrono::steady_clock::now();
    size_t total = 0, lines = 0;
    hgz::GunzipStream g; hgz::AsyncGunzip a;
    if (mode == 0) g.open(argv[1]); else a.open(argv[1]);
    for (;;) {
      long got = mode == 0 ? g.read(buf.data(), buf.size()) : a.read(buf.data(), buf.size());
      if (got <= 0) break;
      total += got;
      for (int pass = 0; pass < 1; ++pass) { const uint8_t* p = buf.data(); const uint8_t* e = p + got; while ((p = (const uint8_t*)memchr(p, '\n', e - p))) { ++lines; ++p; } }   // ~parse cost
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%s: %zu bytes %zu lines %.3f s\n", mode ? "async" : "sync ", total, lines, dt);
  }
}
