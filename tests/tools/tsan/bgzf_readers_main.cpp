#include "../../../include/rustseq_host.h"
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
int main(int argc, char** argv)
{
  for (int rep = 0; rep < 3; ++rep)
    for (unsigned readers : {1u, 2u, 3u, 8u}) {
      uint64_t a, b, c, h; int st;
      int rc = rsm_debug_bgzf_segments(argv[1], readers, 128u << 10, 2 + readers % 3, &a, &b, &c, &h, &st);
      printf("readers %u rc %d status %d segments %lu blocks %lu text %lu hash %016lx\n", readers, rc, st, a, b, c, h);
    }
  return 0;
}
