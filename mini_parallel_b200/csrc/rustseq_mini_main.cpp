// rustseq_mini -- the reference's CLI (main.rs:11-192) on top of librustseq/libswb200.
#include "../../include/rustseq_host.h"
#include <cstdio>
#include <unistd.h>
// The process is done when rsm_main returns: everything it owns (device arenas, CUDA contexts, page-locked buffers) goes
// back with the address space.  Leaving through exit() instead hands the CUDA runtime's own tear-down ~2 s on an eight-GPU
// box (measured: profiles/wgs_scaling_8gpu_r02.txt) for nothing.
int main(int argc, char** argv)
{
  const int rc = rsm_main(argc, argv);
  std::fflush(stdout); std::fflush(stderr);
  _exit(rc);
}
