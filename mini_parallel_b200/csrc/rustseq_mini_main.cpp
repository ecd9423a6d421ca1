// rustseq_mini -- the reference's CLI (main.rs:11-192) on top of librustseq/libswb200.
#include "../../include/rustseq_host.h"
int main(int argc, char** argv) { return rsm_main(argc, argv); }
