// inflate_bench.cu -- the BGZF inflate kernel and the FASTQ index kernels alone, on synthetic FASTQ text.
// Compresses N distinct 65 280-byte blocks with zlib (raw deflate, level 1 by default), replicates them to a segment of
// the size the WGS driver uses, and times launch_inflate_bgzf / launch_fq_index / launch_fq_extract with CUDA events.
// Two quality models: constant 'I' (what tools/bench_wgs.py writes) and noisy (a quality string a sequencer would write:
// mostly literals after deflate).   build: make build/inflate_bench     run: build/inflate_bench [blocks] [level]
#include "../mini_parallel_b200/csrc/swb_kernels.cuh"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>

static uint64_t splitmix(uint64_t& s) { uint64_t x = (s += 0x9E3779B97F4A7C15ull); x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }

static std::string make_text(size_t bytes, bool noisy_quals, uint64_t seed)
{
  std::string t; t.reserve(bytes + 400);
  uint64_t s = seed, k = 0;
  while (t.size() < bytes) {
    char hdr[32]; std::snprintf(hdr, sizeof hdr, "@%010llu\n", (unsigned long long)k++);
    t += hdr;
    for (int i = 0; i < 150; i += 32) { uint64_t x = splitmix(s); for (int j = 0; j < 32 && i + j < 150; ++j) t += "ACGT"[(x >> (2 * j)) & 3]; }
    t += "\n+\n";
    if (!noisy_quals) t.append(150, 'I');
    else for (int i = 0; i < 150; i += 8) { uint64_t x = splitmix(s); for (int j = 0; j < 8 && i + j < 150; ++j) { const unsigned r = (x >> (8 * j)) & 255; t += r < 160 ? 'F' : r < 224 ? ':' : r < 248 ? ',' : '#'; } }
    t += '\n';
  }
  t.resize(bytes);
  return t;
}

int main(int argc, char** argv)
{
  const size_t n_blocks = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 9472;
  const int level = argc > 2 ? std::atoi(argv[2]) : 1;
  const size_t kBlock = 65280, kDistinct = 592;
  for (int noisy = 0; noisy < 2; ++noisy) {
    const std::string text = make_text(kBlock * kDistinct, noisy != 0, 0xB200 + noisy);
    std::vector<uint8_t> comp; std::vector<swb_bgzf_block> proto(kDistinct);
    for (size_t b = 0; b < kDistinct; ++b) {
      z_stream z; std::memset(&z, 0, sizeof z);
      deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
      std::vector<uint8_t> out(deflateBound(&z, kBlock));
      z.next_in = (Bytef*)text.data() + b * kBlock; z.avail_in = kBlock; z.next_out = out.data(); z.avail_out = out.size();
      deflate(&z, Z_FINISH);
      proto[b] = swb_bgzf_block{comp.size(), (uint32_t)z.total_out, (uint32_t)kBlock};
      comp.insert(comp.end(), out.begin(), out.begin() + z.total_out);
      deflateEnd(&z);
    }
    std::vector<swb_bgzf_block> blocks(n_blocks); std::vector<uint64_t> out_off(n_blocks);
    for (size_t b = 0; b < n_blocks; ++b) { blocks[b] = proto[b % kDistinct]; out_off[b] = b * kBlock; }
    const size_t text_bytes = n_blocks * kBlock;
    uint8_t *d_comp, *d_text; swb_bgzf_block* d_blocks; uint64_t* d_off; uint32_t* d_fail;
    cudaMalloc(&d_comp, comp.size() + 64); cudaMalloc(&d_text, text_bytes + 4096); cudaMalloc(&d_blocks, n_blocks * sizeof(swb_bgzf_block));
    cudaMalloc(&d_off, n_blocks * 8); cudaMalloc(&d_fail, 64);
    cudaMemcpy(d_comp, comp.data(), comp.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_blocks, blocks.data(), n_blocks * sizeof(swb_bgzf_block), cudaMemcpyHostToDevice);
    cudaMemcpy(d_off, out_off.data(), n_blocks * 8, cudaMemcpyHostToDevice);
    cudaMemset(d_fail, 0, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 4; ++it) {
      if (it == 1) cudaEventRecord(e0);
      swb::launch_inflate_bgzf(d_comp, d_blocks, n_blocks, d_off, d_text, d_fail, 0);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    uint32_t h_fail[4]; cudaMemcpy(h_fail, d_fail, 16, cudaMemcpyDeviceToHost);
    std::vector<uint8_t> back(kBlock * 3);
    cudaMemcpy(back.data(), d_text + (n_blocks - 3) * kBlock, back.size(), cudaMemcpyDeviceToHost);
    bool same = true;
    for (size_t b = n_blocks - 3; b < n_blocks; ++b) same = same && !std::memcmp(back.data() + (b - (n_blocks - 3)) * kBlock, text.data() + (b % kDistinct) * kBlock, kBlock);
    // index kernels on the inflated text
    const uint64_t begin = 0, end = text_bytes, tiles = swb::fq_tiles(begin, end);
    uint32_t* d_tc; uint64_t *d_tp, *d_scal, *d_beg, *d_end;
    const uint64_t recs = text_bytes / 300 + 16;
    cudaMalloc(&d_tc, tiles * 4 + 64); cudaMalloc(&d_tp, tiles * 8 + 64); cudaMalloc(&d_scal, 64); cudaMalloc(&d_beg, recs * 8); cudaMalloc(&d_end, recs * 8);
    cudaMemset(d_scal, 0, 64);
    float ms_index = 0, ms_em = 0;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      swb::launch_fq_index(d_text, begin, end, d_tc, d_tp, d_scal, reinterpret_cast<uint32_t*>(d_scal + 4) + 1, 1, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_index, e0, e1);
    }
    uint32_t *d_pk, *d_bm;
    cudaMalloc(&d_pk, (tiles + 2) * 1024 + 64); cudaMalloc(&d_bm, (tiles + 2) * 32 + 64);
    for (int it = 0; it < 2; ++it) {                           // (the second pass sees masked text: same work)
      cudaEventRecord(e0);
      swb::launch_fq_extract(d_text, begin, end, d_tc, d_tp, d_scal, d_beg, d_end, recs, reinterpret_cast<unsigned long long*>(d_scal + 1), 1, d_pk, d_bm, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_em, e0, e1);
    }
    cudaFree(d_pk); cudaFree(d_bm);
    std::printf("{\"quals\": \"%s\", \"level\": %d, \"blocks\": %zu, \"text_mb\": %.1f, \"comp_ratio\": %.2f, \"inflate_ms\": %.3f, \"inflate_gb_s\": %.1f, "
                "\"failed\": %u, \"same\": %s, \"index_ms\": %.3f, \"extract_mask_pack_ms\": %.3f, \"cuda\": \"%s\"}\n",
                noisy ? "noisy" : "constant", level, n_blocks, text_bytes / 1e6, (double)(kBlock * kDistinct) / comp.size(), ms, text_bytes / ms / 1e6,
                h_fail[0], same ? "true" : "false", ms_index, ms_em, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d_comp); cudaFree(d_text); cudaFree(d_blocks); cudaFree(d_off); cudaFree(d_fail); cudaFree(d_tc); cudaFree(d_tp); cudaFree(d_scal); cudaFree(d_beg); cudaFree(d_end);
  }
  return 0;
}
