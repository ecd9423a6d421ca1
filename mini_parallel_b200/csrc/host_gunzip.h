// host_gunzip.h -- streaming gzip reader for the HOST FASTQ path (plain single- or multi-member .gz files, the ones the GPU
// BGZF path cannot take): what gzopen/gzread did for FastqReader, with a decoder built for throughput.
//
// The reference inflates through a `zcat` child (aligner.rs:109-120); round 1 replaced it by in-process zlib, which spends
// 88 % of a reader thread in inflate() at ~285 MB/s of FASTQ text per core.  This decoder is the usual modern design -- a
// 64-bit bit buffer refilled with one unaligned load, an 11-bit literal/length table and an 8-bit distance table whose
// entries carry base value, extra-bit count and code length, sub-tables for longer codes, word-wise match copies with a
// broadcast path for distance 1 (quality strings), the CRC-32 folded with PCLMULQDQ -- and keeps gzread's behaviour at the edges (listed at GunzipStream).
// Host only; the device decoder is swb_inflate.cuh.  Checked against zlib on every level / strategy, random read sizes,
// multi-member files, corrupt and truncated input (tests/test_host_gunzip.py).
#pragma once
#include <stdint.h>
#include <string.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>          // crc32() only
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace hgz {

// ---- CRC-32 of the inflated text ----
// zlib's crc32 runs at ~2.3 GB/s, a fifth of this reader's time.  On x86-64 with PCLMULQDQ the checksum is folded 64 bytes
// at a time with carry-less multiplications (the fold-by-4 scheme of Intel's "Fast CRC Computation for Generic Polynomials
// Using PCLMULQDQ", constants for the reflected IEEE polynomial); heads, tails and other CPUs use zlib.  The first use checks
// the folded result against zlib's on a test pattern and falls back for good if they differ.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1")))
static inline uint32_t crc32_fold(const uint8_t* buf, size_t len /* >= 64, multiple of 16 */, uint32_t crc /* inverted */)
{
  alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};
  alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};
  alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0};
  alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0));  x2 = _mm_loadu_si128((const __m128i*)(buf + 16));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 32)); x4 = _mm_loadu_si128((const __m128i*)(buf + 48));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i*)k1k2);
  buf += 64; len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), _mm_loadu_si128((const __m128i*)(buf + 0)));
    x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), _mm_loadu_si128((const __m128i*)(buf + 16)));
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), _mm_loadu_si128((const __m128i*)(buf + 32)));
    x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), _mm_loadu_si128((const __m128i*)(buf + 48)));
    buf += 64; len -= 64;
  }
  x0 = _mm_load_si128((const __m128i*)k3k4);                        // four lanes into one
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16; len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);                          // 128 -> 64 bits
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_xor_si128(_mm_srli_si128(x1, 8), x2);
  x0 = _mm_loadl_epi64((const __m128i*)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(x1, x3), x0, 0x00), x2);
  x0 = _mm_load_si128((const __m128i*)poly);                        // Barrett reduction to 32 bits
  x2 = _mm_clmulepi64_si128(_mm_and_si128(x1, x3), x0, 0x10);
  x2 = _mm_clmulepi64_si128(_mm_and_si128(x2, x3), x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}
#endif

static inline bool crc32_fold_usable()
{
#if defined(__x86_64__)
  static const bool ok = [] {
    if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
    uint8_t t[1024 + 48];
    for (size_t i = 0; i < sizeof t; ++i) t[i] = (uint8_t)(i * 131u + (i >> 3) * 7u + 5u);
    for (size_t len : {(size_t)64, (size_t)80, (size_t)256, (size_t)1024, (size_t)1072}) {
      const uint32_t seed = 0x1234567u * (uint32_t)len;
      if (~crc32_fold(t, len, ~seed) != (uint32_t)crc32_z(seed, t, len)) return false;
    }
    return true;
  }();
  return ok;
#else
  return false;
#endif
}

static inline uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n)
{
#if defined(__x86_64__)
  if (n >= 256 && crc32_fold_usable()) {
    const size_t body = n & ~(size_t)15;
    crc = ~crc32_fold(p, body, ~crc);
    p += body; n -= body;
  }
#endif
  return n ? (uint32_t)crc32_z(crc, p, n) : crc;
}

constexpr int LIT_TB = 11, DIST_TB = 8;
constexpr uint32_t K_LIT = 0x8000u, K_EXC = 0x4000u, K_SUB = 0x2000u;     // entry: value<<16 | kind | extra<<8 | bits, where `bits` is what the
                                                                          // symbol consumes: its code AND its extra bits (one shift per symbol)
constexpr uint32_t V_EOB = 0, V_BAD = 1;

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

static const uint16_t kLenBase[29]  = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t  kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                                       8193, 12289, 16385, 24577};
static const uint8_t  kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t  kClOrder[19]  = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Alphabet { A_CODELEN, A_LITLEN, A_DIST };

static inline uint32_t symbol_entry(Alphabet a, int s)
{
  if (a == A_CODELEN) return (uint32_t)s << 16;
  if (a == A_LITLEN) {
    if (s < 256) return ((uint32_t)s << 16) | K_LIT;
    if (s == 256) return (V_EOB << 16) | K_EXC;
    if (s <= 285) return ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)kLenExtra[s - 257] << 8);
    return (V_BAD << 16) | K_EXC;                                    // 286, 287: in the fixed code, never legal in data
  }
  if (s < 30) return ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 8);
  return (V_BAD << 16) | K_EXC;
}

static inline uint32_t bitrev(uint32_t c, int len)
{
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) { r = (r << 1) | (c & 1u); c >>= 1; }
  return r;
}

// Canonical Huffman code -> two-level decode table.  zlib's acceptance rules (inftrees.c): over-subscribed sets are errors;
// an incomplete set is accepted only when its longest code has one bit (a single distance code, or none at all); unused
// table slots decode to V_BAD.  Returns false for a set zlib rejects.  `table` has room for (1 << tb) + n * (1 << (15 - tb)).
static inline bool build_table(const uint8_t* lens, int n, Alphabet a, int tb, uint32_t* table)
{
  int count[16] = {0};
  for (int s = 0; s < n; ++s) ++count[lens[s]];
  int max = 15;
  while (max > 0 && count[max] == 0) --max;
  const uint32_t bad = (V_BAD << 16) | K_EXC | 1u;                    // consumes a bit, so even a careless caller moves on
  const int main_size = 1 << tb;
  if (max == 0) { for (int i = 0; i < main_size; ++i) table[i] = bad; return true; }   // no codes: any use is an error (zlib does the same)
  int left = 1;
  for (int len = 1; len <= 15; ++len) { left = (left << 1) - count[len]; if (left < 0) return false; }
  if (left > 0 && (a == A_CODELEN || max != 1)) return false;
  uint32_t next[16]; next[0] = 0;                                     // first canonical code of every length
  { uint32_t code = 0; for (int len = 1; len <= 15; ++len) { next[len] = code; code = (code + (uint32_t)count[len]) << 1; } }
  for (int i = 0; i < main_size; ++i) table[i] = bad;
  // sub-table sizes: the longest code under each main-table prefix
  uint8_t sub_bits[1 << LIT_TB];
  if (max > tb) {
    memset(sub_bits, 0, (size_t)main_size);
    uint32_t nx[16]; memcpy(nx, next, sizeof nx);
    for (int s = 0; s < n; ++s) {
      const int len = lens[s];
      if (len <= tb) { if (len) ++nx[len]; continue; }
      const uint32_t rev = bitrev(nx[len]++, len), pre = rev & (uint32_t)(main_size - 1);
      if (len - tb > sub_bits[pre]) sub_bits[pre] = (uint8_t)(len - tb);
    }
    uint32_t off = (uint32_t)main_size;
    for (int p = 0; p < main_size; ++p)
      if (sub_bits[p]) {
        table[p] = (off << 16) | K_EXC | K_SUB | ((uint32_t)sub_bits[p] << 8) | (uint32_t)tb;
        for (uint32_t k = 0; k < (1u << sub_bits[p]); ++k) table[off + k] = bad;
        off += 1u << sub_bits[p];
      }
  }
  for (int s = 0; s < n; ++s) {
    const int len = lens[s];
    if (!len) continue;
    const uint32_t rev = bitrev(next[len]++, len);
    if (len <= tb) {
      const uint32_t se = symbol_entry(a, s), xb = (se & (K_LIT | K_EXC)) ? 0u : (se >> 8) & 15u;
      const uint32_t e = se | ((uint32_t)len + xb);
      for (uint32_t i = rev; i < (uint32_t)main_size; i += 1u << len) table[i] = e;
    } else {
      const uint32_t pre = rev & (uint32_t)(main_size - 1), head = table[pre];
      const uint32_t off = head >> 16, sb = (head >> 8) & 15u;
      const uint32_t se = symbol_entry(a, s), xb = (se & (K_LIT | K_EXC)) ? 0u : (se >> 8) & 15u;
      const uint32_t e = se | ((uint32_t)(len - tb) + xb);
      for (uint32_t i = rev >> tb; i < (1u << sb); i += 1u << (len - tb)) table[off + i] = e;
    }
  }
  return true;
}

// One deflate stream, resumable between symbols.  The caller owns input and output windows:
//   in / in_end      compressed bytes; at least 8 readable bytes follow in_end (padding), `final_input` says no more will come
//   out / out_end    output cursor and limit inside a buffer whose previous `hist` bytes (<= 32768) are this stream's history
struct Inflater {
  uint64_t bitbuf = 0; int bitcnt = 0;
  int state = 0;                     // 0 block header, 1 stored, 2 huffman, 3 done
  bool last_block = false;
  uint32_t stored_left = 0;
  uint32_t lit[(1 << LIT_TB) + 288 * (1 << (15 - LIT_TB))];
  uint32_t dist[(1 << DIST_TB) + 32 * (1 << (15 - DIST_TB))];
  uint32_t clen[1 << 7];
  const char* error = nullptr;
  bool starved = false;              // the FINAL input ended inside the stream: nothing more can be decoded, ever
  // Optional stop at a block boundary (the parallel reader, host_pgunzip.h): with pos_base set, run() returns BOUNDARY in front
  // of the first block header whose bit position -- counted from pos_base -- is >= stop_bit, and leaves it in boundary_bit.
  const uint8_t* pos_base = nullptr;
  uint64_t stop_bit = ~0ull, boundary_bit = 0;

  void reset() { bitbuf = 0; bitcnt = 0; state = 0; last_block = false; stored_left = 0; error = nullptr; starved = false; pos_base = nullptr; stop_bit = ~0ull; }
  // start in the middle of a byte: the stream's next bit is bit (bit & 7) of base[bit >> 3]; returns the input cursor for run()
  const uint8_t* start_at_bit(const uint8_t* base, uint64_t bit)
  {
    reset();
    const uint8_t* in = base + (bit >> 3);
    const int r = (int)(bit & 7);
    if (r) { bitbuf = (uint64_t)(*in++ >> r); bitcnt = 8 - r; }
    return in;
  }

  enum Result { NEED_INPUT, NEED_OUTPUT, DONE, ERROR, BOUNDARY };

  static inline void refill_fast(uint64_t& bb, int& bc, const uint8_t*& in)
  {
    bb |= load64(in) << bc;
    in += (63 - bc) >> 3;
    bc |= 56;
  }
  static inline void refill_safe(uint64_t& bb, int& bc, const uint8_t*& in, const uint8_t* in_end)
  {
    while (bc < 56 && in < in_end) { bb |= (uint64_t)*in++ << bc; bc += 8; }       // never to 64: refill_fast shifts by bc
  }

  bool fixed_tables()
  {
    uint8_t l[288 + 32];
    for (int s = 0; s < 288; ++s) l[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    for (int s = 0; s < 32; ++s) l[288 + s] = 5;
    return build_table(l, 288, A_LITLEN, LIT_TB, lit) && build_table(l + 288, 32, A_DIST, DIST_TB, dist);
  }

  // The dynamic block header, parsed in one go: the caller makes sure all of it (at most 563 bytes) is in [in, in_end) unless
  // the input is final, so running dry in here means a truncated stream.
  Result dynamic_tables(uint64_t& bb, int& bc, const uint8_t*& in, const uint8_t* in_end)
  {
    auto take = [&](int k) -> uint32_t { const uint32_t v = (uint32_t)(bb & ((1ull << k) - 1)); bb >>= k; bc -= k; return v; };
    auto starve = [&]() -> Result { error = "unexpected end of file"; starved = true; return NEED_INPUT; };
    refill_safe(bb, bc, in, in_end);
    if (bc < 14) return starve();
    const int nlit = (int)take(5) + 257, ndist = (int)take(5) + 1, ncl = (int)take(4) + 4;
    if (nlit > 286 || ndist > 30) { error = "too many length or distance symbols"; return ERROR; }
    uint8_t cl[19] = {0};
    for (int i = 0; i < ncl; ++i) {
      if (bc < 3) { refill_safe(bb, bc, in, in_end); if (bc < 3) return starve(); }
      cl[kClOrder[i]] = (uint8_t)take(3);
    }
    if (!build_table(cl, 19, A_CODELEN, 7, clen)) { error = "invalid code lengths set"; return ERROR; }
    uint8_t lens[320 + 138];
    int i = 0;
    while (i < nlit + ndist) {
      if (bc < 14) refill_safe(bb, bc, in, in_end);                  // a code (<= 7 bits) and its extra bits (<= 7)
      const uint32_t e = clen[bb & 127u];
      const int bits = (int)(e & 255u);
      if (bits > bc) return starve();
      if (e & K_EXC) { error = "invalid code lengths set"; return ERROR; }
      bb >>= bits; bc -= bits;
      const int s = (int)(e >> 16);
      if (s < 16) { lens[i++] = (uint8_t)s; continue; }
      const int xb = s == 16 ? 2 : s == 17 ? 3 : 7;
      if (bc < xb) return starve();
      int rep; uint8_t v = 0;
      if (s == 16) { if (i == 0) { error = "invalid bit length repeat"; return ERROR; } v = lens[i - 1]; rep = 3 + (int)take(2); }
      else if (s == 17) rep = 3 + (int)take(3);
      else rep = 11 + (int)take(7);
      if (i + rep > nlit + ndist) { error = "invalid bit length repeat"; return ERROR; }
      for (int k = 0; k < rep; ++k) lens[i++] = v;
    }
    if (lens[256] == 0) { error = "invalid code -- missing end-of-block"; return ERROR; }
    if (!build_table(lens, nlit, A_LITLEN, LIT_TB, lit)) { error = "invalid literal/lengths set"; return ERROR; }
    if (!build_table(lens + nlit, ndist, A_DIST, DIST_TB, dist)) { error = "invalid distances set"; return ERROR; }
    return DONE;
  }

  // Decodes until the stream ends, the output is full, or the input runs dry.  Advances in / out.
  Result run(const uint8_t*& in_ref, const uint8_t* in_end, bool final_input, uint8_t*& out_ref, uint8_t* out_end, uint64_t hist)
  {
    if (starved) return NEED_INPUT;
    const uint8_t* in = in_ref; uint8_t* out = out_ref;
    uint8_t* const out_start = out;
    uint64_t bb = bitbuf; int bc = bitcnt;
    Result res = DONE;
    auto save = [&]() { bitbuf = bb; bitcnt = bc; in_ref = in; out_ref = out; };
    for (;;) {
      if (state == 3) { res = DONE; break; }
      if (state == 0) {                                                // ---- block header ----
        if (pos_base) {
          const uint64_t bp = (uint64_t)(in - pos_base) * 8 - (uint64_t)bc;
          if (bp >= stop_bit) { boundary_bit = bp; res = BOUNDARY; break; }
        }
        // a dynamic header is parsed in one go: wait until it is certainly all there
        if (!final_input && (in_end - in) < 1024) { res = NEED_INPUT; break; }
        refill_safe(bb, bc, in, in_end);
        if (bc < 3) { error = "unexpected end of file"; starved = true; res = NEED_INPUT; break; }
        last_block = bb & 1u; const uint32_t type = (uint32_t)(bb >> 1) & 3u;
        bb >>= 3; bc -= 3;
        if (type == 0) {
          const int drop = bc & 7; bb >>= drop; bc -= drop;            // to the byte boundary
          refill_safe(bb, bc, in, in_end);
          if (bc < 32) { error = "unexpected end of file"; starved = true; res = NEED_INPUT; break; }
          const uint32_t len = (uint32_t)(bb & 0xFFFFu), nlen = (uint32_t)((bb >> 16) & 0xFFFFu);
          bb >>= 32; bc -= 32;
          if ((len ^ nlen) != 0xFFFFu) { error = "invalid stored block lengths"; res = ERROR; break; }
          stored_left = len; state = 1;
        } else if (type == 1) {
          fixed_tables(); state = 2;
        } else if (type == 2) {
          const Result r = dynamic_tables(bb, bc, in, in_end);
          if (r != DONE) { res = r; break; }
          state = 2;
        } else { error = "invalid block type"; res = ERROR; break; }
        continue;
      }
      if (state == 1) {                                                // ---- stored ----
        while (stored_left && bc >= 8) {                               // bytes already in the bit buffer
          if (out >= out_end) break;
          *out++ = (uint8_t)bb; bb >>= 8; bc -= 8; --stored_left;
        }
        if (stored_left && bc >= 8) { res = NEED_OUTPUT; break; }
        if (stored_left) {
          bb = 0; bc = 0;                                              // (bits past bc were never counted: `in` is the next byte)
          const uint64_t n = std::min<uint64_t>(std::min<uint64_t>(stored_left, (uint64_t)(in_end - in)), (uint64_t)(out_end - out));
          memcpy(out, in, n); out += n; in += n; stored_left -= (uint32_t)n;
          if (stored_left) {
            if (out >= out_end) { res = NEED_OUTPUT; break; }
            if (final_input) { error = "unexpected end of file"; starved = true; }
            res = NEED_INPUT; break;
          }
        }
        state = last_block ? 3 : 0;
        continue;
      }
      // ---- huffman symbols ----
      // fast loop: >= 16 input bytes and >= 258 + 16 output bytes in hand, no bounds checks inside.  The table entry of the
      // NEXT symbol is fetched before the copy of a match runs (the load's latency hides behind the copy), and literals run
      // three to a refill.  (Shift counts are written `& 63`: an entry's bit count is below 64, and a 64-bit shift takes its
      // count modulo 64 anyway, so the mask costs no instruction where `& 255` put one on the critical path.)
      bool ended = false;
      if (in_end - in >= 16 && out_end - out >= 280) {
        refill_fast(bb, bc, in);
        uint32_t e = lit[bb & ((1u << LIT_TB) - 1)];
        for (;;) {
          if (e & K_LIT) {
            bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint8_t)(e >> 16);
            e = lit[bb & ((1u << LIT_TB) - 1)];
            if (e & K_LIT) {
              bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint8_t)(e >> 16);
              e = lit[bb & ((1u << LIT_TB) - 1)];
              if (e & K_LIT) {
                bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint8_t)(e >> 16);
                e = lit[bb & ((1u << LIT_TB) - 1)];
              }
            }
            if (!(in_end - in >= 16 && out_end - out >= 280)) break;
            refill_fast(bb, bc, in);                                 // (e looked at the low bits only: still the entry of the next symbol)
            continue;
          }
          if (e & K_EXC) {
            if (e & K_SUB) {
              bb >>= LIT_TB; bc -= LIT_TB;
              e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 15u)) - 1))];
              if (e & K_LIT) {
                bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint8_t)(e >> 16);
                if (!(in_end - in >= 16 && out_end - out >= 280)) break;
                refill_fast(bb, bc, in);
                e = lit[bb & ((1u << LIT_TB) - 1)];
                continue;
              }
            }
            if (e & K_EXC) {
              bb >>= (e & 63u); bc -= (int)(e & 255u);
              if ((e >> 16) == V_EOB && !(e & K_SUB)) { ended = true; break; }
              error = "invalid literal/length code"; save(); return ERROR;
            }
          }
          // a length: base + extra bits, then the distance
          const uint64_t lsaved = bb;
          bb >>= (e & 63u); bc -= (int)(e & 255u);                  // code and extra bits in one shift; the extra bits come from lsaved
          uint32_t d = dist[bb & ((1u << DIST_TB) - 1)];
          const uint32_t lxb = (e >> 8) & 15u;
          const uint32_t len = (e >> 16) + ((uint32_t)(lsaved >> ((e & 255u) - lxb)) & ((1u << lxb) - 1));
          if (d & K_EXC) {
            if (!(d & K_SUB)) { error = "invalid distance code"; save(); return ERROR; }
            bb >>= DIST_TB; bc -= DIST_TB;
            d = dist[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 15u)) - 1))];
            if (d & K_EXC) { error = "invalid distance code"; save(); return ERROR; }
          }
          const uint64_t dsaved = bb;
          bb >>= (d & 63u); bc -= (int)(d & 255u);
          const uint32_t dxb = (d >> 8) & 15u;
          const uint32_t dd = (d >> 16) + ((uint32_t)(dsaved >> ((d & 255u) - dxb)) & ((1u << dxb) - 1));
          if (dd > hist + (uint64_t)(out - out_start)) { error = "invalid distance too far back"; save(); return ERROR; }
          const uint8_t* src = out - dd;
          uint8_t* const end = out + len;
          const bool more = in_end - in >= 16 && out_end - end >= 280;
          // The next symbol's entry, before the copy.  A refill only adds bits above the bc valid ones, so with a table index
          // worth of bits in hand the look-up does not wait for the refill's load -- whose address hangs on this match's bit
          // count -- and the two dependency chains (table look-ups, input pointer) run side by side instead of end to end.
          if (more) {
            if (bc >= LIT_TB) { e = lit[bb & ((1u << LIT_TB) - 1)]; refill_fast(bb, bc, in); }
            else { refill_fast(bb, bc, in); e = lit[bb & ((1u << LIT_TB) - 1)]; }
          }
          if (dd >= 8) {
            store64(out, load64(src)); store64(out + 8, load64(src + 8));          // len >= 3; up to 16 bytes at once
            if (len > 16) { uint8_t* o = out + 16; src += 16; do { store64(o, load64(src)); o += 8; src += 8; } while (o < end); }
          } else if (dd == 1) {
            const uint64_t v = 0x0101010101010101ull * src[0];
            uint8_t* o = out; do { store64(o, v); o += 8; } while (o < end);
          } else {
            uint8_t* o = out; do { *o++ = *src++; } while (o < end);
          }
          out = end;
          if (!more) break;
        }
      }
      if (ended) { state = last_block ? 3 : 0; continue; }
      // careful loop near the ends of the buffers: one symbol at a time, every step checked.  A whole symbol (length code,
      // extra bits, distance code, extra bits) is at most 48 bits: unless the input is final it is decoded only with that
      // many in hand, so a lookup never runs on zero-filled bits; with final input every step counts its bits.
      for (;;) {
        refill_safe(bb, bc, in, in_end);
        if (bc < 48 && !final_input) { res = NEED_INPUT; goto out_of_loop; }
        auto starve = [&]() { error = "unexpected end of file"; starved = true; res = NEED_INPUT; };
        uint32_t e = lit[bb & ((1u << LIT_TB) - 1)];
        int used = 0;
        uint64_t b2 = bb;
        if ((e & (K_EXC | K_SUB)) == (K_EXC | K_SUB)) {
          b2 >>= LIT_TB; used += LIT_TB;
          e = lit[(e >> 16) + (uint32_t)(b2 & ((1u << ((e >> 8) & 15u)) - 1))];
        }
        const uint32_t lxb = (e & (K_LIT | K_EXC)) ? 0u : (e >> 8) & 15u;
        const uint32_t lcode = (e & 255u) - lxb;                       // bits of the code alone
        used += (int)(e & 255u);
        if (used > bc || ((e & K_EXC) && (e >> 16) == V_BAD && bc < 15)) { starve(); goto out_of_loop; }
        const uint32_t len = (e >> 16) + ((uint32_t)(b2 >> lcode) & ((1u << lxb) - 1));
        b2 >>= (e & 255u);
        if (e & K_LIT) {
          if (out >= out_end) { res = NEED_OUTPUT; goto out_of_loop; }
          *out++ = (uint8_t)(e >> 16); bb = b2; bc -= used;
          if (in_end - in >= 16 && out_end - out >= 280) break;       // back to the fast loop
          continue;
        }
        if (e & K_EXC) {
          if ((e >> 16) == V_EOB) { bb = b2; bc -= used; state = last_block ? 3 : 0; break; }
          error = "invalid literal/length code"; save(); return ERROR;
        }
        uint32_t d = dist[b2 & ((1u << DIST_TB) - 1)];
        if ((d & (K_EXC | K_SUB)) == (K_EXC | K_SUB)) {
          b2 >>= DIST_TB; used += DIST_TB;
          d = dist[(d >> 16) + (uint32_t)(b2 & ((1u << ((d >> 8) & 15u)) - 1))];
        }
        const uint32_t dxb = (d & K_EXC) ? 0u : (d >> 8) & 15u;
        used += (int)(d & 255u);
        if (used > bc || ((d & K_EXC) && bc < 48)) { starve(); goto out_of_loop; }
        if (d & K_EXC) { error = "invalid distance code"; save(); return ERROR; }
        const uint32_t dd = (d >> 16) + ((uint32_t)(b2 >> ((d & 255u) - dxb)) & ((1u << dxb) - 1));
        b2 >>= (d & 255u);
        if ((uint64_t)(out_end - out) < len) { res = NEED_OUTPUT; goto out_of_loop; }      // the symbol is not consumed: decoded again later
        if (dd > hist + (uint64_t)(out - out_start)) { error = "invalid distance too far back"; save(); return ERROR; }
        bb = b2; bc -= used;
        { const uint8_t* src = out - dd; for (uint32_t k = 0; k < len; ++k) out[k] = src[k]; }
        out += len;
        if (in_end - in >= 16 && out_end - out >= 280) break;
      }
      continue;
    }
  out_of_loop:
    save();
    return res;
  }
};

// gzread's behaviour, kept:
//   * a file that does not start with the gzip magic is passed through as it is ("transparent" mode);
//   * members are concatenated; bytes after the last member that are not a gzip header are ignored;
//   * a stream that ends early (truncated file) yields what could be decoded, then end of file -- not an error;
//   * corrupt data (bad codes, distances, stored lengths, CRC-32 or length mismatch) is an error: read() returns -1.
class GunzipStream {
 public:
  ~GunzipStream() { close(); }
  bool open(const char* path)
  {
    close();
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) return false;
    ibuf_.assign(kIn + 64, 0); in_ = in_end_ = ibuf_.data(); eof_in_ = false;
    wbuf_.assign(kHist + kOut + 320, 0); w_have_ = kHist; w_read_ = kHist; hist_ = 0;
    mode_ = 0; err_.clear(); done_ = false; first_member_ = true;
    return true;
  }
  void close() { if (fd_ >= 0) ::close(fd_); fd_ = -1; }
  const std::string& error() const { return err_; }
  void set_verify(bool on) { verify_ = on; }                 // benchmarks only: skip the CRC-32 / length check of the trailer

  // up to cap bytes of inflated text; 0 at the end of the data, -1 on corrupt input
  long read(uint8_t* dst, size_t cap)
  {
    size_t got = 0;
    while (got < cap) {
      if (w_read_ < w_have_) {
        const size_t n = std::min(cap - got, w_have_ - w_read_);
        memcpy(dst + got, wbuf_.data() + w_read_, n); w_read_ += n; got += n;
        continue;
      }
      if (!err_.empty()) return got ? (long)got : -1;        // like gzread: what was decoded first, the error on the next call
      if (done_) break;
      produce();                                             // output, or progress towards it (input, next member), or done_ / err_
    }
    return (long)got;
  }

 private:
  static constexpr size_t kIn = 1 << 20, kOut = 2 << 20, kHist = 32768;

  void fill_input()
  {
    const size_t left = (size_t)(in_end_ - in_);
    if (left && in_ != ibuf_.data()) memmove(ibuf_.data(), in_, left);
    in_ = ibuf_.data(); in_end_ = in_ + left;
    while (!eof_in_ && (size_t)(in_end_ - ibuf_.data()) < kIn) {
      const ssize_t n = ::read(fd_, const_cast<uint8_t*>(in_end_), kIn - (size_t)(in_end_ - ibuf_.data()));
      if (n < 0) { if (errno == EINTR) continue; eof_in_ = true; break; }
      if (n == 0) { eof_in_ = true; break; }
      in_end_ += n;
    }
    memset(const_cast<uint8_t*>(in_end_), 0, 64);
  }
  size_t avail() const { return (size_t)(in_end_ - in_); }
  bool want(size_t n) { if (avail() < n && !eof_in_) fill_input(); return avail() >= n; }

  // next piece of output into the window buffer; false when nothing was produced (end, error or need to loop)
  bool produce()
  {
    // slide: the last 32 KiB stay as history
    if (w_have_ > kHist) { memmove(wbuf_.data(), wbuf_.data() + w_have_ - kHist, kHist); }
    w_have_ = kHist; w_read_ = kHist;
    if (mode_ == 0) {                                       // ---- start of a member (or of a transparent file) ----
      if (!want(18)) { /* fewer than 18 bytes left */ }
      if (avail() == 0) { done_ = true; return false; }
      if (avail() < 2 || in_[0] != 0x1f || in_[1] != 0x8b) {
        if (first_member_) { mode_ = 3; return produce_raw(); }
        done_ = true; return false;                         // trailing garbage after a member: ignored
      }
      // header: magic, CM, FLG, MTIME(4), XFL, OS, [FEXTRA], [FNAME], [FCOMMENT], [FHCRC]
      if (!parse_header()) return false;
      inf_.reset(); hist_ = 0; crc_ = (uint32_t)crc32(0L, Z_NULL, 0); isize_ = 0; mode_ = 1; first_member_ = false;
    }
    if (mode_ == 3) return produce_raw();
    if (mode_ == 1) {                                       // ---- deflate data ----
      if (avail() < 4096 && !eof_in_) fill_input();
      uint8_t* out = wbuf_.data() + kHist; uint8_t* const out0 = out;
      const uint8_t* in = in_;
      const Inflater::Result r = inf_.run(in, in_end_, eof_in_, out, out0 + kOut, hist_);
      in_ = in;
      const size_t n = (size_t)(out - out0);
      if (n) { if (verify_) crc_ = crc32_update(crc_, out0, n); isize_ += (uint32_t)n; hist_ = std::min<uint64_t>(kHist, hist_ + n); w_have_ = kHist + n; }
      if (r == Inflater::ERROR) { err_ = inf_.error ? inf_.error : "invalid deflate data"; return n != 0; }
      if (r == Inflater::NEED_INPUT) {
        if (eof_in_ && n == 0) { done_ = true; return false; }          // truncated: what was decoded has been delivered (gzread: Z_BUF_ERROR)
        if (!eof_in_) fill_input();
        return n != 0;
      }
      if (r == Inflater::DONE) {
        // unused whole bytes go back to the input, then the trailer: CRC-32 and ISIZE
        const int spare = inf_.bitcnt >> 3;
        in_ -= spare; inf_.bitcnt = 0; inf_.bitbuf = 0;
        mode_ = 2;
      }
      if (n) return true;
    }
    if (mode_ == 2) {                                       // ---- trailer ----
      want(8);
      if (avail() < 4) { done_ = true; return false; }      // truncated inside the trailer: like any early end
      const uint32_t crc = (uint32_t)in_[0] | ((uint32_t)in_[1] << 8) | ((uint32_t)in_[2] << 16) | ((uint32_t)in_[3] << 24);
      if (verify_ && crc != crc_) { err_ = "incorrect data check"; return false; }      // (zlib, too, checks the CRC as soon as its four bytes are there)
      if (avail() < 8) { done_ = true; return false; }
      const uint32_t isz = (uint32_t)in_[4] | ((uint32_t)in_[5] << 8) | ((uint32_t)in_[6] << 16) | ((uint32_t)in_[7] << 24);
      in_ += 8;
      if (isz != isize_) { err_ = "incorrect length check"; return false; }
      mode_ = 0;
      return false;                                         // loop: next member
    }
    return false;
  }

  bool produce_raw()
  {
    if (avail() == 0 && !eof_in_) fill_input();
    const size_t n = std::min(avail(), kOut);
    if (n == 0) { done_ = true; return false; }
    memcpy(wbuf_.data() + kHist, in_, n); in_ += n; w_have_ = kHist + n;
    return true;
  }

  bool parse_header()
  {
    // the fixed part and every optional field must be in the buffer: headers are tiny next to the 1 MiB input buffer
    want(kIn / 2);
    const uint8_t* p = in_; const uint8_t* e = in_end_;
    auto short_file = [&]() { done_ = true; return false; };                       // truncated inside a header: early end
    if (e - p < 10) return short_file();
    if (p[2] != 8) { err_ = "unknown compression method"; return false; }
    const uint8_t flg = p[3];
    if (flg & 0xE0) { err_ = "unknown header flags set"; return false; }
    p += 10;
    if (flg & 4) { if (e - p < 2) return short_file(); const size_t xl = (size_t)p[0] | ((size_t)p[1] << 8); p += 2; if ((size_t)(e - p) < xl) return short_file(); p += xl; }
    if (flg & 8) { while (p < e && *p) ++p; if (p >= e) return short_file(); ++p; }
    if (flg & 16) { while (p < e && *p) ++p; if (p >= e) return short_file(); ++p; }
    if (flg & 2) { if (e - p < 2) return short_file(); p += 2; }
    in_ = p;
    return true;
  }

  int fd_ = -1;
  std::vector<uint8_t> ibuf_, wbuf_;
  const uint8_t* in_ = nullptr; const uint8_t* in_end_ = nullptr;
  bool eof_in_ = false, done_ = false, first_member_ = true, verify_ = true;
  size_t w_have_ = 0, w_read_ = 0;
  uint64_t hist_ = 0;
  int mode_ = 0;                      // 0 member start, 1 deflate data, 2 trailer, 3 transparent
  uint32_t crc_ = 0, isize_ = 0;
  Inflater inf_;
  std::string err_;
};

// The same reader with the inflating on a thread of its own: the caller's thread parses lines while the next 4 MiB are being
// decoded (a plain gzip stream cannot be split, but inflate and parse can overlap: 0.55 + 0.2 s per 316 MB of FASTQ text
// become max(0.55, 0.2)).  Worth it only where cores are spare -- FastqReader asks for it when the box has at least two per
// file.  Same read() contract as GunzipStream, including "what was decoded first, the error on the next call".
class AsyncGunzip {
 public:
  ~AsyncGunzip() { close(); }
  bool open(const char* path)
  {
    close();
    if (!gs_.open(path)) return false;
    for (auto& b : bufs_) { b.data.resize(kBuf); b.n = 0; }
    free_.clear(); ready_.clear();
    for (size_t i = 0; i < kBufs; ++i) free_.push_back(i);
    stop_ = false; cur_ = kNone; cur_pos_ = 0; ended_ = false; failed_ = false;
    try { th_ = std::thread([this] { produce(); }); }
    catch (const std::exception&) { gs_.close(); errno = EAGAIN; return false; }      // no thread to be had: the caller reports the file as unopenable
    return true;
  }
  void close()
  {
    if (th_.joinable()) {
      { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
      cv_.notify_all();
      th_.join();
    }
    gs_.close();
  }
  const std::string& error() const { return gs_.error(); }

  long read(uint8_t* dst, size_t cap)
  {
    size_t got = 0;
    while (got < cap) {
      if (cur_ == kNone) {
        if (ended_) break;
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !ready_.empty(); });
        cur_ = ready_.front(); ready_.erase(ready_.begin()); cur_pos_ = 0;
      }
      Buf& b = bufs_[cur_];
      if (b.n <= 0) {                                        // the producer's last word: end of data (0) or corrupt data (-1)
        ended_ = true; failed_ = b.n < 0; cur_ = kNone;
        break;
      }
      const size_t n = std::min(cap - got, (size_t)b.n - cur_pos_);
      memcpy(dst + got, b.data.data() + cur_pos_, n); got += n; cur_pos_ += n;
      if (cur_pos_ == (size_t)b.n) {
        { std::lock_guard<std::mutex> lk(mu_); free_.push_back(cur_); }
        cv_.notify_all();
        cur_ = kNone;
      }
    }
    if (got == 0 && failed_) return -1;
    return (long)got;
  }

 private:
  static constexpr size_t kBuf = 4 << 20, kBufs = 3, kNone = ~(size_t)0;
  struct Buf { std::vector<uint8_t> data; long n = 0; };
  void produce()
  {
    for (;;) {
      size_t i;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !free_.empty(); });
        if (stop_) return;
        i = free_.back(); free_.pop_back();
      }
      bufs_[i].n = gs_.read(bufs_[i].data.data(), kBuf);
      const bool last = bufs_[i].n <= 0;
      { std::lock_guard<std::mutex> lk(mu_); ready_.push_back(i); }
      cv_.notify_all();
      if (last) return;
    }
  }
  GunzipStream gs_;
  Buf bufs_[kBufs];
  std::vector<size_t> free_, ready_;
  std::mutex mu_; std::condition_variable cv_;
  std::thread th_;
  bool stop_ = false, ended_ = false, failed_ = false;
  size_t cur_ = kNone, cur_pos_ = 0;
};

}  // namespace hgz
