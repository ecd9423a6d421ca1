// swb_inflate.cuh -- raw DEFLATE (RFC 1951) decoder for one small member, written once for two targets:
//   * device: one WARP per BGZF block (<= 64 KiB of text).  All 32 lanes run the same decode loop on the same
//     bits (uniform control flow, broadcast loads); Huffman table fills and LZ77 copies are split across the lanes.
//   * host  : the same code with one "lane", compiled by g++ for the CPU unit test that compares it with zlib
//     (tests/test_inflate_core.py).  No GPU is needed to check the bit-level logic.
// The FASTQ.gz files of a WGS run are the reference's input (aligner.rs:107-120, inflated there by a `zcat` child);
// blocked gzip (BGZF: bgzip / BCL Convert) makes every 64 KiB block an independent member, so a B200 can inflate
// thousands of blocks at once instead of 16 host threads doing it.
//
// Every loop is bounded by the input or the output size; a malformed stream ends with a non-zero status, never a hang.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SWI_HD __device__ __forceinline__               /* nvcc build: device only; the host build is g++ (CPU unit test) */
#define SWI_OUTLINED __device__ __noinline__
#else
#define SWI_HD inline
#define SWI_OUTLINED inline
#endif

namespace swi {

constexpr int LIT_BITS = 10, DIST_BITS = 8;
constexpr int LIT_FAST = 1 << LIT_BITS, DIST_FAST = 1 << DIST_BITS;

enum Status : int { OK = 0, ERR_INPUT_OVERRUN = 1, ERR_OUTPUT_OVERRUN = 2, ERR_BAD_BLOCK_TYPE = 3, ERR_BAD_STORED = 4,
                    ERR_BAD_LENGTHS = 5, ERR_BAD_SYMBOL = 6, ERR_BAD_DISTANCE = 7, ERR_LENGTH_MISMATCH = 8 };

// Decode tables of one member (shared memory on the device, stack on the host): 4 KiB.
// A fast-table entry says everything the symbol loop needs, so it never touches the base / extra-bits tables:
//   [3:0] code length (0 = longer than the table's index: canonical walk)   [7:4] kind   [15:8] extra bits   [31:16] value
//   kind 0 literal (value = byte) | 1 length or distance (value = base) | 2 end of block | 3 not a legal symbol
//   kind 5 (code length 0) a code longer than the table's index: canonical walk.  Lengths / distances are kind 1 so that
//   the symbol loop's usual case -- a match -- is ONE test per table lookup.
constexpr uint32_t kLongCode = 5u << 4;

struct Tables {
  uint32_t lit_fast[LIT_FAST];
  uint32_t dist_fast[DIST_FAST];
  uint16_t lit_count[16], dist_count[16];
  uint16_t lit_sym[288], dist_sym[32];
  uint8_t  lengths[320];
};

struct Lanes { int lane, n; };     // this thread's index among the n threads that cooperate (host: 0 of 1)

#if defined(__CUDA_ARCH__)
#define SWI_SYNC() __syncwarp()
#else
#define SWI_SYNC() ((void)0)
#endif

// Bit reader: a 64-bit WINDOW of the stream (bits [pos, pos+64) relative to the word-aligned base) is loaded with three
// aligned word loads; `used` counts the bits of it already consumed.  Taking bits is a shift of the window and an add to
// `used` -- no 64-bit buffer to shift and refill per symbol.  The caller guarantees 16 readable bytes after the payload
// (device: the compressed arena has slack; host: bounds-checked loads); bits past the payload are never trusted:
// consuming them trips overrun().
struct Bits {
  const uint8_t* p; uint64_t n;    // payload
  const uint32_t* w;               // p rounded down to a word boundary
  uint32_t pos, end, last;         // window start / first bit after the payload, in bits from w; index of the payload's last word
  uint32_t lo, hi, used;
};

SWI_HD uint32_t load_word(const Bits& b, uint32_t idx)
{
#if defined(__CUDA_ARCH__)
  return b.w[idx];
#else
  const uint8_t* q = reinterpret_cast<const uint8_t*>(b.w) + 4ull * idx;
  uint32_t v = 0;
  for (int k = 0; k < 4; ++k) if (q + k >= b.p && q + k < b.p + b.n) v |= (uint32_t)q[k] << (8 * k);
  return v;
#endif
}

// new window at the first unconsumed bit
SWI_HD void refill(Bits& b)
{
  b.pos += b.used; b.used = 0;
  // A malformed stream can run past its payload (the symbol loop is bounded by the output size, not by the input): the
  // window then stops moving at the payload's last word -- its bits are wrong, which no longer matters: overrun() is
  // already true -- so that no load ever lands more than 12 bytes behind the payload.
  const uint32_t idx = b.pos >> 5 < b.last ? b.pos >> 5 : b.last, s = b.pos & 31u;
  const uint32_t w0 = load_word(b, idx), w1 = load_word(b, idx + 1), w2 = load_word(b, idx + 2);
  b.lo = (uint32_t)((((uint64_t)w1 << 32) | w0) >> s);
  b.hi = (uint32_t)((((uint64_t)w2 << 32) | w1) >> s);
}
SWI_HD void init_bits(Bits& b, const uint8_t* in, uint64_t in_len)
{
  const uintptr_t a = (uintptr_t)in;
  b.p = in; b.n = in_len;
  b.w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  b.pos = (uint32_t)(a & 3) * 8u; b.end = b.pos + (uint32_t)in_len * 8u; b.last = b.end >> 5; b.used = 0;
  refill(b);
}
// the 32 stream bits that start `off` bits into the window (off < 64; the top off-32 of them are zero beyond the window)
SWI_HD uint32_t window(const Bits& b, uint32_t off) { return (uint32_t)((((uint64_t)b.hi << 32) | b.lo) >> off); }
SWI_HD uint32_t low_bits(uint32_t v, uint32_t k) { return v & ((1u << k) - 1u); }   // k <= 31
// k <= 16 bits.  The window is renewed whenever fewer than 32 unconsumed bits would remain, so a caller may take up to
// 32 bits between two calls of its own refill().
SWI_HD uint32_t take(Bits& b, int k)
{
  if (b.used > 32) refill(b);
  const uint32_t v = low_bits(window(b, b.used), (uint32_t)k);
  b.used += (uint32_t)k;
  return v;
}
// bits consumed so far must not exceed the payload
SWI_HD bool overrun(const Bits& b) { return b.pos + b.used > b.end; }

SWI_HD uint32_t bitrev(uint32_t c, int len)
{
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) { r = (r << 1) | (c & 1u); c >>= 1; }
  return r;
}

#if defined(__CUDACC__)
#define SWI_CONST static __device__ __constant__        /* nvcc build: only the device side runs this code */
#else
#define SWI_CONST static const                          /* g++ build of the CPU unit test */
#endif
SWI_CONST uint16_t kLenBase[29]  = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
SWI_CONST uint8_t  kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
SWI_CONST uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193,
                                    12289, 16385, 24577};
SWI_CONST uint8_t  kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
SWI_CONST uint8_t  kClOrder[19]  = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum Alphabet : int { ALPHA_PLAIN = 0, ALPHA_LITLEN = 1, ALPHA_DIST = 2 };

// what the symbol loop needs to know about symbol s of the given alphabet (see Tables)
SWI_HD uint32_t make_entry(int alphabet, int s, int len)
{
  if (alphabet == ALPHA_LITLEN) {
    if (s < 256) return ((uint32_t)s << 16) | (uint32_t)len;
    if (s == 256) return (2u << 4) | (uint32_t)len;
    if (s <= 285) return ((uint32_t)kLenBase[s - 257] << 16) | ((uint32_t)kLenExtra[s - 257] << 8) | (1u << 4) | (uint32_t)len;
    return (3u << 4) | (uint32_t)len;
  }
  if (alphabet == ALPHA_DIST) {
    if (s <= 29) return ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 8) | (1u << 4) | (uint32_t)len;
    return (3u << 4) | (uint32_t)len;
  }
  return ((uint32_t)s << 16) | (uint32_t)len;
}

// Canonical Huffman tables from code lengths.  Returns false on an over-subscribed set.
// (An incomplete set is accepted, as zlib accepts the single-code distance tree; unused codes decode as errors.)
// Kept out of line on the device: inlined three times into inflate_member, nvcc 12.9 -O3 produced a build() that
// reported a valid fixed-code length set as over-subscribed (tools/inflate_gpu_probe.cu reproduces it; -G, a printf in
// the loop or __noinline__ all make it correct).  It runs two or three times per deflate block, the call costs nothing.
SWI_OUTLINED bool build(const uint8_t* lengths, int n, int alphabet, uint32_t* fast, int fast_bits, uint16_t* count, uint16_t* sym, const Lanes& L)
{
  const int fast_size = 1 << fast_bits;
  for (int k = L.lane; k < fast_size; k += L.n) fast[k] = kLongCode;
  uint32_t cnt[16];
  for (int l = 0; l < 16; ++l) cnt[l] = 0;
  for (int s = 0; s < n; ++s) ++cnt[lengths[s]];
  int left = 1;
  for (int l = 1; l < 16; ++l) { left <<= 1; left -= (int)cnt[l]; if (left < 0) return false; }
  uint32_t offs[16], next_code[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + cnt[l];
  uint32_t code = 0;
  for (int l = 1; l < 16; ++l) { next_code[l] = code; code = (code + cnt[l]) << 1; }
  SWI_SYNC();                                       // the cleared table before the fills, the old tables before the new
  if (L.lane == 0) for (int l = 0; l < 16; ++l) count[l] = (uint16_t)(l ? cnt[l] : 0);
  for (int s = 0; s < n; ++s) {
    const int l = lengths[s];
    if (!l) continue;
    if (L.lane == 0) sym[offs[l]] = (uint16_t)s;
    ++offs[l];
    const uint32_t c = next_code[l]++;
    if (l <= fast_bits) {
      const uint32_t rev = bitrev(c, l), e = make_entry(alphabet, s, l);
      for (uint32_t k = rev + ((uint32_t)L.lane << l); k < (uint32_t)fast_size; k += (uint32_t)L.n << l) fast[k] = e;
    }
  }
  SWI_SYNC();
  return true;
}

// An LZ77 copy that has been decoded but not carried out yet (see the symbol loop): len bytes from out[from..] to out[at..].
struct Pending { uint32_t len, from, at; };

// The bytes of a short pending copy between its loads and its stores.  Device: lane i holds byte i; host: all of them.
constexpr uint32_t kShortCopy = 32;
#if defined(__CUDA_ARCH__)
struct CopyRegs { uint8_t v; };
SWI_HD void short_load(CopyRegs& r, const uint8_t* out, const Pending& p, const Lanes& L) { if ((uint32_t)L.lane < p.len) r.v = out[p.from + L.lane]; }
SWI_HD void short_store(const CopyRegs& r, uint8_t* out, const Pending& p, const Lanes& L) { if ((uint32_t)L.lane < p.len) out[p.at + L.lane] = r.v; }
#else
struct CopyRegs { uint8_t v[kShortCopy]; };
SWI_HD void short_load(CopyRegs& r, const uint8_t* out, const Pending& p, const Lanes&) { for (uint32_t i = 0; i < p.len; ++i) r.v[i] = out[p.from + i]; }
SWI_HD void short_store(const CopyRegs& r, uint8_t* out, const Pending& p, const Lanes&) { for (uint32_t i = 0; i < p.len; ++i) out[p.at + i] = r.v[i]; }
#endif

// out[at, at+len) = the len bytes that follow out[from], which repeat with period at-from when the ranges overlap
SWI_HD void copy_match(uint8_t* out, const Pending& p, const Lanes& L)
{
  const uint8_t* src = out + p.from;
  uint8_t* dst = out + p.at;
  const uint32_t dist = p.at - p.from;
  if (dist >= p.len) {                                     // no overlap (the usual case): a plain strided copy
    for (uint32_t i = L.lane; i < p.len; i += L.n) dst[i] = src[i];
  } else if (dist == 1) {                                  // a run of one byte (quality strings)
    const uint8_t v1 = src[0];
    for (uint32_t i = L.lane; i < p.len; i += L.n) dst[i] = v1;
  } else {                                                 // the pattern of `dist` bytes repeats: read only what existed before
    for (uint32_t i = L.lane; i < p.len; i += L.n) dst[i] = src[i % dist];
  }
}

// Canonical bit-by-bit walk for a code longer than the fast table's index; v = the stream bits that start at the code.
// An unused code comes back as kind 3 with code length 0.
SWI_HD uint32_t decode_slow(uint32_t v, int alphabet, const uint16_t* count, const uint16_t* sym)
{
  int code = 0, first = 0, index = 0;
  for (int len = 1; len <= 15; ++len) {
    code |= (int)(v & 1u); v >>= 1;
    const int c = count[len];
    if (code - c < first) return make_entry(alphabet, sym[index + (code - first)], len);
    index += c; first += c; first <<= 1; code <<= 1;
  }
  return 3u << 4;
}

// One symbol: its table entry (code bits consumed).
SWI_HD uint32_t decode(Bits& b, int alphabet, const uint32_t* fast, int fast_bits, const uint16_t* count, const uint16_t* sym)
{
  if (b.used > 32) refill(b);
  const uint32_t v = window(b, b.used);
  uint32_t e = fast[low_bits(v, (uint32_t)fast_bits)];
  if (e == kLongCode) e = decode_slow(v, alphabet, count, sym);
  b.used += e & 15u;
  return e;
}

// Inflate one raw-deflate member of in_len bytes into out[0, out_cap).  *produced = bytes written.
SWI_HD int inflate_member(const uint8_t* in, uint64_t in_len, uint8_t* out, uint32_t out_cap, uint32_t* produced, Tables& T, const Lanes& L)
{
  *produced = 0;
  if (in_len >= (1ull << 28)) return ERR_INPUT_OVERRUN;       // bit positions are 32-bit; a BGZF member is < 64 KiB
  Bits b;
  init_bits(b, in, in_len);
  uint32_t opos = 0;
  int status = OK;
  for (int blocks = 0; blocks < 1 << 20; ++blocks) {          // a member of <= 64 KiB never has this many deflate blocks
    const uint32_t bfinal = take(b, 1), btype = take(b, 2);
    if (btype == 0) {                                          // stored
      b.pos = (b.pos + b.used + 7u) & ~7u; b.used = 0;         // to the next byte boundary (w is word aligned, so this is one of p too)
      refill(b);
      const uint32_t len = take(b, 16), nlen = take(b, 16);
      if ((len ^ nlen) != 0xFFFFu) { status = ERR_BAD_STORED; break; }
      if (opos + len > out_cap) { status = ERR_OUTPUT_OVERRUN; break; }
      const uint64_t src = (uint64_t)((b.pos + b.used) >> 3) - (uint64_t)(in - reinterpret_cast<const uint8_t*>(b.w));
      if (src + len > in_len) { status = ERR_INPUT_OVERRUN; break; }
      for (uint32_t i = L.lane; i < len; i += L.n) out[opos + i] = in[src + i];
      opos += len;
      b.pos += b.used + 8u * len; b.used = 0;
      refill(b);
    } else if (btype == 1 || btype == 2) {
      int nlit, ndist;
      if (btype == 1) {                                        // fixed code
        for (int s = L.lane; s < 288; s += L.n) T.lengths[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
        for (int s = L.lane; s < 30; s += L.n) T.lengths[288 + s] = 5;
        SWI_SYNC();
        nlit = 288; ndist = 30;
      } else {                                                 // dynamic code
        nlit = (int)take(b, 5) + 257; ndist = (int)take(b, 5) + 1;
        const int ncl = (int)take(b, 4) + 4;
        if (nlit > 286 || ndist > 30) { status = ERR_BAD_LENGTHS; break; }
        uint8_t cl[19];
        for (int i = 0; i < 19; ++i) cl[i] = 0;
        for (int i = 0; i < ncl; ++i) cl[kClOrder[i]] = (uint8_t)take(b, 3);
        // the code-length code is tiny: a 7-bit fast table in the distance table's place
        if (!build(cl, 19, ALPHA_PLAIN, T.dist_fast, 7, T.dist_count, T.dist_sym, L)) { status = ERR_BAD_LENGTHS; break; }
        int i = 0;
        uint8_t prev = 0;
        bool bad = false;
        while (i < nlit + ndist) {
          const uint32_t ce = decode(b, ALPHA_PLAIN, T.dist_fast, 7, T.dist_count, T.dist_sym);
          if ((ce >> 4) & 15u) { bad = true; break; }
          const int s = (int)(ce >> 16);
          if (s < 16) { if (L.lane == 0) T.lengths[i] = (uint8_t)s; prev = (uint8_t)s; ++i; continue; }
          int rep; uint8_t v = 0;
          if (s == 16) { if (i == 0) { bad = true; break; } v = prev; rep = 3 + (int)take(b, 2); }
          else if (s == 17) rep = 3 + (int)take(b, 3);
          else rep = 11 + (int)take(b, 7);
          if (i + rep > nlit + ndist) { bad = true; break; }
          if (L.lane == 0) for (int k = 0; k < rep; ++k) T.lengths[i + k] = v;
          i += rep; prev = v;
        }
        if (bad || overrun(b)) { status = bad ? ERR_BAD_LENGTHS : ERR_INPUT_OVERRUN; break; }
        SWI_SYNC();
        if (T.lengths[256] == 0) { status = ERR_BAD_LENGTHS; break; }
        // distance lengths follow the literal/length lengths in the same array: move them to a fixed place
        uint8_t dl[30];
        for (int s = 0; s < 30; ++s) dl[s] = s < ndist ? T.lengths[nlit + s] : 0;
        SWI_SYNC();
        for (int s = L.lane; s < 30; s += L.n) T.lengths[288 + s] = dl[s];
        for (int s = nlit + L.lane; s < 288; s += L.n) T.lengths[s] = 0;
        SWI_SYNC();
        nlit = 288; ndist = 30;
      }
      if (!build(T.lengths, nlit, ALPHA_LITLEN, T.lit_fast, LIT_BITS, T.lit_count, T.lit_sym, L) ||
          !build(T.lengths + 288, ndist, ALPHA_DIST, T.dist_fast, DIST_BITS, T.dist_count, T.dist_sym, L)) { status = ERR_BAD_LENGTHS; break; }
      // ---- symbols ----
      // A literal/length code with its extra bits is at most 20 bits, a distance code with its extra bits at most 28: with
      // no more than 16 bits of the 64-bit window consumed, a whole match decodes from one window.
      // The copy of a match is software-pipelined round the symbol loop: a match decoded in one pass is LOADED at the top
      // of the next pass and STORED at its bottom, with the decoding of the next symbol in between.  In FASTQ text a match
      // is 3-8 bases found up to 32 KiB back -- an L2 round trip the warp would otherwise sit out once per match.  Loads
      // and stores of one copy stay inside one pass (no load is in flight across the back edge, where the compiler would
      // wait for it), copies run in stream order, each after a __syncwarp that makes everything written so far visible.
#if defined(__CUDA_ARCH__)
      const bool pipelined = L.n == (int)kShortCopy;           // one byte per lane (the one-lane debug kernel copies in place)
#else
      const bool pipelined = true;
#endif
      Pending P; P.len = 0; P.from = 0; P.at = 0;
      bool p_short = false;                                    // the pending copy is short and does not overlap itself: pipelined
      CopyRegs R = {};
      for (;;) {                                               // every pass emits >= 1 byte (bounded by out_cap), ends the block or fails
        if (P.len) {
          SWI_SYNC();
          if (p_short) short_load(R, out, P, L);
          else { copy_match(out, P, L); P.len = 0; }
        }
        if (b.used > 16) refill(b);
        const uint32_t v = window(b, b.used);
        uint32_t e = T.lit_fast[v & (LIT_FAST - 1)];
        if (((e >> 4) & 15u) != 1u) {                            // not a length: a literal, the end of the block, a long code
          if (e == kLongCode) e = decode_slow(v, ALPHA_LITLEN, T.lit_count, T.lit_sym);
          const uint32_t kind = (e >> 4) & 15u;
          if (kind == 0) {
            if (opos >= out_cap) { status = ERR_OUTPUT_OVERRUN; break; }
            out[opos] = (uint8_t)(e >> 16);                      // every lane stores the same byte to the same address: no branch
            ++opos;
            b.used += e & 15u;
            if (P.len) { short_store(R, out, P, L); P.len = 0; }
            continue;
          }
          if (kind != 1) { b.used += e & 15u; if (kind != 2) status = ERR_BAD_SYMBOL; break; }
        }
        const uint32_t cl = e & 15u, xb = (e >> 8) & 255u;
        const uint32_t len = (e >> 16) + low_bits(v >> cl, xb);
        const uint32_t at = b.used + cl + xb;                    // <= 36
        const uint32_t d = window(b, at);
        uint32_t de = T.dist_fast[d & (DIST_FAST - 1)];
        if (de == kLongCode) de = decode_slow(d, ALPHA_DIST, T.dist_count, T.dist_sym);
        const uint32_t dcl = de & 15u, dxb = (de >> 8) & 255u;
        const uint32_t dist = (de >> 16) + low_bits(d >> dcl, dxb);
        b.used = at + dcl + dxb;
        if ((((de >> 4) & 15u) != 1) | (dist > opos) | (opos + len > out_cap)) {      // one branch for the three ways a match can be wrong
          status = ((de >> 4) & 15u) != 1 || dist > opos ? ERR_BAD_DISTANCE : ERR_OUTPUT_OVERRUN;
          break;
        }
        if (P.len) short_store(R, out, P, L);
        P.len = len; P.from = opos - dist; P.at = opos;
        p_short = (len <= kShortCopy) & (dist >= len) & pipelined;
        opos += len;
      }
      // every way out of the loop is a break below its top, so a copy still pending here has been loaded, not stored
      if (P.len) short_store(R, out, P, L);                    // (also on the error paths: `produced` bytes are really there)
      SWI_SYNC();
      if (status != OK) break;
      // one input check per deflate block: the bits past the payload are not the stream's, the symbol loop above still
      // ends (it is bounded by out_cap), and the block is rejected here
      if (overrun(b)) { status = ERR_INPUT_OVERRUN; break; }
    } else { status = ERR_BAD_BLOCK_TYPE; break; }
    if (bfinal) break;
  }
  *produced = opos;
  return status;
}

}  // namespace swi
