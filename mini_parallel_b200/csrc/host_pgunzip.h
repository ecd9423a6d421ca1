// host_pgunzip.h -- ONE plain gzip file inflated by SEVERAL host threads (hgz::ParallelGunzip), for the host FASTQ path where
// cores are spare: --test-wgs / count_bases_in_fastq on one file, gpu_align_pair on two, --full-wgs on a box with more cores
// than twice its files.  The reference hands every file to one `zcat` child (aligner.rs:109-120); host_gunzip.h made that an
// in-process decoder, but a deflate stream is serial by construction: a block can start at any BIT, and every match may
// reach 32 KiB back into text that has not been decoded yet.  What lifts both (the two-pass scheme of pugz / rapidgzip):
//   * the compressed file (mmap) is cut into chunks of 1 MiB.  A worker looks for the first bit position in its chunk that
//     parses as the header of a non-final dynamic-Huffman block with complete codes, and decodes from there -- to the first
//     block boundary at or after the chunk's end -- into 16-BIT symbols: a byte, or a MARKER "byte i of the 32 KiB in front
//     of my start", which is what a match reaching behind the start copies (the symbol buffer starts with the 32 768 markers,
//     so a copy needs no special case and markers propagate through copies of copies);
//   * the caller's thread walks the chain: a chunk's symbols are accepted only when the chunk STARTED on the very bit where
//     the text known so far ENDS -- then they are what a serial decoder would have produced from there, whatever the block
//     finder believed -- the last 32 KiB are resolved on the spot (the next chunk's window), the rest by a worker (one table
//     look-up per symbol, CRC-32 of the chunk beside it);
//   * wherever the chain does not close -- a chunk without a dynamic block, a false positive, a stored / fixed / final block
//     on a chunk border, a member border, a decoding error, the end of a truncated file -- the caller's thread decodes from
//     the last certain bit with hgz::Inflater, the serial decoder, up to the next chunk that might fit.  So every byte,
//     every error and every early end is either the serial decoder's own or the replay of a clean decode from a certain
//     position: behaviour at the edges is GunzipStream's (= gzread's) by construction, and tests/test_host_pgunzip.py holds
//     the two against each other on every kind of stream, chunk sizes down to 512 bytes, and mutated files.
//   * a file that begins with a BGZF member needs neither markers nor a block finder: a member's size stands in its header and
//     it has no history.  The workers inflate runs of whole members into bytes (CRC-32 and ISIZE of every member checked by
//     the worker) and the chain takes a run when it starts at the header it stands in front of; everything else as above.
// Files that are not regular files, not gzip, or smaller than three chunks go to GunzipStream unchanged.
#pragma once
#include "host_gunzip.h"
#include <sys/mman.h>
#include <sys/stat.h>
#include <atomic>
#include <cstdlib>
#include <deque>

#define HGZ_HAVE_PARALLEL 1

namespace hgz {

constexpr uint32_t kWin = 32768;
constexpr uint16_t kMarker = 0x8000;                     // symbol = kMarker | index into the 32 KiB in front of the chunk

// ---- what a worker makes of one chunk ----
template <class T> struct RawBuf {   // uninitialised, growable (std::vector would zero-fill tens of megabytes per chunk)
  T* p = nullptr; size_t cap = 0;
  RawBuf() = default;
  RawBuf(const RawBuf&) = delete; RawBuf& operator=(const RawBuf&) = delete;
  RawBuf(RawBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
  RawBuf& operator=(RawBuf&& o) noexcept { if (this != &o) { std::free(p); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; } return *this; }
  ~RawBuf() { std::free(p); }
  void reserve(size_t n) { if (n <= cap) return; void* q = std::realloc(p, n * sizeof(T)); if (!q) throw std::bad_alloc(); p = (T*)q; cap = n; }
  void release() { std::free(p); p = nullptr; cap = 0; }
};

struct SpecChunk {
  bool clean = false;              // decoded without a complaint from start_bit to end_bit
  bool final_block = false;        // ... and end_bit is the first bit after the member's final block
  uint64_t start_bit = 0, end_bit = 0;
  RawBuf<uint16_t> sym;            // kWin marker slots, then the n symbols
  size_t n = 0;
  // A run of BGZF members instead (bgzf_run): whole members from the header at start_byte to the header at end_byte, inflated
  // to n BYTES in `bytes`, every member's CRC-32 and ISIZE checked by the worker.  No markers: a member has no history.
  bool bgzf_run = false;
  uint64_t start_byte = 0, end_byte = 0, members = 0;
  RawBuf<uint8_t> bytes;
};

// The member header bgzip writes: 1f 8b 08 04, MTIME XFL OS, XLEN = 6, 'B' 'C' 02 00, BSIZE (member size - 1).  (BGZF allows other
// extra subfields beside BC; files that carry them are simply not recognised here and take the serial reader.)
static inline bool bgzf_header_at(const uint8_t* base, uint64_t size, uint64_t p, uint64_t* next)
{
  if (p + 18 + 8 > size) return false;
  const uint8_t* h = base + p;
  if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || h[3] != 4 || h[10] != 6 || h[11] != 0 || h[12] != 'B' || h[13] != 'C' || h[14] != 2 || h[15] != 0) return false;
  const uint64_t total = ((uint64_t)h[16] | ((uint64_t)h[17] << 8)) + 1;
  if (total < 18 + 2 + 8 || p + total > size) return false;
  *next = p + total;
  return true;
}

// Inflate the BGZF members from the header at `first` up to the first member that starts at or after `limit` (or that does not
// check out: the run ends in front of it).  R.clean = at least one member, all of them with the right CRC-32 and length.
static inline void bgzf_run_decode(Inflater& T, const uint8_t* base, uint64_t size, uint64_t first, uint64_t limit, bool verify, SpecChunk& R)
{
  R.clean = false; R.bgzf_run = true; R.start_byte = R.end_byte = first; R.n = 0; R.members = 0;
  uint64_t total = 0, q = first, next = 0, count = 0;
  while (q < limit && bgzf_header_at(base, size, q, &next)) {              // sizes first: one buffer for the run
    const uint8_t* t = base + next - 8;
    const uint32_t isz = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    if (isz > 65536) break;                                                // not a BGZF member (they hold <= 64 KiB)
    total += isz; q = next; ++count;
  }
  if (!count) return;
  R.bytes.reserve(total + 512);
  uint8_t* out = R.bytes.p; uint8_t* const out_end = out + total + 320;
  q = first;
  for (uint64_t m = 0; m < count; ++m) {
    bgzf_header_at(base, size, q, &next);
    const uint8_t* t = base + next - 8;
    const uint32_t crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    const uint32_t isz = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    T.reset();
    const uint8_t* in = base + q + 18; uint8_t* o = out;
    const Inflater::Result r = T.run(in, t, true, o, out_end, 0);
    const bool ok = r == Inflater::DONE && in - (T.bitcnt >> 3) == t && (uint64_t)(o - out) == isz && (!verify || crc32_update(0u, out, isz) == crc);
    if (!ok) break;                                                        // the serial reader will say what is wrong with this member
    out += isz; q = next; ++R.members;
  }
  if (!R.members) return;
  R.n = (size_t)(out - R.bytes.p); R.end_byte = q; R.clean = true;
}

// Is bit p the start of a non-final dynamic block whose three codes are acceptable to zlib?  On success T holds its tables.
static inline bool dynamic_header_at(Inflater& T, const uint8_t* base, uint64_t size, uint64_t p)
{
  const uint64_t byte = p >> 3; const int sh = (int)(p & 7);
  if (byte + 24 > size) return false;                     // too close to the end of the file: left to the serial decoder
  const uint64_t w = load64(base + byte) >> sh;           // >= 57 bits
  if ((w & 7u) != 4u) return false;                       // BFINAL = 0, BTYPE = 2
  if (((w >> 3) & 31u) > 29u || ((w >> 8) & 31u) > 29u) return false;
  const int ncl = (int)((w >> 13) & 15u) + 4;
  // the code-length code: ncl x 3 bits from bit p + 17; it must be complete (zlib rejects anything else)
  uint64_t c = load64(base + byte + 2) >> sh;             // bits from p + 16
  if (sh) c |= load64(base + byte + 10) << (64 - sh);
  c >>= 1;
  int kraft = 0;
  for (int i = 0; i < ncl; ++i) { const int l = (int)(c & 7u); c >>= 3; if (l) kraft += 128 >> l; }
  if (kraft != 128) return false;
  T.reset();
  const uint8_t* in = T.start_at_bit(base, p + 3);
  uint64_t bb = T.bitbuf; int bc = T.bitcnt;
  const Inflater::Result r = T.dynamic_tables(bb, bc, in, base + size);
  T.error = nullptr; T.starved = false;
  return r == Inflater::DONE;
}

// Decode from start_bit (a block header) to the first block boundary at or after stop_bit, or to the end of the final block.
// Anything unusual -- an error, running out of input, more than max_syms symbols -- leaves R.clean false: the serial decoder
// will go over the same bits and say what is wrong with them.
static inline void spec_decode(Inflater& T, const uint8_t* base, uint64_t size, uint64_t start_bit, uint64_t stop_bit, size_t max_syms, SpecChunk& R)
{
  R.clean = false; R.final_block = false; R.start_bit = start_bit; R.n = 0;
  const uint8_t* const in_end = base + size;
  T.reset();
  const uint8_t* in = T.start_at_bit(base, start_bit);
  uint64_t bb = T.bitbuf; int bc = T.bitcnt;
  uint64_t pad = 0;                                       // zero bytes fed beyond the end of the file
  if (stop_bit <= start_bit) return;
  size_t cap = std::min(max_syms, (size_t)((stop_bit - start_bit) >> 3) * 6 + 65536);
  R.sym.reserve(kWin + cap + 512);
  for (uint32_t i = 0; i < kWin; ++i) R.sym.p[i] = (uint16_t)(kMarker | i);
  uint16_t* out = R.sym.p + kWin;
  uint16_t* out_end = out + cap;
  auto refill = [&]() {
    if (in_end - in >= 8) { bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; }
    else while (bc < 56) { if (in < in_end) bb |= (uint64_t)*in++ << bc; else ++pad; bc += 8; }
  };
  auto bitpos = [&]() { return (uint64_t)(in - base) * 8 + pad * 8 - (uint64_t)bc; };
  auto grow = [&]() -> bool {
    const size_t have = (size_t)(out - (R.sym.p + kWin));
    if (cap >= max_syms) return false;
    cap = std::min(max_syms, cap * 2);
    R.sym.reserve(kWin + cap + 512);
    out = R.sym.p + kWin + have; out_end = R.sym.p + kWin + cap;
    return true;
  };
  for (;;) {
    // ---- block header ----
    const uint64_t here = bitpos();
    if (here > size * 8) return;
    if (here >= stop_bit && here != start_bit) { R.end_bit = here; R.clean = true; break; }
    refill();
    const bool last = bb & 1u; const uint32_t type = (uint32_t)(bb >> 1) & 3u;
    bb >>= 3; bc -= 3;
    if (type == 0) {
      const int drop = bc & 7; bb >>= drop; bc -= drop;
      refill();
      const uint32_t len = (uint32_t)(bb & 0xFFFFu), nlen = (uint32_t)((bb >> 16) & 0xFFFFu);
      bb >>= 32; bc -= 32;
      if ((len ^ nlen) != 0xFFFFu || pad) return;
      in -= bc >> 3; bb = 0; bc = 0;                       // whole bytes in the bit buffer go back: the stored bytes are read in place
      if ((uint64_t)(in_end - in) < len) return;
      while ((size_t)(out_end - out) < len + 320u) if (!grow()) return;
      for (uint32_t i = 0; i < len; ++i) out[i] = in[i];
      out += len; in += len;
    } else if (type == 1 || type == 2) {
      if (type == 1) T.fixed_tables();
      else {
        const Inflater::Result r = T.dynamic_tables(bb, bc, in, in_end);
        if (r != Inflater::DONE) return;
      }
      const uint32_t* const lit = T.lit; const uint32_t* const dist = T.dist;
      constexpr uint32_t LM = (1u << LIT_TB) - 1, DM = (1u << DIST_TB) - 1;
      // Fast loop while >= 16 bytes of input lie ahead (all but the last symbols of the file): the same shape as the serial
      // decoder's -- the next symbol's table entry is fetched before the match is copied and, with an index worth of bits in
      // hand, before the refill, so the refill's load is off the look-ups' dependency chain.
      bool eob = false;
      if (in_end - in >= 16) {
        bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56;
        uint32_t e = lit[bb & LM];
        for (;;) {
          if (out_end - out < 320) { if (!grow()) return; }
          if (e & K_LIT) {
            bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
            e = lit[bb & LM];
            if (e & K_LIT) {
              bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
              e = lit[bb & LM];
              if (e & K_LIT) {
                bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
                e = lit[bb & LM];
              }
            }
            if (in_end - in < 16) break;
            bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56;      // (e looked at the low bits only: still the next symbol's entry)
            continue;
          }
          if (e & K_EXC) {
            if (e & K_SUB) {
              bb >>= LIT_TB; bc -= LIT_TB;
              e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 15u)) - 1))];
              if (e & K_LIT) {
                bb >>= (e & 63u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
                if (in_end - in < 16) break;
                bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56;
                e = lit[bb & LM];
                continue;
              }
            }
            if (e & K_EXC) {
              bb >>= (e & 63u); bc -= (int)(e & 255u);
              if ((e >> 16) == V_EOB && !(e & K_SUB)) { eob = true; break; }
              return;                                      // not a legal symbol
            }
          }
          const uint64_t lsaved = bb;
          bb >>= (e & 63u); bc -= (int)(e & 255u);
          uint32_t d = dist[bb & DM];
          const uint32_t lxb = (e >> 8) & 15u;
          const uint32_t len = (e >> 16) + ((uint32_t)(lsaved >> ((e & 255u) - lxb)) & ((1u << lxb) - 1));
          if (d & K_EXC) {
            if (!(d & K_SUB)) return;
            bb >>= DIST_TB; bc -= DIST_TB;
            d = dist[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 15u)) - 1))];
            if (d & K_EXC) return;
          }
          const uint64_t dsaved = bb;
          bb >>= (d & 63u); bc -= (int)(d & 255u);
          const uint32_t dxb = (d >> 8) & 15u;
          const uint32_t dd = (d >> 16) + ((uint32_t)(dsaved >> ((d & 255u) - dxb)) & ((1u << dxb) - 1));   // <= 32768: inside the marker slots at worst
          const bool more = in_end - in >= 16;
          if (more) {
            if (bc >= LIT_TB) { e = lit[bb & LM]; bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; }
            else { bb |= load64(in) << bc; in += (63 - bc) >> 3; bc |= 56; e = lit[bb & LM]; }
          }
          const uint16_t* src = out - dd;
          uint16_t* const end = out + len;
          if (dd >= 8) {                                   // 8 symbols = 16 bytes at a time
            uint16_t* o = out;
            do { uint64_t a = 0, b = 0; memcpy(&a, src, 8); memcpy(&b, src + 4, 8); memcpy(o, &a, 8); memcpy(o + 4, &b, 8); o += 8; src += 8; } while (o < end);
          } else if (dd == 1) {
            const uint64_t v = 0x0001000100010001ull * src[0];
            uint16_t* o = out; do { memcpy(o, &v, 8); o += 4; } while (o < end);
          } else {
            uint16_t* o = out; do { *o++ = *src++; } while (o < end);
          }
          out = end;
          if (!more) break;
        }
      }
      // the last symbols of the file, one at a time with the careful refill
      if (!eob)
      for (;;) {
        if (out_end - out < 320) { if (!grow()) return; }
        refill();
        if (pad > 16) return;                              // far beyond the end of the file
        uint32_t e = lit[bb & LM];
        if (e & K_LIT) {                                   // up to three literals per refill (3 x 15 bits)
          bb >>= (e & 255u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
          e = lit[bb & LM];
          if (e & K_LIT) {
            bb >>= (e & 255u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16);
            e = lit[bb & LM];
            if (e & K_LIT) { bb >>= (e & 255u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16); }
          }
          continue;
        }
        if (e & K_EXC) {
          if (e & K_SUB) {
            bb >>= LIT_TB; bc -= LIT_TB;
            e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 15u)) - 1))];
            if (e & K_LIT) { bb >>= (e & 255u); bc -= (int)(e & 255u); *out++ = (uint16_t)(e >> 16); continue; }
          }
          if (e & K_EXC) {
            bb >>= (e & 255u); bc -= (int)(e & 255u);
            if ((e >> 16) == V_EOB && !(e & K_SUB)) break;
            return;                                        // not a legal symbol
          }
        }
        const uint64_t lsaved = bb;
        bb >>= (e & 255u); bc -= (int)(e & 255u);
        uint32_t d = dist[bb & DM];
        const uint32_t lxb = (e >> 8) & 15u;
        const uint32_t len = (e >> 16) + ((uint32_t)(lsaved >> ((e & 255u) - lxb)) & ((1u << lxb) - 1));
        if (d & K_EXC) {
          if (!(d & K_SUB)) return;
          bb >>= DIST_TB; bc -= DIST_TB;
          d = dist[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 15u)) - 1))];
          if (d & K_EXC) return;
        }
        const uint64_t dsaved = bb;
        bb >>= (d & 255u); bc -= (int)(d & 255u);
        const uint32_t dxb = (d >> 8) & 15u;
        const uint32_t dd = (d >> 16) + ((uint32_t)(dsaved >> ((d & 255u) - dxb)) & ((1u << dxb) - 1));   // <= 32768: inside the marker slots at worst
        const uint16_t* src = out - dd;
        uint16_t* const end = out + len;
        if (dd >= 8) {                                     // 8 symbols = 16 bytes at a time
          uint16_t* o = out;
          do { uint64_t a = 0, b = 0; memcpy(&a, src, 8); memcpy(&b, src + 4, 8); memcpy(o, &a, 8); memcpy(o + 4, &b, 8); o += 8; src += 8; } while (o < end);
        } else if (dd == 1) {
          const uint64_t v = 0x0001000100010001ull * src[0];
          uint16_t* o = out; do { memcpy(o, &v, 8); o += 4; } while (o < end);
        } else {
          uint16_t* o = out; do { *o++ = *src++; } while (o < end);
        }
        out = end;
      }
    } else return;
    if (bitpos() > size * 8) return;                       // the block ended on bits the file does not have
    if (last) { R.end_bit = bitpos(); R.final_block = true; R.clean = true; break; }
  }
  R.n = (size_t)(out - (R.sym.p + kWin));
}

// Gzip member header at p: 1 parsed (*next = first byte of the deflate data), 0 truncated inside it, -1 not acceptable (*why)
static inline int parse_gzip_header(const uint8_t* p, const uint8_t* e, const uint8_t** next, const char** why)
{
  if (e - p < 10) return 0;
  if (p[2] != 8) { *why = "unknown compression method"; return -1; }
  const uint8_t flg = p[3];
  if (flg & 0xE0) { *why = "unknown header flags set"; return -1; }
  p += 10;
  if (flg & 4) { if (e - p < 2) return 0; const size_t xl = (size_t)p[0] | ((size_t)p[1] << 8); p += 2; if ((size_t)(e - p) < xl) return 0; p += xl; }
  if (flg & 8) { while (p < e && *p) ++p; if (p >= e) return 0; ++p; }
  if (flg & 16) { while (p < e && *p) ++p; if (p >= e) return 0; ++p; }
  if (flg & 2) { if (e - p < 2) return 0; p += 2; }
  *next = p;
  return 1;
}

class ParallelGunzip {
 public:
  ~ParallelGunzip() { close(); }
  void set_threads(unsigned n) { threads_ = n ? n : 1; }                         // decoder threads beside the caller's; <= 1: GunzipStream
  void set_chunk_bytes(uint64_t n) { chunk_ = std::max<uint64_t>(n, 64); }      // tests: tiny chunks on small files
  void set_verify(bool on) { verify_ = on; serial_.set_verify(on); }
  const std::string& error() const { return use_serial_ ? serial_.error() : err_; }
  bool parallel() const { return !use_serial_; }                                 // false: the file went to GunzipStream
  // what the chain did (tests, SWB_DEBUG): chunks whose symbols were accepted / stretches decoded serially by the caller
  uint64_t chunks_accepted() const { return n_accepted_; }
  uint64_t serial_stretches() const { return n_fallbacks_; }

  bool open(const char* path)
  {
    close();
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    bool par = threads_ > 1 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && (uint64_t)sb.st_size >= 3 * chunk_;
    if (par) {
      size_ = (uint64_t)sb.st_size;
      void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) par = false;
      else { base_ = (const uint8_t*)m; madvise(m, size_, MADV_SEQUENTIAL); }
    }
    ::close(fd);
    const uint8_t* data = nullptr; const char* why = nullptr;
    if (par && !(base_[0] == 0x1f && base_[1] == 0x8b && parse_gzip_header(base_, base_ + size_, &data, &why) == 1)) { unmap(); par = false; }
    if (!par) { use_serial_ = true; return serial_.open(path); }
    use_serial_ = false;
    cbits_ = chunk_ * 8;
    n_chunks_ = (size_ + chunk_ - 1) / chunk_;
    spec_.clear(); spec_.resize(n_chunks_);
    spec_state_.assign(n_chunks_, 0);
    next_spec_ = 0; consumer_chunk_ = 0; stop_ = false; spec_on_ = true; misses_ = 0;
    segs_.clear(); resolve_q_.clear(); queued_bytes_ = 0;
    pos_ = first_data_bit_ = (uint64_t)(data - base_) * 8; hist_ = 0; win_.assign(kWin, 0);
    // A file that begins with a BGZF member is taken for BGZF: the workers then inflate runs of whole members (no block finder, no
    // markers: a member's size is in its header and it has no history), and the chain starts in front of the first header.
    uint64_t next0 = 0;
    bgzf_ = bgzf_header_at(base_, size_, 0, &next0);
    at_boundary_ = bgzf_; hdr_byte_ = 0;
    chain_done_ = false; fb_active_ = false;
    err_.clear(); failed_ = false; ended_ = false; cur_pos_ = 0;
    crc_run_ = 0; isize_run_ = 0; n_accepted_ = 0; n_fallbacks_ = 0;
    // a worker's decoder state and look-up table are allocated here, where a failure can still send the file to the serial reader
    try {
      kits_.clear();
      for (unsigned t = 0; t < threads_; ++t) { kits_.emplace_back(new Kit); kits_.back()->lut.assign(65536, 0); for (int i = 0; i < 256; ++i) kits_.back()->lut[i] = (uint8_t)i; }
      for (unsigned t = 0; t < threads_; ++t) { Kit* kit = kits_[t].get(); pool_.emplace_back([this, kit] { worker(*kit); }); }
    } catch (const std::exception&) { if (pool_.empty()) { close(); use_serial_ = true; return serial_.open(path); } }     // (fewer threads than asked for: carry on)
    return true;
  }

  void close()
  {
    if (!pool_.empty()) {
      { std::lock_guard<std::mutex> lk(mu_); stop_ = true; }
      cv_work_.notify_all();
      for (auto& t : pool_) t.join();
      pool_.clear();
    }
    kits_.clear();
    segs_.clear(); resolve_q_.clear(); spec_.clear(); spec_state_.clear(); sym_pool_.clear(); data_pool_.clear();
    unmap();
    serial_.close();
  }

  // up to cap bytes of inflated text; 0 at the end of the data, -1 on corrupt input (after the bytes in front of it were delivered)
  long read(uint8_t* dst, size_t cap)
  {
    if (use_serial_) return serial_.read(dst, cap);
    try { return read_chain(dst, cap); }
    catch (const std::exception&) { err_ = "out of memory"; failed_ = true; return -1; }       // (a buffer could not be had: nothing else throws in here)
  }

 private:
  long read_chain(uint8_t* dst, size_t cap)
  {
    size_t got = 0;
    while (got < cap) {
      if (failed_) return got ? (long)got : -1;
      if (ended_) break;
      pump();
      if (segs_.empty()) { if (chain_done_) ended_ = true; else wait_for_progress(); continue; }
      Seg& s = *segs_.front();
      if (!s.ready.load(std::memory_order_acquire)) { wait_for_progress(); continue; }
      if (s.kind == Seg::DATA) {
        if (cur_pos_ == 0 && s.n && !s.verified) { crc_run_ = (uint32_t)crc32_combine(crc_run_, s.crc, (z_off_t)s.n); isize_run_ += (uint32_t)s.n; }
        const size_t n = std::min(cap - got, s.n - cur_pos_);
        if (n) memcpy(dst + got, s.data.p + cur_pos_, n);
        got += n; cur_pos_ += n;
        if (cur_pos_ == s.n) pop_front();
        continue;
      }
      if (s.kind == Seg::MEMBER_END) {
        // the trailer, as GunzipStream reads it: fewer than 4 bytes = an early end; the CRC as soon as its 4 bytes are there
        if (s.trailer_bytes < 4) { ended_ = true; break; }
        if (verify_ && s.crc != crc_run_) { err_ = "incorrect data check"; failed_ = true; continue; }
        if (s.trailer_bytes < 8) { ended_ = true; break; }
        if (s.isize != isize_run_) { err_ = "incorrect length check"; failed_ = true; continue; }
        crc_run_ = 0; isize_run_ = 0;
        pop_front();
        continue;
      }
      if (s.kind == Seg::ERROR) { err_ = s.msg; failed_ = true; continue; }
      ended_ = true;                                        // Seg::END
    }
    return (long)got;
  }

 private:
  struct Seg {
    enum Kind { DATA, MEMBER_END, ERROR, END } kind = DATA;
    RawBuf<uint8_t> data; size_t n = 0; uint32_t crc = 0; std::atomic<bool> ready{false};
    std::unique_ptr<SpecChunk> src; std::vector<uint8_t> win;       // DATA still to be resolved by a worker: symbols + the window in front of them
    uint32_t isize = 0; int trailer_bytes = 0;                      // MEMBER_END (crc = the trailer's)
    std::string msg;                                                // ERROR
    bool verified = false;                                          // DATA of a BGZF run: whole members, their CRC-32 / ISIZE already checked
  };

  void unmap() { if (base_) munmap((void*)base_, size_); base_ = nullptr; }
  void pop_front()
  {
    {
      std::lock_guard<std::mutex> lk(mu_);
      queued_bytes_ -= std::min<uint64_t>(queued_bytes_, segs_.front()->n);
      if (segs_.front()->data.p && data_pool_.size() < pool_cap()) data_pool_.push_back(std::move(segs_.front()->data));
    }
    segs_.pop_front(); cur_pos_ = 0;
    cv_work_.notify_all();
  }
  // nothing to deliver and the chain cannot move: until a worker finishes something (a resolve or the chunk the chain waits for)
  void wait_for_progress()
  {
    std::unique_lock<std::mutex> lk(mu_);
    const uint64_t j = chain_chunk();
    cv_done_.wait(lk, [&] {
      if (!segs_.empty() && segs_.front()->ready.load(std::memory_order_acquire)) return true;
      return !chain_done_ && !fb_active_ && j < n_chunks_ && spec_state_[j] == 2 && queued_bytes_ < kMaxQueued;
    });
  }

  // ---- workers: resolve tasks first, then the next chunk to decode ahead of the chain ----
  struct Kit { Inflater inf; std::vector<uint8_t> lut; };   // one per worker: decoder tables (150 KB: not on a thread's stack) and the resolve table
  void worker(Kit& kit)
  {
    Inflater* const T = &kit.inf;
    std::vector<uint8_t>& lut = kit.lut;
    for (;;) {
      Seg* job = nullptr; uint64_t k = 0;
      RawBuf<uint16_t> spare_sym;                             // buffers go round (fresh ones cost a page fault per 4 KiB)
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return stop_ || !resolve_q_.empty() || can_spec(); });
        if (stop_) return;
        if (!resolve_q_.empty()) {
          job = resolve_q_.front(); resolve_q_.pop_front();
          if (!data_pool_.empty()) { job->data = std::move(data_pool_.back()); data_pool_.pop_back(); }
        } else {
          k = next_spec_++; spec_state_[k] = 1;
          if (!sym_pool_.empty()) { spare_sym = std::move(sym_pool_.back()); sym_pool_.pop_back(); }
        }
      }
      if (job) {
        bool ok = true;
        try {
          memcpy(lut.data() + kMarker, job->win.data(), kWin);
          const uint16_t* s = job->src->sym.p + kWin;
          const size_t n = job->src->n;
          job->data.reserve(n);
          uint8_t* o = job->data.p;
          size_t i = 0;
          for (; i + 8 <= n; i += 8) {
            o[i] = lut[s[i]]; o[i + 1] = lut[s[i + 1]]; o[i + 2] = lut[s[i + 2]]; o[i + 3] = lut[s[i + 3]];
            o[i + 4] = lut[s[i + 4]]; o[i + 5] = lut[s[i + 5]]; o[i + 6] = lut[s[i + 6]]; o[i + 7] = lut[s[i + 7]];
          }
          for (; i < n; ++i) o[i] = lut[s[i]];
          job->crc = verify_ ? crc32_update(0u, o, n) : 0u;
        } catch (const std::exception&) { ok = false; }
        if (!ok) { job->kind = Seg::ERROR; job->msg = "out of memory"; job->n = 0; }
        RawBuf<uint16_t> used = std::move(job->src->sym);
        job->src.reset(); job->win = std::vector<uint8_t>();
        {
          std::lock_guard<std::mutex> lk(mu_);
          if (used.p && sym_pool_.size() < pool_cap()) sym_pool_.push_back(std::move(used));
          job->ready.store(true, std::memory_order_release);
        }
        cv_done_.notify_all();
        continue;
      }
      std::unique_ptr<SpecChunk> R;
      try {
        R.reset(new SpecChunk);
        R->sym = std::move(spare_sym);
        const uint64_t from = k * cbits_, until = std::min((k + 1) * cbits_, size_ * 8);
        uint64_t p = from; bool found = false;
        if (bgzf_) {                                                  // the first member header in the chunk, then whole members
          uint64_t next = 0;
          for (p = from >> 3; p < (until >> 3); ++p) if (base_[p] == 0x1f && bgzf_header_at(base_, size_, p, &next)) { found = true; break; }
          if (found) bgzf_run_decode(*T, base_, size_, p, until >> 3, verify_, *R);
          if (!R->clean) { R->bytes.release(); R->n = 0; }
        } else {
          if (k == 0) { p = first_data_bit_; found = p < until; }      // the first member's data: a certain start, whatever its block type
          else for (; p < until; ++p) if (dynamic_header_at(*T, base_, size_, p)) { found = true; break; }
          if (found) spec_decode(*T, base_, size_, p, until, kMaxSyms, *R);
          if (!R->clean) { R->sym.release(); R->n = 0; }
        }
      } catch (const std::exception&) { if (R) { R->clean = false; R->sym.release(); } }
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (k < consumer_chunk_) recycle(R);                  // the chain went past this chunk while it was being decoded
        spec_[k] = std::move(R); spec_state_[k] = 2;
      }
      cv_done_.notify_all();
    }
  }
  void recycle(std::unique_ptr<SpecChunk>& R)               // mu_ held: the symbol buffer back to the pool, the chunk gone
  {
    if (R && R->sym.p && sym_pool_.size() < pool_cap()) sym_pool_.push_back(std::move(R->sym));
    if (R && R->bytes.p && data_pool_.size() < pool_cap()) data_pool_.push_back(std::move(R->bytes));
    R.reset();
  }
  size_t pool_cap() const { return 2 * (size_t)threads_ + 4; }
  bool can_spec() const { return spec_on_ && next_spec_ < n_chunks_ && next_spec_ < consumer_chunk_ + 2 * (uint64_t)threads_ + 2 && queued_bytes_ < kMaxQueued; }

  // ---- the chain (caller's thread) ----
  void push(std::unique_ptr<Seg> s) { { std::lock_guard<std::mutex> lk(mu_); queued_bytes_ += s->n; } segs_.push_back(std::move(s)); }
  void push_last(Seg::Kind k, const char* msg = nullptr)
  {
    std::unique_ptr<Seg> s(new Seg);
    s->kind = k; s->ready.store(true); if (msg) s->msg = msg;
    push(std::move(s));
    chain_done_ = true;
  }

  // As far as the chain goes without waiting for a worker: accept the chunk that starts where the known text ends, or decode
  // serially towards the next one that might fit.  (Serial stretches are work for this thread whenever they are done: done now.)
  void pump()
  {
    for (;;) {
      if (chain_done_) return;
      if (queued_bytes_ >= kMaxQueued) return;              // the reader is behind: nothing more until it has caught up
      if (fb_active_) { fallback_step(); continue; }
      if (at_boundary_) {                                   // in front of a member header (or of the end of the file)
        if (hdr_byte_ >= size_) { push_last(Seg::END); return; }
        if (bgzf_ && spec_on_) {
          const uint64_t jb = chain_chunk();
          std::unique_ptr<SpecChunk> run;
          {
            std::lock_guard<std::mutex> lk(mu_);
            move_consumer_to(jb);
            if (spec_state_[jb] != 2) return;               // not there yet
            if (spec_[jb] && spec_[jb]->clean && spec_[jb]->bgzf_run && spec_[jb]->start_byte == hdr_byte_) run = std::move(spec_[jb]);
            else if (!(spec_[jb] && spec_[jb]->clean && spec_[jb]->bgzf_run && spec_[jb]->start_byte > hdr_byte_)) recycle(spec_[jb]);
          }
          if (run) { accept_run(std::move(run)); misses_ = 0; continue; }
          if (++misses_ >= kMaxMisses) { std::lock_guard<std::mutex> lk(mu_); spec_on_ = false; }
        }
        enter_member();                                     // this member through the serial decoder
        continue;
      }
      const uint64_t j = pos_ / cbits_;
      if (j >= n_chunks_ || !spec_on_ || bgzf_) { start_fallback(~0ull); continue; }     // (BGZF: to the member's end; the next run may fit again)
      std::unique_ptr<SpecChunk> R;
      uint64_t target = (j + 1) * cbits_;
      {
        std::lock_guard<std::mutex> lk(mu_);
        move_consumer_to(j);
        if (spec_state_[j] != 2) return;                    // not there yet
        if (spec_[j] && spec_[j]->clean && spec_[j]->start_bit >= pos_) {
          if (spec_[j]->start_bit == pos_) R = std::move(spec_[j]); else target = spec_[j]->start_bit;
        } else recycle(spec_[j]);                           // nothing found, or it started behind the known text: of no use
      }
      if (R && hist_ < kWin && reaches_before_start(*R)) R.reset();      // a match beyond the start of the member: the serial decoder reports it
      if (R) { accept(std::move(R)); misses_ = 0; continue; }
      if (++misses_ >= kMaxMisses) { std::lock_guard<std::mutex> lk(mu_); spec_on_ = false; }     // not a file this scheme suits (BGZF, fixed codes, ...)
      start_fallback(target);
    }
  }

  uint64_t chain_chunk() const { return std::min((at_boundary_ ? hdr_byte_ * 8 : pos_) / cbits_, n_chunks_ ? n_chunks_ - 1 : 0); }
  void move_consumer_to(uint64_t j)                         // mu_ held
  {
    if (consumer_chunk_ == j) return;
    for (uint64_t k = consumer_chunk_; k < j; ++k) recycle(spec_[k]);      // chunks the chain has passed (a block may span several)
    consumer_chunk_ = j; if (next_spec_ < j) next_spec_ = j;
    cv_work_.notify_all();
  }

  void accept_run(std::unique_ptr<SpecChunk> R)             // whole BGZF members, checked by the worker: their text as it is
  {
    ++n_accepted_;
    std::unique_ptr<Seg> s(new Seg);
    s->kind = Seg::DATA; s->n = R->n; s->data = std::move(R->bytes); s->verified = true; s->ready.store(true);
    hdr_byte_ = R->end_byte; hist_ = 0;
    push(std::move(s));
  }

  // the member header at hdr_byte_: its deflate data is where the chain goes on (or the file ends here, as GunzipStream sees it)
  void enter_member()
  {
    const uint8_t* p = base_ + hdr_byte_; const uint8_t* e = base_ + size_;
    if (e - p < 2 || p[0] != 0x1f || p[1] != 0x8b) { push_last(Seg::END); return; }      // nothing, or bytes that are not a member: ignored
    const uint8_t* data = nullptr; const char* why = nullptr;
    const int r = parse_gzip_header(p, e, &data, &why);
    if (r == 0) { push_last(Seg::END); return; }            // truncated inside a header: an early end
    if (r < 0) { push_last(Seg::ERROR, why); return; }
    pos_ = (uint64_t)(data - base_) * 8; hist_ = 0; at_boundary_ = false;
  }

  bool reaches_before_start(const SpecChunk& R) const
  {
    const uint32_t lowest = (uint32_t)kMarker + (kWin - (uint32_t)hist_);         // markers below this point in front of the member (hist_ = 0: all of them)
    const uint16_t* s = R.sym.p + kWin;
    for (size_t i = 0; i < R.n; ++i) if (s[i] >= kMarker && (uint32_t)s[i] < lowest) return true;
    return false;
  }

  void accept(std::unique_ptr<SpecChunk> R)
  {
    ++n_accepted_;
    std::unique_ptr<Seg> s(new Seg);
    s->kind = Seg::DATA; s->n = R->n; s->win = win_;
    // the next chunk's window: the last 32 KiB of this chunk's text (or what is left of the old window, then all of it)
    const uint16_t* sy = R->sym.p + kWin;
    const size_t n = R->n;
    auto byte_of = [&](uint16_t v) -> uint8_t { return v & kMarker ? win_[v & (kMarker - 1)] : (uint8_t)v; };
    if (n >= kWin) { for (uint32_t i = 0; i < kWin; ++i) scratch_[i] = byte_of(sy[n - kWin + i]); memcpy(win_.data(), scratch_, kWin); }
    else if (n) {
      for (size_t i = 0; i < n; ++i) scratch_[i] = byte_of(sy[i]);
      memmove(win_.data(), win_.data() + n, kWin - n); memcpy(win_.data() + kWin - n, scratch_, n);
    }
    hist_ = std::min<uint64_t>(kWin, hist_ + n);
    pos_ = R->end_bit;
    const bool fin = R->final_block;
    Seg* raw = s.get();
    if (n) s->src = std::move(R); else s->ready.store(true);
    push(std::move(s));
    if (n) { { std::lock_guard<std::mutex> lk(mu_); resolve_q_.push_back(raw); } cv_work_.notify_all(); }
    if (fin) member_end();
  }

  // after a member's final block (pos_ = the bit behind it): its trailer, then the next member's header or the end
  void member_end()
  {
    const uint64_t tb = std::min(size_, (pos_ + 7) >> 3);
    std::unique_ptr<Seg> s(new Seg);
    s->kind = Seg::MEMBER_END; s->ready.store(true);
    const uint64_t left = size_ - tb;
    s->trailer_bytes = (int)std::min<uint64_t>(left, 8);
    const uint8_t* t = base_ + tb;
    if (left >= 4) s->crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    if (left >= 8) s->isize = (uint32_t)t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
    push(std::move(s));
    if (left < 8) { chain_done_ = true; return; }
    at_boundary_ = true; hdr_byte_ = tb + 8;                // (pump goes on from there: a BGZF run, or enter_member())
  }

  // ---- the serial decoder on the caller's thread, one slab of output per step, until a block boundary >= target ----
  void start_fallback(uint64_t target_bit)
  {
    if ((pos_ >> 3) >= size_) { push_last(Seg::END); return; }        // the file ends where deflate data should begin: an early end
    ++n_fallbacks_;
    fb_in_ = fb_inf_.start_at_bit(base_, pos_);
    fb_inf_.pos_base = base_; fb_inf_.stop_bit = target_bit;
    if (fb_buf_.size() < kWin + kSlab + 320) fb_buf_.assign(kWin + kSlab + 320, 0);
    memcpy(fb_buf_.data(), win_.data(), kWin);
    fb_active_ = true;
  }
  void fallback_step()
  {
    uint8_t* const out0 = fb_buf_.data() + kWin; uint8_t* out = out0;
    const Inflater::Result r = fb_inf_.run(fb_in_, base_ + size_, true, out, out0 + kSlab, hist_);
    const size_t n = (size_t)(out - out0);
    if (n) {
      std::unique_ptr<Seg> s(new Seg);
      s->kind = Seg::DATA; s->n = n; s->data.reserve(n); memcpy(s->data.p, out0, n); s->ready.store(true);
      s->crc = verify_ ? crc32_update(0u, out0, n) : 0u;
      push(std::move(s));
      hist_ = std::min<uint64_t>(kWin, hist_ + n);
      memmove(fb_buf_.data(), fb_buf_.data() + n, kWin);    // the last 32 KiB of [window | slab] are the new window
    }
    if (r == Inflater::NEED_OUTPUT) return;
    memcpy(win_.data(), fb_buf_.data(), kWin);
    fb_active_ = false;
    if (r == Inflater::BOUNDARY) { pos_ = fb_inf_.boundary_bit; return; }
    if (r == Inflater::DONE) { pos_ = (uint64_t)(fb_in_ - base_) * 8 - (uint64_t)fb_inf_.bitcnt; member_end(); return; }
    if (r == Inflater::ERROR) { push_last(Seg::ERROR, fb_inf_.error ? fb_inf_.error : "invalid deflate data"); return; }
    push_last(Seg::END);                                    // NEED_INPUT with the whole file in hand: it is truncated -- an early end
  }

  static constexpr size_t kSlab = 4 << 20, kMaxSyms = 48u << 20;      // a chunk that inflates to more than 48 M symbols is left to the serial decoder
  static constexpr uint64_t kMaxQueued = 256ull << 20;                  // text waiting for the reader: the chain and the workers pause beyond it
  static constexpr unsigned kMaxMisses = 32;
  unsigned threads_ = 1;
  uint64_t chunk_ = 1 << 20, cbits_ = 8 << 20, n_chunks_ = 0;
  bool verify_ = true, use_serial_ = true;
  GunzipStream serial_;
  const uint8_t* base_ = nullptr; uint64_t size_ = 0;
  uint64_t first_data_bit_ = 0;
  // shared with the workers (mu_)
  std::mutex mu_; std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> pool_; std::vector<std::unique_ptr<Kit>> kits_;
  std::vector<std::unique_ptr<SpecChunk>> spec_; std::vector<uint8_t> spec_state_;      // 0 untouched, 1 being decoded, 2 done
  uint64_t next_spec_ = 0, consumer_chunk_ = 0, queued_bytes_ = 0;
  std::deque<Seg*> resolve_q_;
  std::vector<RawBuf<uint16_t>> sym_pool_; std::vector<RawBuf<uint8_t>> data_pool_;
  bool stop_ = false, spec_on_ = true;
  // caller's thread only
  std::deque<std::unique_ptr<Seg>> segs_;
  uint64_t pos_ = 0, hist_ = 0; std::vector<uint8_t> win_; uint8_t scratch_[kWin];
  bool chain_done_ = false, fb_active_ = false, failed_ = false, ended_ = false;
  bool bgzf_ = false, at_boundary_ = false; uint64_t hdr_byte_ = 0;      // BGZF file; the chain stands in front of the member header at hdr_byte_
  unsigned misses_ = 0;
  Inflater fb_inf_; const uint8_t* fb_in_ = nullptr; std::vector<uint8_t> fb_buf_;
  size_t cur_pos_ = 0; uint32_t crc_run_ = 0, isize_run_ = 0;
  uint64_t n_accepted_ = 0, n_fallbacks_ = 0;
  std::string err_;
};

}  // namespace hgz
