// hlm_pgzip.h — a plain gzip file (RFC 1952 / RFC 1951: ONE deflate stream, which is what `gzip`
// and most sequencing pipelines write) inflated on several host threads.  The reference reads
// r1=/r2= .gz files through boost::iostreams' gzip_decompressor (src/preselect.cpp:807-831,
// :987-1006): one zlib stream, one thread - 0.3 GB/s of text per file.
//
// A deflate stream cannot be entered in the middle by zlib: a block may start at any BIT, and
// its matches reach up to 32 KiB back into text the thread has not seen.  Both are solved the way
// pugz / rapidgzip do:
//   * the compressed bytes are cut into chunks; for every chunk but the first the start of a
//     dynamic-Huffman block is searched from the cut on, bit by bit (block header + both code
//     descriptions must be well formed: a handful of candidates per megabyte survive, and a
//     candidate is only kept if the chunk decodes from it without an error);
//   * a chunk is decoded into 16-bit symbols: a literal byte, or a MARKER "the byte at offset w
//     of the 32 KiB that precede this chunk"; when the previous chunk's text is known the markers
//     are replaced.
// Correctness does not rest on the search: chunk c is decoded until it reaches the start chunk
// c+1 assumed, and chunk c+1 is accepted only if chunk c ended on exactly that bit - otherwise it
// is decoded again from where chunk c really ended.  Since chunk 0 starts behind the gzip header,
// every accepted byte is what a sequential decoder produces.  CRC-32 and ISIZE of every member
// are checked (per-chunk CRCs merged with crc32_combine); concatenated members are followed.
// A truncated or damaged stream is an error, never a shorter output.
#ifndef HLM_PGZIP_H
#define HLM_PGZIP_H

#include <zlib.h>

#include "hlm_crc32.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace hlm_pgz {

constexpr int kWin = 32768;
constexpr uint16_t kMarker = 0x8000;  // 0x8000 + w: byte w of the 32 KiB before the chunk

struct BitIn {
  const uint8_t* base;
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf = 0;
  int cnt = 0;           // valid bits in buf
  bool overrun = false;  // bits were taken past the end of the data
  BitIn(const uint8_t* data, size_t n, uint64_t bit) : base(data), p(data + (bit >> 3)), end(data + n) {
    refill();
    take((int)(bit & 7));
  }
  inline void refill() {
    if (cnt < 0) {  // the fast decode loop takes bits without looking: past the end of the data
      overrun = true;
      cnt = 0;
      buf = 0;
    }
    if (end - p >= 8) {
      uint64_t v;
      memcpy(&v, p, 8);
      buf |= v << cnt;
      int adv = (63 - cnt) >> 3;
      p += adv;
      cnt += adv * 8;
    } else {
      while (cnt <= 56 && p < end) {
        buf |= (uint64_t)*p++ << cnt;
        cnt += 8;
      }
    }
  }
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void take(int n) {
    buf >>= n;
    cnt -= n;
    if (cnt < 0) {
      overrun = true;
      cnt = 0;
      buf = 0;
    }
  }
  inline uint32_t get(int n) {
    if (cnt < n) refill();
    uint32_t v = peek(n);
    take(n);
    return v;
  }
  uint64_t bitpos() const { return (uint64_t)(p - base) * 8 - (uint64_t)cnt; }
  void align() { take(cnt & 7); }
};

// canonical Huffman code: a direct table for the first PB bits, the classic count / sorted-symbol
// walk for the (rare) longer codes
template <int PB, int MAXSYM>
struct Huff {
  uint16_t tab[1 << PB];  // (symbol << 4) | length, 0: longer than PB bits
  uint16_t cnt[16];
  uint16_t sorted[MAXSYM];
  int n_used = 0, maxlen = 0;
  // 0: complete code; 1: incomplete (or empty); -1: over-subscribed
  int build(const uint8_t* len, int n) {
    memset(cnt, 0, sizeof cnt);
    for (int i = 0; i < n; i++) cnt[len[i]]++;
    n_used = n - cnt[0];
    cnt[0] = 0;
    int left = 1;
    maxlen = 0;
    for (int l = 1; l <= 15; l++) {
      left <<= 1;
      left -= cnt[l];
      if (left < 0) return -1;
      if (cnt[l]) maxlen = l;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
    for (int i = 0; i < n; i++)
      if (len[i]) sorted[offs[len[i]]++] = (uint16_t)i;
    memset(tab, 0, sizeof tab);
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= 15; l++) {
      for (int k = 0; k < cnt[l]; k++, code++, idx++) {
        if (l > PB) continue;
        uint32_t rev = 0;
        for (int b = 0; b < l; b++) rev |= ((code >> b) & 1u) << (l - 1 - b);
        uint16_t e = (uint16_t)((sorted[idx] << 4) | l);
        for (uint32_t x = rev; x < (1u << PB); x += 1u << l) tab[x] = e;
      }
      code <<= 1;
    }
    return left > 0 ? 1 : 0;
  }
  // symbol, or -1 (no such code / input exhausted)
  inline int decode(BitIn& in) const {
    uint16_t e = tab[in.peek(PB)];
    if (e & 15) {
      in.take(e & 15);
      return e >> 4;
    }
    int code = 0, first = 0, index = 0;
    uint64_t b = in.buf;
    for (int l = 1; l <= maxlen; l++) {
      code |= (int)(b & 1);
      b >>= 1;
      int c = cnt[l];
      if (code - c < first) {
        in.take(l);
        return sorted[index + (code - first)];
      }
      index += c;
      first += c;
      first <<= 1;
      code <<= 1;
    }
    return -1;
  }
};

constexpr int kLitBits = 11, kDistBits = 8;
constexpr uint32_t kFLit = 1u << 4, kFEob = 1u << 5;
typedef Huff<kLitBits, 288> LitHuff;
typedef Huff<kDistBits, 32> DistHuff;
typedef Huff<7, 19> PreHuff;

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t kPreOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// the header of a dynamic block (the three bits BFINAL / BTYPE already taken): both codes
inline bool read_dynamic(BitIn& in, LitHuff& lit, DistHuff& dist, bool strict) {
  in.refill();
  int hlit = (int)in.get(5) + 257, hdist = (int)in.get(5) + 1, hclen = (int)in.get(4) + 4;
  if (hlit > 286 || hdist > 30) return false;
  uint8_t pl[19];
  memset(pl, 0, sizeof pl);
  for (int i = 0; i < hclen; i++) pl[kPreOrder[i]] = (uint8_t)in.get(3);
  PreHuff pre;
  int r = pre.build(pl, 19);
  if (r < 0 || (r > 0 && (strict || pre.n_used != 1))) return false;
  uint8_t len[286 + 30];
  int n = 0, total = hlit + hdist;
  while (n < total) {
    in.refill();
    int s = pre.decode(in);
    if (s < 0 || in.overrun) return false;
    if (s < 16) {
      len[n++] = (uint8_t)s;
      continue;
    }
    int rep, v = 0;
    if (s == 16) {
      if (n == 0) return false;
      v = len[n - 1];
      rep = 3 + (int)in.get(2);
    } else if (s == 17) {
      rep = 3 + (int)in.get(3);
    } else {
      rep = 11 + (int)in.get(7);
    }
    if (n + rep > total) return false;
    while (rep--) len[n++] = (uint8_t)v;
  }
  if (len[256] == 0) return false;  // no end-of-block code
  r = lit.build(len, hlit);
  if (r < 0 || (r > 0 && (strict || lit.n_used != 1))) return false;
  r = dist.build(len + hlit, hdist);
  // zlib accepts an incomplete distance code only when it has a single code (or none)
  if (r < 0 || (r > 0 && dist.n_used > 1)) return false;
  return !in.overrun;
}

inline void fixed_codes(LitHuff& lit, DistHuff& dist) {
  uint8_t l[288], d[30];
  for (int i = 0; i < 144; i++) l[i] = 8;
  for (int i = 144; i < 256; i++) l[i] = 9;
  for (int i = 256; i < 280; i++) l[i] = 7;
  for (int i = 280; i < 288; i++) l[i] = 8;
  for (int i = 0; i < 30; i++) d[i] = 5;
  lit.build(l, 288);
  dist.build(d, 30);  // 30 of 32 codes: incomplete by definition, codes 30 / 31 never decode
}

// The symbols of one Huffman-coded block (its code descriptions already read into lit / dist) up to
// its end-of-block code.  T = uint16_t with MARK: matches that reach before the output become
// window markers (a chunk entered in the middle of a stream); T = uint8_t without: such a match
// is an error.  `room(pos)` is called when pos passes `limit`: it makes space (and may move the
// buffer, returned through `o`) or fails; after it, 1024 more symbols must fit.
// Returns false on invalid data or when the input ran out.
template <class T, bool MARK, class Room>
inline bool decode_block(BitIn& in, const LitHuff& lit, const DistHuff& dist, T*& o, size_t& pos, size_t& limit,
                         size_t member_base, bool member_known, Room room) {
  uint32_t flit[1 << kLitBits], fdist[1 << kDistBits];
  // direct tables: what a code means, ready to use (literal byte / length base and number of extra
  // bits / end of block), so that a literal costs one look-up and up to three literals are taken
  // per refill of the bit buffer
  for (uint32_t i = 0; i < (1u << kLitBits); i++) {
    uint16_t e = lit.tab[i];
    uint32_t l = e & 15u, sy = e >> 4, f = 0;
    if (l) {
      if (sy < 256) f = kFLit | (sy << 16) | l;
      else if (sy == 256) f = kFEob | l;
      else if (sy - 257 < 29) f = ((uint32_t)kLenBase[sy - 257] << 16) | ((uint32_t)kLenExtra[sy - 257] << 8) | l;
    }
    flit[i] = f;  // 0: a longer (or invalid) code - the generic decoder takes over
  }
  for (uint32_t i = 0; i < (1u << kDistBits); i++) {
    uint16_t e = dist.tab[i];
    uint32_t l = e & 15u, sy = e >> 4;
    fdist[i] = (l && sy < 30) ? (((uint32_t)kDistBase[sy] << 16) | ((uint32_t)kDistExtra[sy] << 8) | l) : 0;
  }
  for (;;) {
    if (pos >= limit) {
      // also the place where a decode that ran off the data, or a candidate start that is not one
      // and expands without bound, is stopped
      if (in.cnt < 0 || in.overrun) return false;
      if (!room(pos)) return false;
    }
    in.refill();
    uint32_t f = flit[in.buf & ((1u << kLitBits) - 1)];
    if (f & kFLit) {
      o[pos++] = (T)(f >> 16);
      in.buf >>= (f & 15);
      in.cnt -= (int)(f & 15);
      f = flit[in.buf & ((1u << kLitBits) - 1)];
      if (f & kFLit) {
        o[pos++] = (T)(f >> 16);
        in.buf >>= (f & 15);
        in.cnt -= (int)(f & 15);
        f = flit[in.buf & ((1u << kLitBits) - 1)];
        if (f & kFLit) {
          o[pos++] = (T)(f >> 16);
          in.buf >>= (f & 15);
          in.cnt -= (int)(f & 15);
          continue;
        }
      }
      in.refill();  // a match needs up to 48 bits
    }
    int len;
    if (f == 0) {  // a code longer than the table's index, or an invalid one
      int sl = lit.decode(in);
      if (sl < 256) {
        if (sl < 0) return false;
        o[pos++] = (T)sl;
        continue;
      }
      if (sl == 256) break;
      sl -= 257;
      if (sl >= 29) return false;
      len = kLenBase[sl] + (int)in.get(kLenExtra[sl]);
    } else {
      in.buf >>= (f & 15);
      in.cnt -= (int)(f & 15);
      if (f & kFEob) break;
      uint32_t xb = (f >> 8) & 31u;
      len = (int)(f >> 16) + (int)(in.buf & ((1u << xb) - 1));
      in.buf >>= xb;
      in.cnt -= (int)xb;
    }
    size_t dd;
    uint32_t fd = fdist[in.buf & ((1u << kDistBits) - 1)];
    if (fd) {
      in.buf >>= (fd & 15);
      in.cnt -= (int)(fd & 15);
      uint32_t xb = (fd >> 8) & 31u;
      dd = (size_t)(fd >> 16) + (size_t)(in.buf & ((1u << xb) - 1));
      in.buf >>= xb;
      in.cnt -= (int)xb;
    } else {
      int ds = dist.decode(in);
      if (ds < 0 || ds >= 30) return false;
      dd = kDistBase[ds] + (size_t)in.get(kDistExtra[ds]);
    }
    if (in.cnt < 0 || in.overrun) return false;
    size_t avail = pos - member_base;
    if (dd > avail) {
      // reaches before the text produced in the current member
      // (member_base == 0 here: an unknown member start can only be the chunk's own start)
      if (!MARK || member_known || dd > avail + kWin) return false;
      for (int k = 0; k < len; k++, pos++)
        o[pos] = pos >= dd ? o[pos - dd] : (T)(kMarker + (kWin - (dd - pos)));
    } else {
      const T* src = o + pos - dd;
      T* dst = o + pos;
      if (dd * sizeof(T) >= 16) {
        // sixteen bytes per step; the last step may write past the match, inside the slack every
        // buffer keeps behind `limit`
        for (size_t k = 0; k < (size_t)len * sizeof(T); k += 16) memcpy((char*)dst + k, (const char*)src + k, 16);
      } else {
        for (int k = 0; k < len; k++) dst[k] = src[k];
      }
      pos += (size_t)len;
    }
  }
  return in.cnt >= 0 && !in.overrun;
}

// One raw deflate stream (a BGZF block) of known size into bytes: the stream must end with its
// final block exactly at out_len bytes.  `scratch` holds at least out_len + 2048 bytes.
inline bool inflate_raw(const uint8_t* d, size_t n, uint8_t* scratch, size_t out_len) {
  BitIn in(d, n, 0);
  LitHuff lit;
  DistHuff dist;
  uint8_t* o = scratch;
  size_t pos = 0, limit = out_len + 1;  // room() is only ever asked when the stream is too long
  for (;;) {
    in.refill();
    uint32_t bfinal = in.get(1), btype = in.get(2);
    if (in.overrun || btype == 3) return false;
    if (btype == 0) {
      in.align();
      in.refill();
      uint32_t len = in.get(16), nlen = in.get(16);
      if (in.overrun || (len ^ 0xFFFFu) != nlen || pos + len > out_len) return false;
      while (len && in.cnt >= 8) {
        o[pos++] = (uint8_t)in.get(8);
        len--;
      }
      if (len) {
        if ((size_t)(in.end - in.p) < len) return false;
        in.buf = 0;
        in.cnt = 0;
        memcpy(o + pos, in.p, len);
        pos += len;
        in.p += len;
      }
    } else {
      if (btype == 1) fixed_codes(lit, dist);
      else if (!read_dynamic(in, lit, dist, false)) return false;
      if (!decode_block<uint8_t, false>(in, lit, dist, o, pos, limit, 0, true, [](size_t) { return false; }))
        return false;
    }
    if (pos > out_len) return false;
    if (bfinal) return pos == out_len;
  }
}

// growable array of 16-bit symbols that is never zero-filled
struct SymBuf {
  uint16_t* p = nullptr;
  size_t cap = 0, n = 0;
  SymBuf() {}
  SymBuf(const SymBuf&) = delete;
  SymBuf& operator=(const SymBuf&) = delete;
  ~SymBuf() { free(p); }
  bool grow(size_t want) {
    size_t c = cap ? cap : ((size_t)1 << 20);
    while (c < want) c *= 2;
    uint16_t* q = (uint16_t*)realloc(p, c * sizeof(uint16_t));
    if (!q) return false;
    p = q;
    cap = c;
    return true;
  }
  size_t size() const { return n; }
  void clear() { n = 0; }
  uint16_t operator[](size_t i) const { return p[i]; }
  const uint16_t* data() const { return p; }
};

// where a gzip member ended inside a chunk's output
struct MemberEnd {
  size_t off;  // output symbols of the chunk that belong to members ending here or earlier
  uint32_t crc, isize;
};

struct Chunk {
  uint64_t start_bit = 0, end_bit = 0;  // decoded [start, end): end = a block boundary (or the end of the data)
  bool has_start = false;               // a block start was found (or is known) for this chunk
  bool known_start = false;             // start_bit is exact (first chunk of a round)
  bool member_start = false;            // start_bit is the first block of a member: no markers allowed
  bool ok = false;                      // decoded without an error up to end_bit
  bool hit_end = false;                 // the data ended here (after a final block and its trailer)
  SymBuf sym;                           // literals and markers
  std::vector<MemberEnd> ends;
  std::vector<uint32_t> seg_crc;        // CRC-32 of the output segments between member ends
  void reset() {
    has_start = known_start = member_start = ok = hit_end = false;
    sym.clear();
    ends.clear();
    seg_crc.clear();
  }
};

// RFC 1952 member header at byte `at`; returns the byte offset of the deflate data or 0
inline size_t gzip_header(const uint8_t* d, size_t n, size_t at) {
  if (n - at < 18 || d[at] != 0x1f || d[at + 1] != 0x8b || d[at + 2] != 8) return 0;
  int flg = d[at + 3];
  size_t p = at + 10;
  if (flg & 4) {
    if (p + 2 > n) return 0;
    p += 2 + (size_t)(d[p] | (d[p + 1] << 8));
  }
  if (flg & 8) {
    while (p < n && d[p]) p++;
    p++;
  }
  if (flg & 16) {
    while (p < n && d[p]) p++;
    p++;
  }
  if (flg & 2) p += 2;
  return p < n ? p : 0;
}

// Decodes blocks from c.start_bit until the position reaches `stop_bit` (checked at block
// boundaries) or the data ends.  c.member_start: text before the start does not exist (markers are
// an error); otherwise matches that reach before the chunk become markers.  `max_sym`: give up
// (return false) when more output than that is produced - a candidate start that is not one can
// expand without bound.
inline bool decode_chunk(const uint8_t* d, size_t n, Chunk& c, uint64_t stop_bit, size_t max_sym) {
  BitIn in(d, n, c.start_bit);
  LitHuff lit;
  DistHuff dist;
  SymBuf& sb = c.sym;
  sb.clear();
  size_t pos = 0;           // symbols produced
  size_t member_base = 0;   // output offset at which the current member started (if inside the chunk)
  bool member_known = c.member_start;
  c.ok = false;
  c.hit_end = false;
  c.ends.clear();
  if (!sb.cap && !sb.grow((size_t)1 << 20)) return false;
  uint16_t* o = sb.p;
  for (;;) {
    if (in.bitpos() >= stop_bit) break;
    in.refill();
    uint32_t bfinal = in.get(1), btype = in.get(2);
    if (in.overrun || btype == 3) return false;
    if (btype == 0) {
      in.align();
      in.refill();
      uint32_t len = in.get(16), nlen = in.get(16);
      if (in.overrun || (len ^ 0xFFFFu) != nlen) return false;
      if (pos + len + 1024 > sb.cap) {
        if (!sb.grow(pos + len + 1024)) return false;
        o = sb.p;
      }
      // the bit buffer holds whole bytes now: drain it, then copy straight from the data
      while (len && in.cnt >= 8) {
        o[pos++] = (uint16_t)in.get(8);
        len--;
      }
      if (len) {
        if ((size_t)(in.end - in.p) < len) return false;
        in.buf = 0;
        in.cnt = 0;
        for (uint32_t i = 0; i < len; i++) o[pos + i] = in.p[i];
        pos += len;
        in.p += len;
      }
    } else {
      if (btype == 1) fixed_codes(lit, dist);
      else if (!read_dynamic(in, lit, dist, false)) return false;
      size_t limit = sb.cap - 1024;
      auto room = [&](size_t at) {
        if (at > max_sym) return false;
        if (!sb.grow(sb.cap * 2)) return false;
        o = sb.p;
        limit = sb.cap - 1024;
        return true;
      };
      if (!decode_block<uint16_t, true>(in, lit, dist, o, pos, limit, member_base, member_known, room)) return false;
    }
    if (pos > max_sym) return false;
    if (bfinal) {
      // member trailer, then either the end of the data or the next member
      in.align();
      uint64_t byte = in.bitpos() >> 3;
      if (byte + 8 > n) return false;
      MemberEnd me;
      me.off = pos;
      memcpy(&me.crc, d + byte, 4);
      memcpy(&me.isize, d + byte + 4, 4);
      c.ends.push_back(me);
      size_t next = (size_t)byte + 8;
      while (next < n && d[next] == 0) next++;  // zero padding after a member (gzip accepts it)
      if (next >= n) {
        c.hit_end = true;
        break;
      }
      size_t data = gzip_header(d, n, next);
      if (!data) return false;
      in = BitIn(d, n, (uint64_t)data * 8);
      member_base = pos;
      member_known = true;
    }
  }
  c.end_bit = c.hit_end ? (uint64_t)n * 8 : in.bitpos();
  sb.n = pos;
  c.ok = true;
  return true;
}

// first bit >= from (and < to) at which a non-final dynamic block with well-formed code
// descriptions starts; `to` when there is none
inline uint64_t find_block(const uint8_t* d, size_t n, uint64_t from, uint64_t to) {
  LitHuff lit;
  DistHuff dist;
  for (uint64_t bit = from; bit < to; bit++) {
    size_t byte = (size_t)(bit >> 3);
    if (byte + 8 > n) break;
    uint64_t v;
    memcpy(&v, d + byte, 8);
    v >>= (bit & 7);
    // BFINAL = 0, BTYPE = 2 (bits 1, 0 in stream order), HLIT <= 29, HDIST <= 29
    if ((v & 7) != 4) continue;
    if (((v >> 3) & 31) > 29 || ((v >> 8) & 31) > 29) continue;
    BitIn in(d, n, bit + 3);
    if (!read_dynamic(in, lit, dist, true)) continue;
    return bit;
  }
  return to;
}

}  // namespace hlm_pgz

class HlmPgzip {
 public:
  HlmPgzip(const uint8_t* data, size_t size, int threads)
      : d_(data), n_(size), threads_(threads < 1 ? 1 : threads) {
    size_t at = hlm_pgz::gzip_header(d_, n_, 0);
    if (!at) {
      err = "not a gzip file";
      failed_ = finished_ = true;
      return;
    }
    next_bit_ = (uint64_t)at * 8;
    next_member_start_ = true;
    const char* e = getenv("HLM_PGZIP_CHUNK_KB");  // the tests shrink the chunks to their inputs
    long kb = e ? atol(e) : 0;
    chunk_bytes_ = kb > 0 ? (size_t)kb << 10 : (size_t)2 << 20;
    for (int i = 0; i < threads_; i++) chunks_.emplace_back(new hlm_pgz::Chunk());
    // rounds are decoded ahead of the consumer by one producer thread (which fans every round out
    // over `threads` workers): the caller parses round k while round k + 1 is inflated
    producer_ = std::thread([this] { produce(); });
  }
  ~HlmPgzip() {
    {
      std::lock_guard<std::mutex> g(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (producer_.joinable()) producer_.join();
    for (Text& t : free_texts_) free(t.p);
    for (auto& r : queue_)
      for (Text& t : r) free(t.p);
    for (Text& t : cur_) free(t.p);
  }
  HlmPgzip(const HlmPgzip&) = delete;
  HlmPgzip& operator=(const HlmPgzip&) = delete;
  std::string err;
  int64_t n_slow = 0, n_chunks = 0;  // chunks decoded again from a corrected start / all chunks
  // a block of text; its memory is recycled between the producer and fill() and never zeroed
  struct Text {
    uint8_t* p = nullptr;
    size_t n = 0, cap = 0;
  };
  // Up to `cap` bytes of text to dst; 0 at the end of the file, -1 with `err` set.
  long fill(uint8_t* dst, size_t cap) {
    size_t got = 0;
    while (got < cap) {
      if (cur_part_ >= cur_.size()) {  // next round
        std::unique_lock<std::mutex> g(mu_);
        for (Text& t : cur_) free_texts_.push_back(t);
        cur_.clear();
        cur_part_ = 0;
        cur_at_ = 0;
        cv_.wait(g, [this] { return !queue_.empty() || finished_; });
        if (queue_.empty()) {  // finished: the end of the file, or an error after the last good round
          if (failed_ && !got) return -1;
          break;
        }
        cur_.swap(queue_.front());
        queue_.pop_front();
        cv_.notify_all();
        continue;
      }
      Text& part = cur_[cur_part_];
      size_t k = std::min(cap - got, part.n - cur_at_);
      memcpy(dst + got, part.p + cur_at_, k);
      got += k;
      cur_at_ += k;
      if (cur_at_ >= part.n) {
        cur_part_++;
        cur_at_ = 0;
      }
    }
    return (long)got;
  }

 private:
  void produce() {
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [this] { return stop_ || queue_.size() < 2; });
        if (stop_) break;
      }
      std::vector<Text> parts;
      bool ok = false, last = false;
      try {
        ok = !failed_ && round(parts);
        last = done_;
      } catch (const std::exception&) {
        err = "out of memory while inflating";
      }
      std::lock_guard<std::mutex> g(mu_);
      if (ok) {
        queue_.push_back(std::move(parts));
      } else {
        for (Text& t : parts) free_texts_.push_back(t);
        failed_ = true;
      }
      if (!ok || last) {
        finished_ = true;
        cv_.notify_all();
        break;
      }
      cv_.notify_all();
    }
    std::lock_guard<std::mutex> g(mu_);
    finished_ = true;
    cv_.notify_all();
  }
  template <class F>
  void parallel(size_t n, F f) {
    std::atomic<size_t> next{0};
    auto work = [&] {
      for (size_t i; (i = next.fetch_add(1)) < n;) f(i);
    };
    size_t nt = std::min((size_t)threads_, n);
    std::vector<std::thread> th;
    for (size_t w = 1; w < nt; w++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  bool fail(const std::string& m) {
    err = m;
    return false;
  }
  // one round: the next `threads` chunks of compressed bytes
  static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  }
  // a recycled block of at least n bytes (the smallest that fits, else a new one)
  Text text_buffer(size_t n) {
    Text t;
    {
      std::lock_guard<std::mutex> g(mu_);
      size_t best = (size_t)-1;
      for (size_t i = 0; i < free_texts_.size(); i++)
        if (free_texts_[i].cap >= n && (best == (size_t)-1 || free_texts_[i].cap < free_texts_[best].cap)) best = i;
      if (best == (size_t)-1 && !free_texts_.empty()) best = 0;  // too small: replaced below
      if (best != (size_t)-1) {
        t = free_texts_[best];
        free_texts_.erase(free_texts_.begin() + (long)best);
      }
    }
    if (t.cap < n) {
      free(t.p);
      t.cap = n + (n >> 3) + 4096;
      t.p = (uint8_t*)malloc(t.cap);
      if (!t.p) throw std::bad_alloc();
    }
    t.n = n;
    return t;
  }
  bool round(std::vector<Text>& parts) {
    using namespace hlm_pgz;
    const bool trace = getenv("HLM_PGZIP_TRACE") != nullptr;
    double t0 = now(), t1 = t0, t2 = t0, t3 = t0, t4 = t0;
    const uint64_t total_bits = (uint64_t)n_ * 8;
    const size_t K = chunks_.size();
    // the first rounds use smaller chunks (1/8, 1/4, 1/2 of the size): the consumer gets its first
    // text after a fraction of a full round
    const int shrink = round_no_ < 3 ? 3 - round_no_ : 0;
    round_no_++;
    const uint64_t cb = (uint64_t)std::max<size_t>(chunk_bytes_ >> shrink, 1024) * 8;
    // nominal cuts: chunk i covers compressed bits [cut[i], cut[i + 1])
    std::vector<uint64_t> cut(K + 1);
    for (size_t i = 0; i <= K; i++) cut[i] = std::min(total_bits, next_bit_ + cb * i);
    for (auto& c : chunks_) c->reset();
    chunks_[0]->has_start = chunks_[0]->known_start = true;
    chunks_[0]->start_bit = next_bit_;
    chunks_[0]->member_start = next_member_start_;
    // 1. candidate starts
    parallel(K - 1, [&](size_t j) {
      size_t i = j + 1;
      if (cut[i] >= cut[i + 1]) return;
      uint64_t b = find_block(d_, n_, cut[i], cut[i + 1]);
      if (b < cut[i + 1]) {
        chunks_[i]->has_start = true;
        chunks_[i]->start_bit = b;
      }
    });
    t1 = now();
    // 2. every chunk decoded up to the start the next one assumed
    auto stop_of = [&](size_t i) {
      for (size_t k = i + 1; k < K; k++)
        if (chunks_[k]->has_start) return chunks_[k]->start_bit;
      return cut[K];
    };
    parallel(K, [&](size_t i) {
      Chunk& c = *chunks_[i];
      if (!c.has_start) return;
      uint64_t stop = stop_of(i);
      // a candidate start is given 64 bytes of text per compressed byte (FASTQ: 3 - 6); a chunk that
      // really expands further is decoded again below, from its verified start and without a limit
      size_t max_sym = c.known_start ? (size_t)-1 : (size_t)((stop - c.start_bit) / 8 + 1) * 64 + ((size_t)1 << 20);
      decode_chunk(d_, n_, c, stop, max_sym);
    });
    t2 = now();
    // 3. stitch: a chunk counts only if the one before it ended on its start
    uint64_t expect = next_bit_;
    bool mstart = next_member_start_;
    std::vector<size_t> order;
    for (size_t i = 0; i < K; i++) {
      Chunk& c = *chunks_[i];
      if (!c.has_start) continue;
      if (expect >= total_bits) {
        c.has_start = false;
        continue;
      }
      if (i > 0 && c.start_bit < expect) {  // lies inside what the previous chunk decoded
        c.has_start = false;
        continue;
      }
      if (c.start_bit != expect || !c.ok) {
        if (c.start_bit == expect && c.known_start)
          return fail("corrupt or truncated gzip stream (deflate data)");
        // decoded again from where the text really continues (window known: no markers needed,
        // but the same decoder is used and the markers resolved below)
        n_slow++;
        c.start_bit = expect;
        c.member_start = mstart;
        c.known_start = true;
        if (!decode_chunk(d_, n_, c, stop_of(i), (size_t)-1))
          return fail("corrupt or truncated gzip stream (deflate data)");
      }
      order.push_back(i);
      n_chunks++;
      expect = c.end_bit;
      mstart = false;
      if (c.hit_end) {
        expect = total_bits;
        for (size_t k = i + 1; k < K; k++) chunks_[k]->has_start = false;
        break;
      }
    }
    if (order.empty()) return fail("corrupt or truncated gzip stream");
    t3 = now();
    // 4. windows (sequential, 32 KiB each), then markers -> text and CRCs in parallel
    std::vector<std::vector<uint8_t>> win(order.size());
    std::vector<uint8_t> w = window_;
    for (size_t q = 0; q < order.size(); q++) {
      win[q] = w;
      Chunk& c = *chunks_[order[q]];
      // the window after this chunk: its last 32 KiB, resolved against w
      size_t ns = c.sym.size();
      std::vector<uint8_t> nw(kWin, 0);
      size_t take = std::min(ns, (size_t)kWin);
      for (size_t k = 0; k < take; k++) {
        uint16_t v = c.sym[ns - take + k];
        nw[kWin - take + k] = v < 256 ? (uint8_t)v : w[v - kMarker];
      }
      if (take < (size_t)kWin) memcpy(nw.data(), w.data() + take, kWin - take);
      w.swap(nw);
    }
    parts.resize(order.size());
    for (size_t q = 0; q < order.size(); q++) parts[q] = text_buffer(chunks_[order[q]]->sym.size());
    parallel(order.size(), [&](size_t q) {
      Chunk& c = *chunks_[order[q]];
      const std::vector<uint8_t>& pw = win[q];
      const uint16_t* s = c.sym.data();
      uint8_t* o = parts[q].p;
      size_t ns = c.sym.size();
      size_t k = 0;
      for (; k + 64 <= ns; k += 64) {  // markers are rare: 64 symbols at a time when there is none
        uint16_t any = 0;
        for (int j = 0; j < 64; j++) any |= s[k + j];
        if (!(any & kMarker)) {
          for (int j = 0; j < 64; j++) o[k + j] = (uint8_t)s[k + j];
        } else {
          for (int j = 0; j < 64; j++) {
            uint16_t v = s[k + j];
            o[k + j] = v < 256 ? (uint8_t)v : pw[v - kMarker];
          }
        }
      }
      for (; k < ns; k++) {
        uint16_t v = s[k];
        o[k] = v < 256 ? (uint8_t)v : pw[v - kMarker];
      }
      size_t a = 0;
      for (size_t e = 0; e <= c.ends.size(); e++) {
        size_t b = e < c.ends.size() ? c.ends[e].off : ns;
        c.seg_crc.push_back(hlm_crc32(0, o + a, b - a));
        a = b;
      }
    });
    t4 = now();
    // 5. member trailers; hand the text over
    for (size_t i : order) {
      Chunk& c = *chunks_[i];
      size_t a = 0;
      for (size_t e = 0; e <= c.ends.size(); e++) {
        size_t b = e < c.ends.size() ? c.ends[e].off : c.sym.size();
        run_crc_ = (uint32_t)crc32_combine(run_crc_, c.seg_crc[e], (z_off_t)(b - a));
        run_len_ += b - a;
        if (e < c.ends.size()) {
          if (run_crc_ != c.ends[e].crc || (uint32_t)run_len_ != c.ends[e].isize)
            return fail("corrupt or truncated gzip stream (CRC-32 / length of a member)");
          run_crc_ = 0;
          run_len_ = 0;
        }
        a = b;
      }
    }
    // text that expands far more than FASTQ does (long runs of one byte: up to 1032 : 1) would make
    // a round of full-size chunks hold gigabytes: the chunks shrink when a round expanded > 12 : 1
    {
      size_t text = 0;
      for (Text& part : parts) text += part.n;
      uint64_t used = (expect - next_bit_) / 8 + 1;
      if (text / used > 12 && chunk_bytes_ > ((size_t)64 << 10)) chunk_bytes_ /= 2;
    }
    if (trace)
      fprintf(stderr, "[pgzip] round: find %.1f ms, decode %.1f, stitch %.1f, resolve %.1f, hand over %.1f (%zu chunks)\n",
              (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (now() - t4) * 1e3, order.size());
    window_.swap(w);
    next_bit_ = expect;
    next_member_start_ = false;
    Chunk& last = *chunks_[order.back()];
    if (last.hit_end) {
      done_ = true;
      if (last.ends.empty() || last.ends.back().off != last.sym.size())
        return fail("corrupt or truncated gzip stream (no final block)");
    } else if (next_bit_ >= total_bits) {
      return fail("corrupt or truncated gzip stream (ends inside a member)");
    }
    return true;
  }

  const uint8_t* d_;
  size_t n_;
  int threads_;
  size_t chunk_bytes_ = (size_t)2 << 20;
  int round_no_ = 0;
  std::vector<std::unique_ptr<hlm_pgz::Chunk>> chunks_;
  uint64_t next_bit_ = 0;
  bool next_member_start_ = true;
  std::vector<uint8_t> window_ = std::vector<uint8_t>(hlm_pgz::kWin, 0);
  // rounds handed from the producer thread to fill()
  std::thread producer_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::vector<Text>> queue_;
  std::vector<Text> free_texts_;  // consumed blocks on their way back to the producer
  bool stop_ = false, finished_ = false;
  std::vector<Text> cur_;
  size_t cur_part_ = 0, cur_at_ = 0;
  uint32_t run_crc_ = 0;
  uint64_t run_len_ = 0;
  bool done_ = false, failed_ = false;
};

#endif
