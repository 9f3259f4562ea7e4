// A DEFLATE (RFC 1951) decoder for BGZF members: see inflate.h.
//
// Written from the format specification: canonical Huffman codes decoded through two-level tables (one lookup for every code of
// up to 11 / 8 bits), a 64-bit bit buffer refilled with one unaligned load, literals and matches written straight into the
// caller's buffer, matches copied eight bytes at a time where source and destination are far enough apart.  Two loops: a fast
// one while input and output both have slack (no per-symbol bounds checks), a careful one for the ends.  Every failure mode
// (bad code lengths, a distance before the start of the output, input or output exhausted, output not filled exactly) returns
// false and the caller falls back to zlib, so a member this decoder cannot take is never a wrong answer.
#include "inflate.h"

#include <cstring>

namespace brq {

namespace {

constexpr int LIT_BITS = 11, DIST_BITS = 8;
// table entry: [7:0] bits to consume, [11:8] extra bits, [15:12] kind, [31:16] value (literal / base / subtable offset)
enum : uint32_t { K_LITERAL = 0, K_BASE = 1, K_END = 2, K_SUBTABLE = 3, K_INVALID = 4 };
inline uint32_t entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t bits) { return value << 16 | kind << 12 | extra << 8 | bits; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

struct Tables {
  uint32_t lit[(1 << LIT_BITS) + 2048];    // primary + subtables (at most 2^15 - 2^11 codes behind them, far fewer in practice)
  uint32_t dist[(1 << DIST_BITS) + 1024];
};

inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = r << 1 | (v & 1); v >>= 1; }
  return r;
}

// Builds the two-level decode table of a canonical prefix code.  lens[s] = code length of symbol s (0 = unused).
// what(s, remaining bits) -> the entry of symbol s.  Returns false for an over-subscribed code, or an incomplete one unless it
// has a single code (which the format allows for distances).
template <class What>
bool build_table(const uint8_t* lens, int n_sym, int primary_bits, uint32_t* table, size_t capacity, What&& what) {
  int count[16] = {0};
  for (int s = 0; s < n_sym; ++s) ++count[lens[s]];
  count[0] = 0;
  int left = 1, n_codes = 0;
  for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; n_codes += count[l]; }
  if (left > 0 && n_codes != 1) return false;
  uint32_t next_code[16];
  { uint32_t code = 0; for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; } }
  const uint32_t primary_size = 1u << primary_bits;
  for (uint32_t i = 0; i < primary_size; ++i) table[i] = entry(0, K_INVALID, 0, 1);
  // the longest code behind every primary prefix decides its subtable's size
  uint8_t sub_bits[1 << LIT_BITS];
  memset(sub_bits, 0, primary_size);
  uint32_t codes[288 + 32];
  {
    uint32_t nc[16];
    memcpy(nc, next_code, sizeof nc);
    for (int s = 0; s < n_sym; ++s) {
      const int l = lens[s];
      if (!l) continue;
      const uint32_t rev = reverse_bits(nc[l]++, l);
      codes[s] = rev;
      if (l > primary_bits) { uint8_t& b = sub_bits[rev & (primary_size - 1)]; if (l - primary_bits > b) b = (uint8_t)(l - primary_bits); }
    }
  }
  size_t used = primary_size;
  for (uint32_t i = 0; i < primary_size; ++i) {
    if (!sub_bits[i]) continue;
    const size_t size = (size_t)1 << sub_bits[i];
    if (used + size > capacity) return false;
    table[i] = entry((uint32_t)used, K_SUBTABLE, sub_bits[i], (uint32_t)primary_bits);
    for (size_t k = 0; k < size; ++k) table[used + k] = entry(0, K_INVALID, 0, 1);
    used += size;
  }
  for (int s = 0; s < n_sym; ++s) {
    const int l = lens[s];
    if (!l) continue;
    const uint32_t rev = codes[s];
    if (l <= primary_bits) {
      const uint32_t e = what(s, (uint32_t)l);
      for (uint32_t i = rev; i < primary_size; i += 1u << l) table[i] = e;
    } else {
      const uint32_t p = table[rev & (primary_size - 1)];
      const uint32_t base = p >> 16, bits = (p >> 8) & 15u;
      const uint32_t e = what(s, (uint32_t)(l - primary_bits));
      for (uint32_t i = rev >> primary_bits; i < (1u << bits); i += 1u << (l - primary_bits)) table[base + i] = e;
    }
  }
  return true;
}

inline uint32_t lit_entry(int s, uint32_t bits) {
  if (s < 256) return entry((uint32_t)s, K_LITERAL, 0, bits);
  if (s == 256) return entry(0, K_END, 0, bits);
  if (s > 285) return entry(0, K_INVALID, 0, bits);
  return entry(LEN_BASE[s - 257], K_BASE, LEN_EXTRA[s - 257], bits);
}
inline uint32_t dist_entry(int s, uint32_t bits) {
  if (s > 29) return entry(0, K_INVALID, 0, bits);
  return entry(DIST_BASE[s], K_BASE, DIST_EXTRA[s], bits);
}

struct Bits {
  const uint8_t* in;
  const uint8_t* in_end;
  uint64_t buf = 0;
  int cnt = 0;    // bits of buf that count (bits above them, if any, are the input's next bits already: refill ORs them again)
  int past = 0;   // zero bytes appended behind the input's end so far (a stream that consumes them is cut short)
  // at least 56 bits that count: one unaligned load while eight input bytes are readable, byte by byte near the end
  inline void refill() {
    if (in + 8 <= in_end) {
      uint64_t w;
      memcpy(&w, in, 8);
      buf |= w << cnt;
      in += (63 - cnt) >> 3;
      cnt |= 56;
    } else {
      while (cnt <= 56) {
        if (in < in_end) buf |= (uint64_t)*in++ << cnt; else ++past;
        cnt += 8;
      }
    }
  }
  inline bool overrun() const { return cnt < 8 * past; }   // bits were taken from behind the input's end
  inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
  inline void drop(int n) { buf >>= n; cnt -= n; }
};

}  // namespace

bool fast_inflate(const uint8_t* src, size_t src_len, uint8_t* dst, size_t dst_len) {
  Bits b;
  b.in = src; b.in_end = src + src_len;
  uint8_t* out = dst;
  uint8_t* const out_end = dst + dst_len;
  static thread_local Tables T;
  static thread_local bool fixed_built = false;
  static thread_local Tables F;   // the fixed code's tables
  bool last = false;
  while (!last) {
    b.refill();
    if (b.overrun()) return false;
    last = b.peek(1) != 0; b.drop(1);
    const uint32_t type = b.peek(2); b.drop(2);
    const Tables* tab;
    if (type == 0) {  // stored: skip to a byte boundary, LEN, NLEN, bytes
      b.drop(b.cnt & 7);
      if (b.overrun()) return false;
      // the bit buffer holds whole bytes of the input: give them back
      const uint8_t* p = b.in - ((b.cnt >> 3) - b.past);
      b.past = 0;
      b.buf = 0; b.cnt = 0;
      if (p + 4 > b.in_end) return false;
      const uint32_t len = p[0] | p[1] << 8, nlen = p[2] | p[3] << 8;
      if ((len ^ nlen) != 0xFFFFu) return false;
      p += 4;
      if (p + len > b.in_end || out + len > out_end) return false;
      memcpy(out, p, len);
      out += len; b.in = p + len;
      continue;
    } else if (type == 1) {
      if (!fixed_built) {
        uint8_t lens[288 + 32];
        for (int s = 0; s < 144; ++s) lens[s] = 8;
        for (int s = 144; s < 256; ++s) lens[s] = 9;
        for (int s = 256; s < 280; ++s) lens[s] = 7;
        for (int s = 280; s < 288; ++s) lens[s] = 8;
        if (!build_table(lens, 288, LIT_BITS, F.lit, sizeof F.lit / 4, lit_entry)) return false;
        for (int s = 0; s < 32; ++s) lens[s] = 5;
        if (!build_table(lens, 32, DIST_BITS, F.dist, sizeof F.dist / 4, dist_entry)) return false;
        fixed_built = true;
      }
      tab = &F;
    } else if (type == 2) {
      b.refill();
      const uint32_t hlit = b.peek(5) + 257; b.drop(5);
      const uint32_t hdist = b.peek(5) + 1; b.drop(5);
      const uint32_t hclen = b.peek(4) + 4; b.drop(4);
      if (hlit > 286 || hdist > 30) return false;
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t cl[19] = {0};
      for (uint32_t i = 0; i < hclen; ++i) { if (b.cnt < 3) b.refill(); cl[order[i]] = (uint8_t)b.peek(3); b.drop(3); }
      uint32_t cl_table[1 << 7];
      if (!build_table(cl, 19, 7, cl_table, 1 << 7, [](int s, uint32_t bits) { return entry((uint32_t)s, K_LITERAL, 0, bits); })) return false;
      uint8_t lens[288 + 32 + 138];
      uint32_t i = 0;
      while (i < hlit + hdist) {
        b.refill();
        if (b.overrun()) return false;
        const uint32_t e = cl_table[b.peek(7)];
        if (((e >> 12) & 15u) != K_LITERAL) return false;
        b.drop((int)(e & 255u));
        const uint32_t s = e >> 16;
        if (s < 16) { lens[i++] = (uint8_t)s; continue; }
        uint32_t rep, val = 0;
        if (s == 16) { if (!i) return false; val = lens[i - 1]; rep = 3 + b.peek(2); b.drop(2); }
        else if (s == 17) { rep = 3 + b.peek(3); b.drop(3); }
        else { rep = 11 + b.peek(7); b.drop(7); }
        if (i + rep > hlit + hdist) return false;
        memset(lens + i, (int)val, rep);
        i += rep;
      }
      if (!lens[256]) return false;   // no end-of-block code
      if (!build_table(lens, (int)hlit, LIT_BITS, T.lit, sizeof T.lit / 4, lit_entry)) return false;
      if (!build_table(lens + hlit, (int)hdist, DIST_BITS, T.dist, sizeof T.dist / 4, dist_entry)) {
        // a block of literals only may declare one unused distance code of length zero
        bool none = true;
        for (uint32_t k = 0; k < hdist; ++k) none = none && lens[hlit + k] == 0;
        if (!none) return false;
        for (uint32_t k = 0; k < (1u << DIST_BITS); ++k) T.dist[k] = entry(0, K_INVALID, 0, 1);
      }
      tab = &T;
    } else {
      return false;
    }

    // ---- the block's symbols
    const uint32_t* lit = tab->lit;
    const uint32_t* dist = tab->dist;
    for (;;) {
      // fast loop: every symbol needs at most 15 + 5 + 15 + 13 = 48 bits; a match writes at most 258 (+ 7 of overshoot) bytes
      while (b.in + 8 <= b.in_end && out_end - out >= 258 + 16) {
        b.refill();
        uint32_t e = lit[b.peek(LIT_BITS)];
        if (((e >> 12) & 15u) == K_SUBTABLE) { b.drop(LIT_BITS); e = lit[(e >> 16) + b.peek((int)((e >> 8) & 15u))]; }
        b.drop((int)(e & 255u));
        uint32_t kind = (e >> 12) & 15u;
        if (kind == K_LITERAL) {
          *out++ = (uint8_t)(e >> 16);
          // a second literal from the same refill, more often than not (BAM payloads are mostly literals; a third and a fourth
          // were measured and do not pay)
          e = lit[b.peek(LIT_BITS)];
          if (((e >> 12) & 15u) == K_LITERAL) { b.drop((int)(e & 255u)); *out++ = (uint8_t)(e >> 16); }
          continue;
        }
        if (kind != K_BASE) { if (kind == K_END) goto block_done; return false; }
        {
          const uint32_t lx = (e >> 8) & 15u;
          const uint32_t len = (e >> 16) + b.peek((int)lx);
          b.drop((int)lx);
          uint32_t d = dist[b.peek(DIST_BITS)];
          if (((d >> 12) & 15u) == K_SUBTABLE) { b.drop(DIST_BITS); d = dist[(d >> 16) + b.peek((int)((d >> 8) & 15u))]; }
          b.drop((int)(d & 255u));
          if (((d >> 12) & 15u) != K_BASE) return false;
          const uint32_t dx = (d >> 8) & 15u;
          const uint32_t distance = (d >> 16) + b.peek((int)dx);
          b.drop((int)dx);
          if (distance > (size_t)(out - dst)) return false;
          const uint8_t* from = out - distance;
          uint8_t* to = out;
          out += len;
          if (distance >= 8) {
            do { uint64_t w; memcpy(&w, from, 8); memcpy(to, &w, 8); from += 8; to += 8; } while (to < out);
          } else if (distance == 1) {
            memset(to, *from, len);
          } else {
            do { *to++ = *from++; } while (to < out);
          }
        }
      }
      // careful loop: one symbol at a time with every bound checked
      {
        b.refill();
        if (b.overrun()) return false;
        uint32_t e = lit[b.peek(LIT_BITS)];
        if (((e >> 12) & 15u) == K_SUBTABLE) { b.drop(LIT_BITS); e = lit[(e >> 16) + b.peek((int)((e >> 8) & 15u))]; }
        b.drop((int)(e & 255u));
        if (b.overrun()) return false;
        const uint32_t kind = (e >> 12) & 15u;
        if (kind == K_LITERAL) { if (out >= out_end) return false; *out++ = (uint8_t)(e >> 16); continue; }
        if (kind == K_END) goto block_done;
        if (kind != K_BASE) return false;
        const uint32_t lx = (e >> 8) & 15u;
        const uint32_t len = (e >> 16) + b.peek((int)lx);
        b.drop((int)lx);
        uint32_t d = dist[b.peek(DIST_BITS)];
        if (((d >> 12) & 15u) == K_SUBTABLE) { b.drop(DIST_BITS); d = dist[(d >> 16) + b.peek((int)((d >> 8) & 15u))]; }
        b.drop((int)(d & 255u));
        if (((d >> 12) & 15u) != K_BASE) return false;
        const uint32_t dx = (d >> 8) & 15u;
        const uint32_t distance = (d >> 16) + b.peek((int)dx);
        b.drop((int)dx);
        if (b.overrun()) return false;
        if (distance > (size_t)(out - dst) || len > (size_t)(out_end - out)) return false;
        const uint8_t* from = out - distance;
        for (uint32_t k = 0; k < len; ++k) out[k] = from[k];
        out += len;
      }
    }
  block_done:;
  }
  return out == out_end && !b.overrun();
}

}  // namespace brq
