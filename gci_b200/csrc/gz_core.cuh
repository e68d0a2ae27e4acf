// `.depth.gz` encoder core (GCI.py:99-143): the per-run arithmetic shared by the CUDA kernels of gzip.cu and a
// host harness (tests/test_gz_core_host.py compiles this header with g++ and inflates the result with zlib).
//
// The depth text is a sequence of RUNS: value v repeated k times = the line "v\n" (Lb bytes) k times.  A run is
// encoded as the literals of one line followed by LZ77 matches at distance Lb (3..258 bytes each) in one fixed-
// Huffman DEFLATE block (RFC 1951 §3.2.6) per gzip member.  Everything a run contributes — its bit count, its bits,
// the CRC-32 of its text — depends only on (v, k), so runs are encoded independently and stitched together with
// prefix sums: bit offsets add, and CRCs combine through
//     R(A || B) = R(A) * x^(8|B|)  +  R(B)          in GF(2)[x] / P        (R = CRC register with zero init / xorout)
//     R(line^k) = sum over the set bits j of k, high to low:  acc = acc * x^(8 Lb 2^j) + R(line^(2^j))
// with table-driven powers, so a run of thousands of identical lines costs a handful of 32-bit carry-less
// multiplications instead of a byte-serial walk over its text.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GZ_HD __host__ __device__ __forceinline__
#else
#define GZ_HD inline
#endif

constexpr uint32_t GZ_POLY = 0xEDB88320u;      // CRC-32 (reflected)
constexpr int GZ_MEMBER = 8192;                // positions per gzip member
constexpr int GZ_POW_BITS = 6;                 // x^(8n): n split into 6-bit digits
constexpr int GZ_POW_DIGITS = 4;               // n < 2^24 text bytes per member (8192 positions x 12 bytes = 98 304)
constexpr int GZ_RUN_LOG = 14;                 // runs of up to 2^14 - 1 lines inside one member (8192 positions)
constexpr int GZ_TBL_VALUES = 1024;            // depth values with a precomputed R(line^(2^j)) row

// a * b mod P, operands and result in the reflected representation (bit 31 = x^0), as zlib's multmodp
GZ_HD uint32_t gz_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    p ^= (a & 0x80000000u) ? b : 0u;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1u) ? GZ_POLY : 0u);
  }
  return p;
}

struct GzTables {
  const uint32_t* crc_byte;   // [256]  classic byte table
  const uint32_t* pow8;       // [GZ_POW_DIGITS][64]: x^(8 * d * 64^i)
  const uint32_t* line_pow;   // [12][GZ_RUN_LOG]: x^(8 * Lb * 2^j)
  const uint32_t* run_tbl;    // [GZ_TBL_VALUES][GZ_RUN_LOG]: R(line_v ^ (2^j))
};

// x^(8 n) mod P
GZ_HD uint32_t gz_pow8(const GzTables& t, uint32_t n) {
  uint32_t r = t.pow8[n & 63u];
  n >>= GZ_POW_BITS;
#pragma unroll 1
  for (int i = 1; i < GZ_POW_DIGITS && n; i++, n >>= GZ_POW_BITS)
    if (n & 63u) r = gz_mulmod(r, t.pow8[i * 64 + (n & 63u)]);
  return r;
}

// the line "v\n": bytes packed little-endian into (lo: bytes 0-7, hi: bytes 8-11), returns its length
GZ_HD int gz_format_line(int32_t v, uint64_t* lo, uint32_t* hi) {
  uint32_t u = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
  uint8_t tmp[12];
  int n = 0;
  do {
    tmp[n++] = (uint8_t)('0' + u % 10u);
    u /= 10u;
  } while (u);
  uint64_t l = 0;
  uint32_t h = 0;
  int k = 0;
  if (v < 0) { l |= (uint64_t)'-'; k = 1; }
  while (n) {
    const uint64_t c = tmp[--n];
    if (k < 8) l |= c << (8 * k); else h |= (uint32_t)c << (8 * (k - 8));
    k++;
  }
  if (k < 8) l |= (uint64_t)'\n' << (8 * k); else h |= (uint32_t)'\n' << (8 * (k - 8));
  *lo = l;
  *hi = h;
  return k + 1;
}

GZ_HD uint32_t gz_line_byte(uint64_t lo, uint32_t hi, int i) {
  return i < 8 ? (uint32_t)(lo >> (8 * i)) & 0xffu : (hi >> (8 * (i - 8))) & 0xffu;
}

// R(bytes) continued from register state r (zero init / no xorout): the classic table walk
GZ_HD uint32_t gz_crc_line(const GzTables& t, uint32_t r, uint64_t lo, uint32_t hi, int Lb) {
#pragma unroll 1
  for (int i = 0; i < Lb; i++) r = t.crc_byte[(r ^ gz_line_byte(lo, hi, i)) & 0xffu] ^ (r >> 8);
  return r;
}

// R(line^k), 1 <= k < 2^GZ_RUN_LOG
GZ_HD uint32_t gz_crc_run(const GzTables& t, int32_t v, uint64_t lo, uint32_t hi, int Lb, uint32_t k) {
  const uint32_t* lp = t.line_pow + Lb * GZ_RUN_LOG;
  uint32_t acc = 0;
  if (v >= 0 && v < GZ_TBL_VALUES) {
    const uint32_t* row = t.run_tbl + (size_t)v * GZ_RUN_LOG;
#pragma unroll 1
    for (int j = GZ_RUN_LOG - 1; j >= 0; j--)
      if ((k >> j) & 1u) acc = (acc ? gz_mulmod(acc, lp[j]) : 0u) ^ row[j];
    return acc;
  }
  // value without a table row: build R(line^(2^j)) on the fly (rare: depth >= 1024 or negative)
  uint32_t blk[GZ_RUN_LOG];
  blk[0] = gz_crc_line(t, 0u, lo, hi, Lb);
  int top = 0;
  while ((k >> (top + 1)) != 0) top++;
  for (int j = 0; j < top; j++) blk[j + 1] = gz_mulmod(blk[j], lp[j]) ^ blk[j];
  for (int j = top; j >= 0; j--)
    if ((k >> j) & 1u) acc = (acc ? gz_mulmod(acc, lp[j]) : 0u) ^ blk[j];
  return acc;
}

// ---- DEFLATE bit arithmetic (fixed Huffman) ------------------------------------------------------------
// bits of one match of `len` (3..258) bytes at distance `dist` (1..12)
GZ_HD int gz_match_bits(int len, int dist) {
  int bits;
  if (len == 258) bits = 8;                         // symbol 285
  else if (len <= 10) bits = 7;                     // symbols 257..264
  else {
    const int l = len - 3;                          // 8..254
    int kx = 0;
    while ((l >> (kx + 3)) != 0) kx++;              // extra bits 1..5: floor(log2 l) - 2
    const int sym = 257 + 4 * (kx + 1) + ((l >> kx) & 3);
    bits = (sym < 280 ? 7 : 8) + kx;
  }
  const int d = dist - 1;
  bits += 5 + (d < 4 ? 0 : (d < 8 ? 1 : 2));
  return bits;
}

// how `rem` (>= 3) bytes repeating at distance Lb split into matches: n258 full ones, then up to two shorter ones
struct GzSplit { uint32_t n258; int a, b; };         // a, b: lengths of the trailing matches (0 = none)
GZ_HD GzSplit gz_split(uint32_t rem) {
  GzSplit s;
  s.n258 = rem / 258u;
  const int r = (int)(rem % 258u);
  s.a = s.b = 0;
  if (r >= 3) s.a = r;
  else if (r > 0) { s.n258 -= 1; s.a = 255; s.b = r + 3; }   // never leave a 1- or 2-byte tail (rem >= 3 => n258 >= 1)
  return s;
}

// bits of a whole run: one literal line + (k - 1) repeats
GZ_HD uint64_t gz_run_bits(int Lb, uint32_t k) {
  uint64_t bits = 8ull * Lb;                          // '-', digits and '\n' are all 8-bit literals (< 144)
  const uint32_t rem = (uint32_t)Lb * (k - 1u);
  if (rem == 0) return bits;
  if (rem < 3) return bits + 8ull * rem;
  const GzSplit s = gz_split(rem);
  bits += (uint64_t)s.n258 * gz_match_bits(258, Lb);
  if (s.a) bits += gz_match_bits(s.a, Lb);
  if (s.b) bits += gz_match_bits(s.b, Lb);
  return bits;
}

// ---- bit emission ---------------------------------------------------------------------------------------
// Sink: `void or_word(uint64_t word_index, uint32_t bits)` ORs into a zero-initialised 32-bit word array
template <typename Sink>
struct GzBitPut {
  Sink& sink;
  uint64_t word;     // next word index
  uint64_t buf;
  int fill;
  GZ_HD GzBitPut(Sink& s, uint64_t bit_offset) : sink(s), word(bit_offset >> 5), buf(0), fill((int)(bit_offset & 31)) {}
  GZ_HD void put(uint32_t v, int bits) {              // LSB first
    buf |= (uint64_t)v << fill;
    fill += bits;
    if (fill >= 32) {
      sink.or_word(word++, (uint32_t)buf);
      buf >>= 32;
      fill -= 32;
    }
  }
  GZ_HD void flush() {
    if (fill > 0 && (uint32_t)buf) sink.or_word(word, (uint32_t)buf);
    buf = 0;
  }
};

GZ_HD uint32_t gz_rev(uint32_t v, int bits) {         // Huffman codes go MSB first
#if defined(__CUDA_ARCH__)
  return __brev(v) >> (32 - bits);
#else
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

template <typename Sink>
GZ_HD void gz_put_literal(GzBitPut<Sink>& w, uint32_t c) {
  if (c < 144) w.put(gz_rev(0x30 + c, 8), 8); else w.put(gz_rev(0x190 + (c - 144), 9), 9);
}

template <typename Sink>
GZ_HD void gz_put_match(GzBitPut<Sink>& w, int len, int dist) {
  int sym, kx = 0, extra = 0;
  if (len == 258) sym = 285;
  else if (len <= 10) sym = 254 + len;
  else {
    const int l = len - 3;
    while ((l >> (kx + 3)) != 0) kx++;
    sym = 257 + 4 * (kx + 1) + ((l >> kx) & 3);
    extra = l & ((1 << kx) - 1);
  }
  if (sym < 280) w.put(gz_rev(sym - 256, 7), 7); else w.put(gz_rev(0xC0 + (sym - 280), 8), 8);
  if (kx) w.put((uint32_t)extra, kx);
  const int d = dist - 1;
  if (d < 4) w.put(gz_rev(d, 5), 5);
  else {
    const int kd = d < 8 ? 1 : 2;
    w.put(gz_rev(2 * (kd + 1) + ((d >> kd) & 1), 5), 5);
    w.put((uint32_t)(d & ((1 << kd) - 1)), kd);
  }
}

template <typename Sink>
GZ_HD void gz_put_run(GzBitPut<Sink>& w, uint64_t lo, uint32_t hi, int Lb, uint32_t k) {
  for (int i = 0; i < Lb; i++) gz_put_literal(w, gz_line_byte(lo, hi, i));
  const uint32_t rem = (uint32_t)Lb * (k - 1u);
  if (rem == 0) return;
  if (rem < 3) {
    for (uint32_t i = 0; i < rem; i++) gz_put_literal(w, gz_line_byte(lo, hi, (int)(i % (uint32_t)Lb)));
    return;
  }
  const GzSplit s = gz_split(rem);
#pragma unroll 1
  for (uint32_t i = 0; i < s.n258; i++) gz_put_match(w, 258, Lb);
  if (s.a) gz_put_match(w, s.a, Lb);
  if (s.b) gz_put_match(w, s.b, Lb);
}

// ---- table construction (host; the CUDA side uploads the result once per context) ------------------------
// layout of one flat uint32 array: [crc_byte 256 | pow8 DIGITS*64 | line_pow 12*RUN_LOG | run_tbl VALUES*RUN_LOG]
constexpr int GZ_OFF_POW8 = 256;
constexpr int GZ_OFF_LINE_POW = GZ_OFF_POW8 + GZ_POW_DIGITS * 64;
constexpr int GZ_OFF_RUN_TBL = GZ_OFF_LINE_POW + 12 * GZ_RUN_LOG;
constexpr int GZ_TABLE_WORDS = GZ_OFF_RUN_TBL + GZ_TBL_VALUES * GZ_RUN_LOG;

inline GzTables gz_tables_view(const uint32_t* flat) {
  GzTables t;
  t.crc_byte = flat;
  t.pow8 = flat + GZ_OFF_POW8;
  t.line_pow = flat + GZ_OFF_LINE_POW;
  t.run_tbl = flat + GZ_OFF_RUN_TBL;
  return t;
}

inline void gz_build_tables(uint32_t* flat) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1u) ? GZ_POLY ^ (c >> 1) : c >> 1;
    flat[i] = c;
  }
  // x^8 in the reflected representation: bit 31 is x^0, so x^8 is bit 23
  const uint32_t x8 = 1u << 23;
  uint32_t base = x8;                                   // x^(8 * 64^i)
  for (int i = 0; i < GZ_POW_DIGITS; i++) {
    uint32_t* row = flat + GZ_OFF_POW8 + i * 64;
    row[0] = 0x80000000u;                               // x^0
    for (int d = 1; d < 64; d++) row[d] = gz_mulmod(row[d - 1], base);
    base = gz_mulmod(row[63], base);
  }
  GzTables t = gz_tables_view(flat);
  for (int Lb = 0; Lb < 12; Lb++) {
    uint32_t p = gz_pow8(t, (uint32_t)Lb);
    for (int j = 0; j < GZ_RUN_LOG; j++) {
      flat[GZ_OFF_LINE_POW + Lb * GZ_RUN_LOG + j] = p;
      p = gz_mulmod(p, p);
    }
  }
  for (int v = 0; v < GZ_TBL_VALUES; v++) {
    uint64_t lo;
    uint32_t hi;
    const int Lb = gz_format_line(v, &lo, &hi);
    uint32_t r = gz_crc_line(t, 0u, lo, hi, Lb);
    for (int j = 0; j < GZ_RUN_LOG; j++) {
      flat[GZ_OFF_RUN_TBL + v * GZ_RUN_LOG + j] = r;
      r = gz_mulmod(r, t.line_pow[Lb * GZ_RUN_LOG + j]) ^ r;
    }
  }
}
