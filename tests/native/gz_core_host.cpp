// Host harness for gci_b200/csrc/gz_core.cuh: encodes a depth array member by member with exactly the per-run
// functions the CUDA kernels use (run bits, bit emission, CRC of a run, CRC combination), sequentially.
// Built and driven by tests/test_gz_core_host.py, which inflates the result with zlib.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../gci_b200/csrc/gz_core.cuh"

struct HostSink {
  uint32_t* w;
  void or_word(uint64_t i, uint32_t v) { w[i] |= v; }
};

extern "C" int64_t gz_host_encode(const int32_t* depth, int64_t n, const char* header, int header_len, uint8_t* out,
                                  int64_t cap) {
  static std::vector<uint32_t> flat;
  if (flat.empty()) { flat.resize(GZ_TABLE_WORDS); gz_build_tables(flat.data()); }
  const GzTables t = gz_tables_view(flat.data());
  memset(out, 0, (size_t)cap);
  HostSink sink{reinterpret_cast<uint32_t*>(out)};
  int64_t off = 0;                                             // byte offset of the current member
  const int64_t n_members = n > 0 ? (n + GZ_MEMBER - 1) / GZ_MEMBER : 1;
  for (int64_t m = 0; m < n_members; m++) {
    const int64_t a = m * GZ_MEMBER, b = n < a + GZ_MEMBER ? n : a + GZ_MEMBER;
    if (off + 64 + (b - a) * 14 + header_len * 2 > cap) return -1;
    const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
    for (int i = 0; i < 10; i++) out[off + i] |= hdr[i];
    GzBitPut<HostSink> w(sink, (uint64_t)(off + 10) * 8);
    w.put(3, 3);                                               // BFINAL = 1, BTYPE = 01
    uint32_t R = 0;
    uint64_t text = 0, bits = 3;
    if (m == 0)
      for (int i = 0; i < header_len; i++) {
        const uint32_t c = (uint8_t)header[i];
        gz_put_literal(w, c);
        bits += c < 144 ? 8 : 9;
        R = t.crc_byte[(R ^ c) & 0xffu] ^ (R >> 8);
        text++;
      }
    for (int64_t i = a; i < b;) {
      int64_t j = i + 1;
      while (j < b && depth[j] == depth[i]) j++;
      const uint32_t k = (uint32_t)(j - i);
      uint64_t lo;
      uint32_t hi;
      const int Lb = gz_format_line(depth[i], &lo, &hi);
      const uint64_t want = gz_run_bits(Lb, k);
      const uint64_t before = w.word * 32 + w.fill;
      gz_put_run(w, lo, hi, Lb, k);
      if (w.word * 32 + w.fill - before != want) return -2;    // bit count and emission must agree
      bits += want;
      const uint32_t tb = (uint32_t)Lb * k;
      R = gz_mulmod(R, gz_pow8(t, tb)) ^ gz_crc_run(t, depth[i], lo, hi, Lb, k);
      text += tb;
      i = j;
    }
    w.put(0, 7);                                               // end of block
    bits += 7;
    w.flush();
    const uint32_t crc = ~(gz_mulmod(0xffffffffu, gz_pow8(t, (uint32_t)text)) ^ R);
    int64_t p = off + 10 + (int64_t)((bits + 7) / 8);
    for (int i = 0; i < 4; i++) out[p++] |= (uint8_t)(crc >> (8 * i));
    for (int i = 0; i < 4; i++) out[p++] |= (uint8_t)((uint32_t)text >> (8 * i));
    off = p;
  }
  return off;
}
