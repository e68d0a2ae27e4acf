// `.depth.gz` on the GPU: int -> decimal text -> DEFLATE -> gzip members (GCI.py:99-143, the reference's
// single largest cost: it formats and gzips every base in Python).
//
// The depth text is long runs of identical short lines ("23\n23\n23\n..."), so a specialised encoder is
// enough: a new value is written as literals, every repeat of the previous line becomes LZ77 matches at
// distance = line length (up to 258 bytes per 13-bit match), fixed Huffman codes (RFC 1951 §3.2.6).  The
// reference already writes a multi-member gzip file (one member per thread slice, `cat`-ed together,
// GCI.py:134-143), so every chunk of GZ_CHUNK positions becomes its own member with its own CRC-32 and
// ISIZE: chunks are independent, one thread encodes one chunk straight from the int32 depth array (the
// text never exists in memory), and only the compressed bytes cross PCIe.
#include <algorithm>

#include "common.cuh"

constexpr int GZ_CHUNK = 8192;                 // positions per gzip member
constexpr int GZ_SLOT = GZ_CHUNK * 14 + 512;   // worst case: 12 text bytes per position at 9 bits each + framing
constexpr int GZ_THREADS = 32;

__constant__ uint32_t c_crc_table[256];

struct BitWriter {
  uint8_t* p;
  unsigned long long buf;
  int n;
  __device__ __forceinline__ void put(uint32_t v, int bits) {   // LSB-first
    buf |= (unsigned long long)v << n;
    n += bits;
    while (n >= 8) {
      *p++ = (uint8_t)buf;
      buf >>= 8;
      n -= 8;
    }
  }
  __device__ __forceinline__ void put_huff(uint32_t code, int bits) {   // Huffman codes go MSB-first
    put(__brev(code) >> (32 - bits), bits);
  }
  __device__ __forceinline__ void literal(uint32_t c) {
    if (c < 144) put_huff(0x30 + c, 8); else put_huff(0x190 + (c - 144), 9);
  }
  __device__ __forceinline__ void flush_byte() {
    if (n > 0) {
      *p++ = (uint8_t)buf;
      buf = 0;
      n = 0;
    }
  }
};

// one match of `len` (3..258) bytes at distance `dist` (1..16)
__device__ __forceinline__ void put_match(BitWriter& w, int len, int dist) {
  // length symbol (RFC 1951 §3.2.5)
  int sym, extra_bits, extra;
  if (len == 258) { sym = 285; extra_bits = 0; extra = 0; }
  else if (len <= 10) { sym = 254 + len; extra_bits = 0; extra = 0; }
  else {
    const int l = len - 3;                          // 8..254
    const int k = 31 - __clz(l) - 2;                // extra bits: 1..5
    sym = 257 + 4 * (k + 1) + ((l >> k) & 3);
    extra_bits = k;
    extra = l & ((1 << k) - 1);
  }
  if (sym < 280) w.put_huff(sym - 256, 7); else w.put_huff(0xC0 + (sym - 280), 8);
  if (extra_bits) w.put(extra, extra_bits);
  // distance symbol: 1..4 -> 0..3; 5-6 -> 4; 7-8 -> 5; 9-12 -> 6; 13-16 -> 7
  const int d = dist - 1;
  int dsym, dbits, dextra;
  if (d < 4) { dsym = d; dbits = 0; dextra = 0; }
  else {
    const int k = 31 - __clz(d) - 1;                // 1 or 2 here
    dsym = 2 * (k + 1) + ((d >> k) & 1);
    dbits = k;
    dextra = d & ((1 << k) - 1);
  }
  w.put_huff(dsym, 5);
  if (dbits) w.put(dextra, dbits);
}

__device__ __forceinline__ void put_matches(BitWriter& w, long long bytes, int dist, const uint8_t* line, int len) {
  // `bytes` more bytes repeating the last `dist`-byte line; matches must be >= 3 bytes long
  if (bytes <= 0) return;
  if (bytes < 3) {
    for (int k = 0; k < (int)bytes; k++) w.literal(line[k % len]);
    return;
  }
  while (bytes > 0) {
    int m = bytes > 258 ? 258 : (int)bytes;
    if (bytes - m > 0 && bytes - m < 3) m -= 3;     // never leave a 1- or 2-byte tail
    put_match(w, m, dist);
    bytes -= m;
  }
}

__device__ __forceinline__ int format_line(int v, uint8_t* line) {
  unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
  uint8_t tmp[12];
  int n = 0;
  do {
    tmp[n++] = (uint8_t)('0' + u % 10u);
    u /= 10u;
  } while (u);
  int k = 0;
  if (v < 0) line[k++] = '-';
  while (n) line[k++] = tmp[--n];
  line[k++] = '\n';
  return k;
}

__global__ void __launch_bounds__(GZ_THREADS)
gzip_chunks_kernel(const int32_t* __restrict__ depth, int64_t count, const uint8_t* __restrict__ header,
                   int header_len, uint8_t* __restrict__ slots, int32_t* __restrict__ sizes, int64_t n_chunks) {
  const int64_t chunk = blockIdx.x * (int64_t)GZ_THREADS + threadIdx.x;
  if (chunk >= n_chunks) return;
  const int64_t i0 = chunk * GZ_CHUNK;
  const int64_t i1 = min(count, i0 + GZ_CHUNK);
  uint8_t* out = slots + chunk * GZ_SLOT;
  // gzip header: magic, deflate, no flags, mtime 0, xfl 0, OS unknown
  const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, 0, 0xff};
  for (int k = 0; k < 10; k++) out[k] = hdr[k];
  BitWriter w{out + 10, 0ull, 0};
  w.put(3, 3);                                       // BFINAL = 1, BTYPE = 01 (fixed Huffman)
  uint32_t crc = 0xffffffffu;
  uint32_t isize = 0;
  if (chunk == 0) {
    for (int k = 0; k < header_len; k++) {
      const uint8_t c = header[k];
      w.literal(c);
      crc = c_crc_table[(crc ^ c) & 0xffu] ^ (crc >> 8);
    }
    isize += header_len;
  }
  uint8_t line[12];
  int len = 0, prev = 0;
  bool have = false;
  long long pending = 0;                             // bytes repeating the current line, not yet emitted
  for (int64_t i = i0; i < i1; i++) {
    const int v = depth[i];
    if (have && v == prev) {
      pending += len;
    } else {
      put_matches(w, pending, len, line, len);
      pending = 0;
      len = format_line(v, line);
      for (int k = 0; k < len; k++) w.literal(line[k]);
      prev = v;
      have = true;
    }
    for (int k = 0; k < len; k++) crc = c_crc_table[(crc ^ line[k]) & 0xffu] ^ (crc >> 8);
    isize += len;
  }
  put_matches(w, pending, len, line, len);
  w.put_huff(0, 7);                                  // end of block
  w.flush_byte();
  crc = ~crc;
  uint8_t* p = w.p;
  for (int k = 0; k < 4; k++) *p++ = (uint8_t)(crc >> (8 * k));
  for (int k = 0; k < 4; k++) *p++ = (uint8_t)(isize >> (8 * k));
  sizes[chunk] = (int32_t)(p - out);
}

__global__ void gzip_pack_kernel(const uint8_t* __restrict__ slots, const int32_t* __restrict__ sizes,
                                 const int64_t* __restrict__ off, uint8_t* __restrict__ out) {
  const int64_t chunk = blockIdx.x;
  const uint8_t* src = slots + chunk * GZ_SLOT;
  uint8_t* dst = out + off[chunk];
  for (int k = threadIdx.x; k < sizes[chunk]; k += blockDim.x) dst[k] = src[k];
}

static bool g_crc_ready[64] = {false};

static int ensure_crc_table(gci_ctx* ctx) {
  if (ctx->device < 64 && g_crc_ready[ctx->device]) return GCI_OK;
  uint32_t t[256];
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    t[i] = c;
  }
  GCI_CUDA_TRY(ctx, cudaMemcpyToSymbol(c_crc_table, t, sizeof t));
  if (ctx->device < 64) g_crc_ready[ctx->device] = true;
  return GCI_OK;
}

extern "C" int gci_depth_gzip(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count,
                              const char* header, int32_t header_len, char* out, int64_t cap, int64_t* n_bytes) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_bytes || header_len < 0 || (header_len && !header))
    return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_gzip: track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || first < 0 || count < 0 ||
      first + count > ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_depth_gzip: bad range");
  if (header_len > 4096) return ctx->fail(GCI_E_ARG, "gci_depth_gzip: header too long");
  GCI_TRY(ensure_crc_table(ctx));
  *n_bytes = 0;
  if (header_len > 400) return ctx->fail(GCI_E_ARG, "gci_depth_gzip: header longer than 400 bytes");
  DevBuf &slots = ctx->tmp[9], &sizes = ctx->tmp[8], &offs = ctx->tmp[7], &hdr = ctx->tmp[6], &packed = ctx->tmp[3];
  // size query (out == NULL) and fetch (out != NULL) of the same range share one encoding pass: the packed
  // members stay on the device between the two calls
  unsigned long long key = 1469598103934665603ull;
  auto mix = [&key](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
  mix((unsigned long long)track); mix((unsigned long long)contig); mix((unsigned long long)first);
  mix((unsigned long long)count);
  for (int k = 0; k < header_len; k++) mix((unsigned char)header[k]);
  if (!(ctx->gz_valid && ctx->gz_key == key)) {
    ctx->gz_valid = false;
    const int64_t n_chunks = std::max<int64_t>(1, (count + GZ_CHUNK - 1) / GZ_CHUNK);
    const int32_t* d = t.depth.as<int32_t>() + ctx->pos_off[contig] + first;
    GCI_TRY(ctx->ensure(slots, (size_t)n_chunks * GZ_SLOT));
    GCI_TRY(ctx->ensure(sizes, 4 * (size_t)n_chunks));
    GCI_TRY(ctx->ensure(offs, 8 * (size_t)(n_chunks + 1)));
    ctx->stage_begin(GCI_ST_TEXT);
    if (header_len) GCI_TRY(gci_h2d(ctx, hdr, header, (size_t)header_len));
    gzip_chunks_kernel<<<(unsigned)((n_chunks + GZ_THREADS - 1) / GZ_THREADS), GZ_THREADS, 0, ctx->stream>>>(
        d, count, header_len ? hdr.as<uint8_t>() : nullptr, header_len, slots.as<uint8_t>(), sizes.as<int32_t>(),
        n_chunks);
    GCI_LAUNCH_CHECK(ctx);
    GCI_TRY(gci_exclusive_scan_i64_from_i32(ctx, sizes.as<int32_t>(), offs.as<int64_t>(), n_chunks,
                                            offs.as<int64_t>() + n_chunks));
    int64_t* h_total = (int64_t*)ctx->pinned(sizeof(int64_t));
    if (!h_total) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
    GCI_TRY(gci_d2h(ctx, h_total, offs.as<int64_t>() + n_chunks, sizeof(int64_t)));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->gz_total = *h_total;
    GCI_TRY(ctx->ensure(packed, (size_t)ctx->gz_total));
    gzip_pack_kernel<<<(unsigned)n_chunks, 128, 0, ctx->stream>>>(slots.as<uint8_t>(), sizes.as<int32_t>(),
                                                                 offs.as<int64_t>(), packed.as<uint8_t>());
    GCI_LAUNCH_CHECK(ctx);
    ctx->stage_end();
    ctx->gz_key = key;
    ctx->gz_valid = true;
  }
  *n_bytes = ctx->gz_total;
  if (!out) return GCI_OK;
  if (cap < ctx->gz_total)
    return ctx->fail(GCI_E_ARG, "gzip buffer too small (%lld < %lld)", (long long)cap, (long long)ctx->gz_total);
  ctx->stage_begin(GCI_ST_D2H);
  GCI_TRY(gci_d2h(ctx, out, packed.p, (size_t)ctx->gz_total));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->gz_valid = false;     // the scratch buffers are shared with other entry points
  return GCI_OK;
}
