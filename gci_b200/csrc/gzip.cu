// `.depth.gz` on the GPU: int -> decimal text -> DEFLATE -> gzip members (GCI.py:99-143, the reference's single
// largest cost: it formats and gzips every base in Python).
//
// Depth is piecewise constant (it only changes where a read starts or ends), so the text is a sequence of RUNS of
// identical lines and the encoder works on runs, not on bases (gz_core.cuh has the per-run arithmetic):
//   1. gz_mark_kernel    one streaming pass over the depth array: a bit per position where a run starts (value
//                        differs from the previous base, or a gzip member starts), run starts counted per tile
//   2. exclusive scan    of the tile counts -> where every tile's runs go
//   3. gz_runs_kernel    (position, value) of every run from the bitmask: 8 B per run
//   4. gz_size_kernel    warp per member (8192 positions; the reference also writes a multi-member file,
//                        GCI.py:134-143): bits of every run -> byte size of the member
//   5. exclusive scan    of the member sizes -> byte offsets in the packed output
//   6. gz_emit_kernel    warp per member: every lane encodes whole runs (one literal line + LZ77 matches at distance
//                        = line length, fixed Huffman) at its bit offset (warp prefix sum) and the CRC-32 of its text
//                        by table powers; the CRCs combine through x^(8 * bytes after the run)
// Only pass 1 touches the 4 B/base array; everything after it works on ~8 B per run (about 1/300 of the bases for
// 30x HiFi coverage).  Only the compressed bytes (about 0.03 B per base) cross PCIe.
#include <algorithm>

#include "common.cuh"
#include "gz_core.cuh"

struct GzSeg {            // one range of one contig = a sequence of gzip members (the first one carries the header)
  int64_t g0;             // global padded position of the contig's first base
  int64_t first, end;     // contig-relative positions [first, end)
  int64_t tile_off;       // index of the range's first tile in this call's tile list
  int64_t member_off;     // index of the range's first member in this call's member list
  int32_t hdr_off, hdr_len, hdr_bits;
  uint32_t hdr_crc;       // R(header): CRC register after the header bytes, zero init
};

struct GzCall {
  const GzSeg* seg;
  int32_t n_seg;
  int64_t n_tiles, n_members;
};

__device__ __forceinline__ int gz_find_seg_by_tile(const GzCall& c, int64_t w) {
  int lo = 0, hi = c.n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (c.seg[mid].tile_off <= w) lo = mid; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ int gz_find_seg_by_member(const GzCall& c, int64_t m) {
  int lo = 0, hi = c.n_seg;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (c.seg[mid].member_off <= m) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- pass 1: run-start bits --------------------------------------------------------------------------------
// warp per 1024-position tile; per round of 128 positions a lane holds 4 consecutive depths (one int4 load: the warp
// reads 512 contiguous bytes), compares each with its predecessor (the lane below, or the carry of the previous
// round) and the nibbles of 8 lanes are packed into one 32-bit word with 3 shuffles
__global__ void __launch_bounds__(256)
gz_mark_kernel(GzCall call, const int32_t* __restrict__ depth, uint32_t* __restrict__ bits, int32_t* __restrict__ tile_cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= call.n_tiles) return;
  const GzSeg s = call.seg[gz_find_seg_by_tile(call, w)];
  const int64_t p0 = ((s.first >> 10) + (w - s.tile_off)) << 10;        // contig-relative first position of the tile
  const int32_t* d = depth + s.g0 + p0;
  int carry = p0 > 0 ? d[-1] : 0;                                       // depth of the position before the round
  int count = 0;
  // a tile inside the range (all but the first / last tile of a contig's range): a run starts where the value
  // changes, plus at the tile's first position when a member or the range starts there
  const bool inner = p0 >= s.first && p0 + GCI_TILE <= s.end;
  const bool forced0 = p0 == s.first || (p0 & (GZ_MEMBER - 1)) == 0;
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int4 v = *reinterpret_cast<const int4*>(d + r * 128 + lane * 4);
    int prev = __shfl_up_sync(0xffffffffu, v.w, 1);
    if (lane == 0) prev = carry;
    carry = __shfl_sync(0xffffffffu, v.w, 31);
    uint32_t nib;
    if (inner) {
      nib = (v.x != prev ? 1u : 0u) | (v.y != v.x ? 2u : 0u) | (v.z != v.y ? 4u : 0u) | (v.w != v.z ? 8u : 0u);
      if (r == 0 && lane == 0 && forced0) nib |= 1u;
    } else {
      const int64_t p = p0 + r * 128 + lane * 4;
      nib = 0;
      const int x[5] = {prev, v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int64_t q = p + k;
        const bool in = q >= s.first && q < s.end;
        const bool start = q == s.first || (q & (GZ_MEMBER - 1)) == 0 || x[k + 1] != x[k];
        nib |= (in && start ? 1u : 0u) << k;
      }
    }
    uint32_t word = nib << ((lane & 7) * 4);
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7) == 0) {
      bits[w * 32 + r * 4 + (lane >> 3)] = word;
      count += __popc(word);
    }
  }
  count = warp_sum(count);
  if (lane == 0) tile_cnt[w] = count;
}

// ---- pass 3: (position, value) per run ------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gz_runs_kernel(GzCall call, const int32_t* __restrict__ depth, const uint32_t* __restrict__ bits,
               const int32_t* __restrict__ tile_run, int32_t* __restrict__ run_pos, int32_t* __restrict__ run_val) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= call.n_tiles) return;
  uint32_t word = bits[w * 32 + lane];
  const int c = __popc(word);
  if (__all_sync(0xffffffffu, c == 0)) return;
  const GzSeg s = call.seg[gz_find_seg_by_tile(call, w)];
  const int64_t p0 = ((s.first >> 10) + (w - s.tile_off)) << 10;
  int idx = tile_run[w] + warp_incl_scan(c, lane) - c;
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    const int64_t p = p0 + lane * 32 + b;
    run_pos[idx] = (int32_t)p;
    run_val[idx] = depth[s.g0 + p];
    idx++;
  }
}

// ---- passes 4 and 6: warp per member ---------------------------------------------------------------------------
struct DevSink {
  uint32_t* w;
  __device__ __forceinline__ void or_word(uint64_t i, uint32_t v) { atomicOr(w + i, v); }
};

__device__ __forceinline__ void or_byte(uint32_t* out, int64_t byte, uint32_t v) {
  if (v) atomicOr(out + (byte >> 2), v << (8 * (int)(byte & 3)));
}

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

template <bool EMIT>
__global__ void __launch_bounds__(256)
gz_member_kernel(GzCall call, const int32_t* __restrict__ tile_run, const int32_t* __restrict__ run_pos,
                 const int32_t* __restrict__ run_val, const uint8_t* __restrict__ headers, GzTables tab,
                 int32_t* __restrict__ msize, const int64_t* __restrict__ moff, uint32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= call.n_members) return;
  const GzSeg s = call.seg[gz_find_seg_by_member(call, m)];
  const bool first_member = m == s.member_off;
  const int64_t j = (s.first >> 13) + (m - s.member_off);               // member on the contig's absolute grid
  static_assert(GZ_MEMBER == 8192, "member shift");
  int64_t a = max(s.first, j << 13), b = min(s.end, (j + 1) << 13);
  if (s.end <= s.first) a = b = s.first;
  int r0 = 0, r1 = 0;
  if (b > a) {
    r0 = tile_run[s.tile_off + ((a >> 10) - (s.first >> 10))];
    r1 = tile_run[s.tile_off + (((b + 1023) >> 10) - (s.first >> 10))];
  }
  uint32_t bits_total = 3u + (first_member ? (uint32_t)s.hdr_bits : 0u);
  uint32_t text_total = first_member ? (uint32_t)s.hdr_len : 0u;
  uint32_t R = first_member ? s.hdr_crc : 0u;
  DevSink sink{out};
  int64_t byte0 = 0;
  if (EMIT) {
    byte0 = moff[m];
    if (lane < 10) {
      const uint32_t hdr = lane == 0 ? 0x1fu : lane == 1 ? 0x8bu : lane == 2 ? 8u : lane == 9 ? 0xffu : 0u;
      or_byte(out, byte0 + lane, hdr);                                  // magic, deflate, no flags, mtime 0, OS unknown
    }
    if (lane == 0) {
      GzBitPut<DevSink> w(sink, (uint64_t)(byte0 + 10) * 8);
      w.put(3, 3);                                                      // BFINAL = 1, BTYPE = 01 (fixed Huffman)
      if (first_member)
        for (int i = 0; i < s.hdr_len; i++) gz_put_literal(w, headers[s.hdr_off + i]);
      w.flush();
    }
  }
#pragma unroll 1
  for (int base = r0; base < r1; base += 32) {
    const int r = base + lane;
    const bool act = r < r1;
    uint32_t rb = 0, tb = 0, k = 0;
    int Lb = 0, v = 0;
    uint64_t lo = 0;
    uint32_t hi = 0;
    if (act) {
      const int pos = run_pos[r];
      v = run_val[r];
      const int64_t nxt = r + 1 < r1 ? (int64_t)run_pos[r + 1] : b;
      k = (uint32_t)(nxt - pos);
      Lb = gz_format_line(v, &lo, &hi);
      rb = (uint32_t)gz_run_bits(Lb, k);
      tb = (uint32_t)Lb * k;
    }
    const uint32_t ib = warp_incl_scan_u32(rb, lane);
    const uint32_t batch_bits = __shfl_sync(0xffffffffu, ib, 31);
    if (EMIT) {
      const uint32_t it = warp_incl_scan_u32(tb, lane);
      const uint32_t batch_text = __shfl_sync(0xffffffffu, it, 31);
      uint32_t Ri = 0;
      if (act) {
        GzBitPut<DevSink> w(sink, (uint64_t)(byte0 + 10) * 8 + bits_total + (ib - rb));
        gz_put_run(w, lo, hi, Lb, k);
        w.flush();
        Ri = gz_crc_run(tab, v, lo, hi, Lb, k);
        const uint32_t after = batch_text - it;                         // text bytes of the batch behind this run
        if (after) Ri = gz_mulmod(Ri, gz_pow8(tab, after));
      }
      Ri = __reduce_xor_sync(0xffffffffu, Ri);
      R = (R ? gz_mulmod(R, gz_pow8(tab, batch_text)) : 0u) ^ Ri;
      text_total += batch_text;
    }
    bits_total += batch_bits;
  }
  bits_total += 7;                                                       // end of block: seven zero bits
  const int32_t bytes = 10 + (int32_t)((bits_total + 7) >> 3) + 8;
  if (!EMIT) {
    if (lane == 0) msize[m] = bytes;
    return;
  }
  if (lane < 8) {
    const uint32_t crc = ~(gz_mulmod(0xffffffffu, gz_pow8(tab, text_total)) ^ R);
    const uint32_t word = lane < 4 ? crc : text_total;                  // CRC-32, ISIZE (little endian)
    or_byte(out, byte0 + bytes - 8 + lane, (word >> (8 * (lane & 3))) & 0xffu);
  }
}

// ---- host driver --------------------------------------------------------------------------------------------
struct GzRange { int32_t contig; int64_t first, count; const char* header; int32_t header_len; };

static std::vector<uint32_t>& gz_host_tables() {
  static std::vector<uint32_t> flat;
  if (flat.empty()) {
    flat.resize(GZ_TABLE_WORDS);
    gz_build_tables(flat.data());
  }
  return flat;
}

// encode the ranges (in order) into ctx->gz_packed; *total = bytes; range_bytes[n + 1] (optional) = byte offset of
// every range in the packed stream
static int gz_encode(gci_ctx* ctx, Track& t, const std::vector<GzRange>& ranges, int64_t* total, int64_t* range_bytes) {
  const std::vector<uint32_t>& flat = gz_host_tables();
  if (ctx->gz_tables.cap == 0) {
    GCI_TRY(gci_h2d(ctx, ctx->gz_tables, flat.data(), sizeof(uint32_t) * flat.size()));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  const GzTables host_tab = gz_tables_view(flat.data());
  const GzTables tab = gz_tables_view(ctx->gz_tables.as<uint32_t>());
  std::vector<GzSeg> segs(ranges.size());
  std::vector<uint8_t> hdr;
  int64_t n_tiles = 0, n_members = 0;
  for (size_t i = 0; i < ranges.size(); i++) {
    const GzRange& r = ranges[i];
    GzSeg& s = segs[i];
    s.g0 = ctx->pos_off[r.contig];
    s.first = r.first;
    s.end = r.first + r.count;
    s.tile_off = n_tiles;
    s.member_off = n_members;
    s.hdr_off = (int32_t)hdr.size();
    s.hdr_len = r.header_len;
    s.hdr_bits = 0;
    s.hdr_crc = 0;
    for (int k = 0; k < r.header_len; k++) {
      const uint8_t c = (uint8_t)r.header[k];
      hdr.push_back(c);
      s.hdr_bits += c < 144 ? 8 : 9;
      s.hdr_crc = host_tab.crc_byte[(s.hdr_crc ^ c) & 0xffu] ^ (s.hdr_crc >> 8);
    }
    if (r.count > 0) {
      n_tiles += ((s.end - 1) >> 10) - (s.first >> 10) + 1;
      n_members += ((s.end - 1) >> 13) - (s.first >> 13) + 1;
    } else {
      n_members += 1;                                                   // header-only member
    }
  }
  if (n_tiles >= (int64_t(1) << 31) / 32 * 32 || n_members >= (int64_t(1) << 31))
    return ctx->fail(GCI_E_ARG, "gzip: range too large for one call");
  ctx->stage_begin(GCI_ST_TEXT);
  GCI_TRY(gci_h2d(ctx, ctx->gz_seg, segs.data(), sizeof(GzSeg) * segs.size()));
  GCI_TRY(gci_h2d(ctx, ctx->gz_hdr, hdr.data(), hdr.size()));
  GCI_CUDA_TRY(ctx, cudaEventRecord(ctx->h2d_done, ctx->stream));       // segs / hdr are locals
  GzCall call{ctx->gz_seg.as<GzSeg>(), (int32_t)segs.size(), n_tiles, n_members};
  GCI_TRY(ctx->ensure(ctx->gz_bits, 128 * (size_t)std::max<int64_t>(1, n_tiles)));
  GCI_TRY(ctx->ensure(ctx->gz_tile_cnt, 4 * (size_t)(n_tiles + 1)));
  GCI_TRY(ctx->ensure(ctx->gz_tile_run, 4 * (size_t)(n_tiles + 1)));
  GCI_TRY(ctx->ensure(ctx->gz_msize, 4 * (size_t)(n_members + 1)));
  GCI_TRY(ctx->ensure(ctx->gz_moff, 8 * (size_t)(n_members + 2)));
  int32_t* tile_cnt = ctx->gz_tile_cnt.as<int32_t>();
  int32_t* tile_run = ctx->gz_tile_run.as<int32_t>();
  int64_t* h_pin = (int64_t*)ctx->pinned(16);
  if (!h_pin) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  int64_t n_runs = 0;
  if (n_tiles) {
    gz_mark_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, ctx->stream>>>(call, t.depth.as<int32_t>(),
                                                                          ctx->gz_bits.as<uint32_t>(), tile_cnt);
    GCI_LAUNCH_CHECK(ctx);
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(tile_cnt + n_tiles, 0, 4, ctx->stream));
    GCI_TRY(gci_exclusive_scan_i32(ctx, tile_cnt, tile_run, n_tiles + 1));
    int32_t* h32 = (int32_t*)h_pin;
    GCI_TRY(gci_d2h(ctx, h32, tile_run + n_tiles, 4));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    n_runs = *h32;
    if (n_runs < 0) return ctx->fail(GCI_E_ARG, "gzip: more than 2^31 depth runs in one call");
    GCI_TRY(ctx->ensure(ctx->gz_run_pos, 4 * (size_t)std::max<int64_t>(1, n_runs)));
    GCI_TRY(ctx->ensure(ctx->gz_run_val, 4 * (size_t)std::max<int64_t>(1, n_runs)));
    gz_runs_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, ctx->stream>>>(
        call, t.depth.as<int32_t>(), ctx->gz_bits.as<uint32_t>(), tile_run, ctx->gz_run_pos.as<int32_t>(),
        ctx->gz_run_val.as<int32_t>());
    GCI_LAUNCH_CHECK(ctx);
  } else {
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(tile_run, 0, 4, ctx->stream));
    GCI_TRY(ctx->ensure(ctx->gz_run_pos, 16));
    GCI_TRY(ctx->ensure(ctx->gz_run_val, 16));
  }
  const unsigned mgrid = (unsigned)((n_members + 7) / 8);
  gz_member_kernel<false><<<mgrid, 256, 0, ctx->stream>>>(call, tile_run, ctx->gz_run_pos.as<int32_t>(),
                                                          ctx->gz_run_val.as<int32_t>(), ctx->gz_hdr.as<uint8_t>(), tab,
                                                          ctx->gz_msize.as<int32_t>(), nullptr, nullptr);
  GCI_LAUNCH_CHECK(ctx);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->gz_msize.as<int32_t>() + n_members, 0, 4, ctx->stream));
  GCI_TRY(gci_exclusive_scan_i64_from_i32(ctx, ctx->gz_msize.as<int32_t>(), ctx->gz_moff.as<int64_t>(), n_members + 1,
                                          nullptr));
  GCI_TRY(gci_d2h(ctx, h_pin, ctx->gz_moff.as<int64_t>() + n_members, 8));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *total = *h_pin;
  const size_t padded = ((size_t)*total + 3 + 8) & ~size_t(3);
  GCI_TRY(ctx->ensure(ctx->gz_packed, padded));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->gz_packed.p, 0, padded, ctx->stream));
  gz_member_kernel<true><<<mgrid, 256, 0, ctx->stream>>>(call, tile_run, ctx->gz_run_pos.as<int32_t>(),
                                                         ctx->gz_run_val.as<int32_t>(), ctx->gz_hdr.as<uint8_t>(), tab,
                                                         nullptr, ctx->gz_moff.as<int64_t>(),
                                                         ctx->gz_packed.as<uint32_t>());
  GCI_LAUNCH_CHECK(ctx);
  if (range_bytes) {
    // byte offset of every range = offset of its first member
    std::vector<int64_t> idx(ranges.size() + 1);
    for (size_t i = 0; i < ranges.size(); i++) idx[i] = segs[i].member_off;
    idx[ranges.size()] = n_members;
    int64_t* h = (int64_t*)ctx->pinned(8 * (ranges.size() + 1));
    if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
    for (size_t i = 0; i <= ranges.size(); i++)
      GCI_TRY(gci_d2h(ctx, h + i, ctx->gz_moff.as<int64_t>() + idx[i], 8));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(range_bytes, h, 8 * (ranges.size() + 1));
  }
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaEventSynchronize(ctx->h2d_done));
  return GCI_OK;
}

static unsigned long long gz_key_of(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count,
                                    const char* header, int32_t header_len) {
  unsigned long long key = 1469598103934665603ull;
  auto mix = [&key](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
  mix((unsigned long long)track); mix((unsigned long long)contig); mix((unsigned long long)first);
  mix((unsigned long long)count); mix(ctx->epoch);
  for (int k = 0; k < header_len; k++) mix((unsigned char)header[k]);
  return key | 1ull;
}

extern "C" {

int gci_depth_gzip(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count, const char* header,
                   int32_t header_len, char* out, int64_t cap, int64_t* n_bytes) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_bytes || header_len < 0 || (header_len && !header))
    return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_gzip: track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || first < 0 || count < 0 ||
      first + count > ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_depth_gzip: bad range");
  if (header_len > 4096) return ctx->fail(GCI_E_ARG, "gci_depth_gzip: header longer than 4096 bytes");
  *n_bytes = 0;
  // the size query (out == NULL) and the fetch of the same range share one encoding pass; the key covers the
  // context's epoch, so anything that may have changed the track in between invalidates the packed bytes
  const unsigned long long key = gz_key_of(ctx, track, contig, first, count, header, header_len);
  if (!(ctx->gz_valid && ctx->gz_key == key)) {
    ctx->gz_valid = false;
    std::vector<GzRange> one{{contig, first, count, header, header_len}};
    GCI_TRY(gz_encode(ctx, t, one, &ctx->gz_total, nullptr));
    ctx->gz_key = key;
    ctx->gz_valid = true;
  }
  *n_bytes = ctx->gz_total;
  if (!out) return GCI_OK;
  if (cap < ctx->gz_total)
    return ctx->fail(GCI_E_ARG, "gzip buffer too small (%lld < %lld)", (long long)cap, (long long)ctx->gz_total);
  ctx->stage_begin(GCI_ST_D2H);
  GCI_TRY(gci_d2h(ctx, out, ctx->gz_packed.p, (size_t)ctx->gz_total));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->gz_valid = false;
  return GCI_OK;
}

// every selected contig of the track in one pass: header i (bytes [header_off[i], header_off[i + 1]) of `headers`,
// e.g. ">name\n") followed by the contig's depth lines
int gci_depth_gzip_track(gci_ctx* ctx, int32_t track, const char* headers, const int64_t* header_off, char* out,
                         int64_t cap, int64_t* n_bytes, int64_t* contig_bytes) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_bytes || !header_off) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_gzip_track: track %d holds no depth", track);
  std::vector<GzRange> ranges;
  for (int c = 0; c < ctx->n_contigs; c++) {
    if (!ctx->selected[c]) continue;
    const int64_t hl = header_off[c + 1] - header_off[c];
    if (hl < 0 || hl > 4096 || (hl && !headers)) return ctx->fail(GCI_E_ARG, "gci_depth_gzip_track: bad header %d", c);
    ranges.push_back({c, 0, ctx->len[c], headers ? headers + header_off[c] : nullptr, (int32_t)hl});
  }
  *n_bytes = 0;
  std::vector<int64_t> rb(ranges.size() + 1, 0);
  int64_t total = 0;
  ctx->gz_valid = false;
  if (!ranges.empty()) GCI_TRY(gz_encode(ctx, t, ranges, &total, contig_bytes ? rb.data() : nullptr));
  *n_bytes = total;
  if (contig_bytes) {
    size_t k = 0;
    for (int c = 0; c < ctx->n_contigs; c++) {
      contig_bytes[c] = rb[k];
      if (ctx->selected[c]) k++;
    }
    contig_bytes[ctx->n_contigs] = total;
  }
  if (!out) return GCI_OK;
  if (cap < total) return ctx->fail(GCI_E_ARG, "gzip buffer too small (%lld < %lld)", (long long)cap, (long long)total);
  ctx->stage_begin(GCI_ST_D2H);
  GCI_TRY(gci_d2h(ctx, out, ctx->gz_packed.p, (size_t)total));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

}  // extern "C"
