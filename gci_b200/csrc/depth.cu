// Depth stage: survivors -> per-tile event buckets -> per-base depth tiles (+ fused issue flags),
// N-run masking, two-type max, stand-alone flags, depth text.
// Reference: GCI.py:302-306 (accumulate), :315-329 (mask), :332-353 (max), :110-117 (text).
#include <algorithm>

#include "common.cuh"

int gci_exclusive_scan_u64(gci_ctx* ctx, const unsigned long long* in, unsigned long long* out, int64_t n,
                           unsigned long long* total_dev);

// ================================================================================================
// K5  event buckets
// ================================================================================================
// Every survivor contributes +1 at a = start+fl and -1 at b = end-fl+1 (both normalised like a Python
// slice, GCI.py:304-306).  Instead of a genome-sized delta array (zero-fill + scatter + scan = 12 B per
// base) the events are bucketed per depth tile: tile_pack[t] accumulates (count, net) in one 64-bit
// atomic, an exclusive scan of tile_pack gives every tile its event offset (low word) and the depth
// carried into it (high word: sum of all earlier nets; a contig's nets sum to zero, so the carry
// restarts at 0 on every contig without any segmentation).

struct Slice { long long a, b; int64_t tile_a, tile_b; bool ok; };

__device__ __forceinline__ Slice survivor_slice(int32_t c, int32_t s, int32_t e, int32_t fl,
                                                const int64_t* __restrict__ len, const int64_t* __restrict__ tile_off) {
  Slice r;
  r.ok = false;
  if (c < 0) return r;
  const long long L = len[c];
  r.a = py_slice_index((long long)s + fl, L);
  r.b = py_slice_index((long long)e - fl + 1, L);
  if (r.a >= r.b) return r;
  const int64_t t0 = tile_off[c];
  if (tile_off[c + 1] == t0) return r;   // contig without depth storage (not selected / not owned)
  r.ok = true;
  r.tile_a = t0 + r.a / GCI_TILE;
  r.tile_b = t0 + r.b / GCI_TILE;
  return r;
}

constexpr unsigned long long EV_PLUS = 1ull + (1ull << 32);            // count+1, net+1
constexpr unsigned long long EV_MINUS = 1ull + 0xffffffff00000000ull;  // count+1, net-1

__global__ void bucket_count_kernel(uint32_t n_reads, const int32_t* __restrict__ sc, const int32_t* __restrict__ ss,
                                    const int32_t* __restrict__ se, int32_t fl, const int64_t* __restrict__ len,
                                    const int64_t* __restrict__ tile_off, unsigned long long* __restrict__ tile_pack,
                                    long long* __restrict__ sums) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t c = -1;
  long long covered = 0;
  if (r < n_reads) {
    c = sc[r];
    const Slice sl = survivor_slice(c, ss[r], se[r], fl, len, tile_off);
    if (sl.ok) {
      atomicAdd(tile_pack + sl.tile_a, EV_PLUS);
      atomicAdd(tile_pack + sl.tile_b, EV_MINUS);
      covered = sl.b - sl.a;
    } else {
      c = -1;
    }
  }
  // sum of depth per contig = sum of slice lengths; aggregate per warp when the warp agrees on a contig
  const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
  if (act == 0) return;
  const int leader = __ffs(act) - 1;
  const int32_t c0 = __shfl_sync(0xffffffffu, c, leader);
  const bool uniform = __all_sync(0xffffffffu, c < 0 || c == c0);
  if (uniform) {
    const long long t = warp_sum_ll(covered);
    if ((threadIdx.x & 31) == leader) atomicAdd((unsigned long long*)(sums + c0), (unsigned long long)t);
  } else if (c >= 0) {
    atomicAdd((unsigned long long*)(sums + c), (unsigned long long)covered);
  }
}

__global__ void bucket_fill_kernel(uint32_t n_reads, const int32_t* __restrict__ sc, const int32_t* __restrict__ ss,
                                   const int32_t* __restrict__ se, int32_t fl, const int64_t* __restrict__ len,
                                   const int64_t* __restrict__ tile_off,
                                   const unsigned long long* __restrict__ tile_scan, uint32_t* __restrict__ cursor,
                                   uint16_t* __restrict__ events) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const Slice sl = survivor_slice(sc[r], ss[r], se[r], fl, len, tile_off);
  if (!sl.ok) return;
  {
    const uint32_t base = (uint32_t)(tile_scan[sl.tile_a] & 0xffffffffull);
    const uint32_t k = atomicAdd(cursor + sl.tile_a, 1u);
    events[base + k] = (uint16_t)(((uint32_t)(sl.a % GCI_TILE) << 1) | 0u);
  }
  {
    const uint32_t base = (uint32_t)(tile_scan[sl.tile_b] & 0xffffffffull);
    const uint32_t k = atomicAdd(cursor + sl.tile_b, 1u);
    events[base + k] = (uint16_t)(((uint32_t)(sl.b % GCI_TILE) << 1) | 1u);
  }
}

// ================================================================================================
// K6  depth tiles (+ fused flags)
// ================================================================================================
// One CTA per tile of 8192 positions.  The tile's events are scattered into a shared-memory delta
// array, each warp scans a contiguous 1024-position slice (lane-consecutive int4: conflict-free LDS.128
// and fully coalesced 512-byte STG.128 per warp instruction), the carry into the tile comes from the
// tile scan, so there is no dependency between CTAs: the kernel is a pure streaming write of 4 B per
// base (+ 1 bit per base of flags).

// pack per-lane 4-bit nibbles (positions 4*lane .. 4*lane+3 of a 128-position warp row) into four
// 32-bit words held by lanes 0, 8, 16, 24
__device__ __forceinline__ uint32_t pack_nibbles(uint32_t nib, int lane) {
  uint32_t w = nib << ((lane & 7) * 4);
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  w |= __shfl_xor_sync(0xffffffffu, w, 4);
  return w;
}

// shared-memory layout of the delta tile: 4 pad words after every 32 positions, so that a lane's 8
// consecutive positions (two LDS.128) never collide with the other lanes of its quarter-warp
__device__ __forceinline__ int pad_idx(int p) { return p + ((p >> 5) << 2); }
constexpr int GCI_TILE_WORDS = GCI_TILE + GCI_TILE / 8;      // 9216 words = 36 KB

// 8 flag bits of one lane (positions idx..idx+7) -> 32-bit words held by lanes 0, 4, 8, ...
__device__ __forceinline__ uint32_t pack_bytes(uint32_t byte, int lane) {
  uint32_t w = byte << ((lane & 3) * 8);
  w |= __shfl_xor_sync(0xffffffffu, w, 1);
  w |= __shfl_xor_sync(0xffffffffu, w, 2);
  return w;
}

__device__ __forceinline__ uint32_t flag8(const int (&o)[8], int lo1, uint32_t span) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) m |= ((uint32_t)(o[k] - lo1) < span ? 1u : 0u) << k;
  return m;
}

// FLAGS: also emit the issue bit (lo < depth <= hi) of every position.
// lo1 = lo + 1, span = number of admissible depth values (0 = none).
template <bool FLAGS>
__global__ void __launch_bounds__(GCI_TILE_THREADS)
depth_tile_kernel(const unsigned long long* __restrict__ tile_pack, const unsigned long long* __restrict__ tile_scan,
                  const uint16_t* __restrict__ events, const int64_t* __restrict__ tile_off,
                  const int64_t* __restrict__ len, int32_t n_contigs, int32_t* __restrict__ depth,
                  uint32_t* __restrict__ flags, int32_t lo1, uint32_t span) {
  __shared__ __align__(16) int s_delta[GCI_TILE_WORDS];
  __shared__ int s_part[GCI_TILE_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t tile = blockIdx.x;
  constexpr int PER_WARP = GCI_TILE / (GCI_TILE_THREADS / 32);   // 1024
  constexpr int ITERS = PER_WARP / 256;                          // 4

#pragma unroll
  for (int v = tid; v < GCI_TILE_WORDS / 4; v += GCI_TILE_THREADS)
    reinterpret_cast<int4*>(s_delta)[v] = make_int4(0, 0, 0, 0);
  if (tid < GCI_TILE_THREADS / 32) s_part[tid] = 0;
  __syncthreads();
  const unsigned long long pk = tile_pack[tile], sc = tile_scan[tile];
  const uint32_t n_ev = (uint32_t)(pk & 0xffffffffull);
  const uint32_t ev0 = (uint32_t)(sc & 0xffffffffull);
  const int base = (int)(uint32_t)(sc >> 32);                    // depth carried into the tile
  for (uint32_t i = tid; i < n_ev; i += GCI_TILE_THREADS) {
    const uint32_t e = events[ev0 + i];
    const int d = (e & 1u) ? -1 : 1;
    const int p = (int)(e >> 1);
    atomicAdd(&s_delta[pad_idx(p)], d);
    atomicAdd(&s_part[p / PER_WARP], d);
  }
  __syncthreads();

  // contig of this tile; `valid` = positions of the tile that are real bases (the rest is padding)
  const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
  const int64_t left = len[c] - (tile - tile_off[c]) * GCI_TILE;
  const int valid = left >= GCI_TILE ? GCI_TILE : (int)left;
  int carry = base;
  for (int j = 0; j < warp; j++) carry += s_part[j];
  int32_t* __restrict__ out = depth + tile * GCI_TILE;
  uint32_t* __restrict__ fout = flags + tile * (GCI_TILE / 32);
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int idx = warp * PER_WARP + it * 256 + lane * 8;
    const int w = pad_idx(idx);
    const int4 v0 = *reinterpret_cast<const int4*>(&s_delta[w]);
    const int4 v1 = *reinterpret_cast<const int4*>(&s_delta[w + 4]);
    int o[8];
    o[0] = v0.x; o[1] = o[0] + v0.y; o[2] = o[1] + v0.z; o[3] = o[2] + v0.w;
    o[4] = o[3] + v1.x; o[5] = o[4] + v1.y; o[6] = o[5] + v1.z; o[7] = o[6] + v1.w;
    const int incl = warp_incl_scan(o[7], lane);
    const int run = carry + incl - o[7];
    carry += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] += run;
    if (idx + 8 > valid) {   // only the last tile of a contig: padding behind the last base stays 0
#pragma unroll
      for (int k = 0; k < 8; k++) if (idx + k >= valid) o[k] = 0;
    }
    *reinterpret_cast<int4*>(out + idx) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4*>(out + idx + 4) = make_int4(o[4], o[5], o[6], o[7]);
    if (FLAGS) {
      uint32_t m = flag8(o, lo1, span);
      if (idx + 8 > valid) m &= idx >= valid ? 0u : (0xffu >> (idx + 8 - valid));
      const uint32_t word = pack_bytes(m, lane);
      if ((lane & 3) == 0) fout[idx >> 5] = word;
    }
  }
}

// ================================================================================================
// stand-alone flags (resume path / different thresholds), two-type max, N-run mask
// ================================================================================================
__global__ void __launch_bounds__(GCI_TILE_THREADS)
flags_kernel(const int32_t* __restrict__ depth, const int64_t* __restrict__ tile_off, const int64_t* __restrict__ len,
             int32_t n_contigs, uint32_t* __restrict__ flags, int32_t lo, int32_t hi) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t tile = blockIdx.x;
  const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
  const int64_t cpos0 = (tile - tile_off[c]) * GCI_TILE;
  const int64_t L = len[c];
  const int64_t gbase = tile * GCI_TILE;
#pragma unroll
  for (int it = 0; it < GCI_TILE / (GCI_TILE_THREADS * 4); it++) {
    const int idx = it * GCI_TILE_THREADS * 4 + tid * 4;
    const int4 m = *reinterpret_cast<const int4*>(&depth[gbase + idx]);
    const int64_t cp = cpos0 + idx;
    uint32_t nib = 0;
    nib |= (m.x > lo && m.x <= hi && cp < L) ? 1u : 0u;
    nib |= (m.y > lo && m.y <= hi && cp + 1 < L) ? 2u : 0u;
    nib |= (m.z > lo && m.z <= hi && cp + 2 < L) ? 4u : 0u;
    nib |= (m.w > lo && m.w <= hi && cp + 3 < L) ? 8u : 0u;
    const uint32_t w = pack_nibbles(nib, lane);
    if ((lane & 7) == 0) flags[(gbase + idx) >> 5] = w;
  }
}

__global__ void __launch_bounds__(GCI_TILE_THREADS)
max_kernel(const int32_t* __restrict__ da, const int32_t* __restrict__ db, int32_t* __restrict__ dout,
           const int64_t* __restrict__ tile_off, const int64_t* __restrict__ len, int32_t n_contigs,
           uint32_t* __restrict__ flags, int32_t lo, int32_t hi, long long* __restrict__ sums) {
  __shared__ long long s_red[GCI_TILE_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t tile = blockIdx.x;
  const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
  const int64_t cpos0 = (tile - tile_off[c]) * GCI_TILE;
  const int64_t L = len[c];
  const int64_t gbase = tile * GCI_TILE;
  long long acc = 0;
#pragma unroll
  for (int it = 0; it < GCI_TILE / (GCI_TILE_THREADS * 4); it++) {
    const int idx = it * GCI_TILE_THREADS * 4 + tid * 4;
    const int4 a = *reinterpret_cast<const int4*>(&da[gbase + idx]);
    const int4 b = *reinterpret_cast<const int4*>(&db[gbase + idx]);
    const int4 m = make_int4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
    *reinterpret_cast<int4*>(&dout[gbase + idx]) = m;
    acc += (long long)m.x + m.y + m.z + m.w;      // padding is 0 in both inputs
    const int64_t cp = cpos0 + idx;
    uint32_t nib = 0;
    nib |= (m.x > lo && m.x <= hi && cp < L) ? 1u : 0u;
    nib |= (m.y > lo && m.y <= hi && cp + 1 < L) ? 2u : 0u;
    nib |= (m.z > lo && m.z <= hi && cp + 2 < L) ? 4u : 0u;
    nib |= (m.w > lo && m.w <= hi && cp + 3 < L) ? 8u : 0u;
    const uint32_t w = pack_nibbles(nib, lane);
    if ((lane & 7) == 0) flags[(gbase + idx) >> 5] = w;
  }
  acc = warp_sum_ll(acc);
  if (lane == 0) s_red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int j = 0; j < GCI_TILE_THREADS / 32; j++) t += s_red[j];
    if (t) atomicAdd((unsigned long long*)(sums + c), (unsigned long long)t);
  }
}

// per-contig sum of a loaded track (resume path)
__global__ void __launch_bounds__(GCI_TILE_THREADS)
sum_kernel(const int32_t* __restrict__ depth, const int64_t* __restrict__ tile_off, int32_t n_contigs,
           long long* __restrict__ sums) {
  __shared__ long long s_red[GCI_TILE_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t tile = blockIdx.x;
  const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
  const int64_t gbase = tile * GCI_TILE;
  long long acc = 0;
#pragma unroll
  for (int it = 0; it < GCI_TILE / (GCI_TILE_THREADS * 4); it++) {
    const int4 m = *reinterpret_cast<const int4*>(&depth[gbase + it * GCI_TILE_THREADS * 4 + tid * 4]);
    acc += (long long)m.x + m.y + m.z + m.w;
  }
  acc = warp_sum_ll(acc);
  if (lane == 0) s_red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    long long t = 0;
    for (int j = 0; j < GCI_TILE_THREADS / 32; j++) t += s_red[j];
    if (t) atomicAdd((unsigned long long*)(sums + c), (unsigned long long)t);
  }
}

struct NRuns {
  const int32_t* contig;
  const int64_t* start;
  const int64_t* end;
  int64_t n;
};

// depths[target][s:e] = 0 (GCI.py:328): one CTA per N-run; keeps sums and flags consistent
__global__ void mask_kernel(NRuns nr, const int64_t* __restrict__ tile_off, int32_t* __restrict__ depth,
                            uint32_t* __restrict__ flags, int flags_valid, int32_t lo, int32_t hi,
                            long long* __restrict__ sums) {
  __shared__ long long s_red[8];
  const int64_t k = blockIdx.x;
  const int32_t c = nr.contig[k];
  const int64_t g0 = tile_off[c] * GCI_TILE;
  const bool zero_is_issue = (0 > lo && 0 <= hi);
  long long removed = 0;
  for (int64_t p = nr.start[k] + threadIdx.x; p < nr.end[k]; p += blockDim.x) {
    removed += depth[g0 + p];
    depth[g0 + p] = 0;
    if (flags_valid) {
      const uint32_t bit = 1u << ((g0 + p) & 31);
      if (zero_is_issue) atomicOr(&flags[(g0 + p) >> 5], bit); else atomicAnd(&flags[(g0 + p) >> 5], ~bit);
    }
  }
  removed = warp_sum_ll(removed);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = removed;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int j = 0; j < (int)(blockDim.x >> 5); j++) t += s_red[j];
    if (t) atomicAdd((unsigned long long*)(sums + c), (unsigned long long)(-t));
  }
}

// ================================================================================================
// depth text (GCI.py:115-117): "%d\n" per base, produced on the GPU
// ================================================================================================
__device__ __forceinline__ int dec_len(int v) {
  unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
  int n = 1;
  if (u >= 10u) n = 2;
  if (u >= 100u) n = 3;
  if (u >= 1000u) n = 4;
  if (u >= 10000u) n = 5;
  if (u >= 100000u) n = 6;
  if (u >= 1000000u) n = 7;
  if (u >= 10000000u) n = 8;
  if (u >= 100000000u) n = 9;
  if (u >= 1000000000u) n = 10;
  return n + (v < 0 ? 1 : 0) + 1;   // digits + sign + '\n'
}

__global__ void text_len_kernel(const int32_t* __restrict__ depth, int64_t n, int32_t* __restrict__ out_len) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out_len[i] = dec_len(depth[i]);
}

__global__ void text_write_kernel(const int32_t* __restrict__ depth, int64_t n, const int64_t* __restrict__ off,
                                  char* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = depth[i];
  const int len = dec_len(v);
  char* p = out + off[i];
  unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
  p[len - 1] = '\n';
  int k = len - 2;
  do {
    p[k--] = (char)('0' + u % 10u);
    u /= 10u;
  } while (u);
  if (v < 0) p[0] = '-';
}

// ================================================================================================
// host drivers
// ================================================================================================
static NRuns make_nruns(gci_ctx* ctx) {
  NRuns nr;
  nr.contig = ctx->d_nr_contig.as<int32_t>();
  nr.start = ctx->d_nr_start.as<int64_t>();
  nr.end = ctx->d_nr_end.as<int64_t>();
  nr.n = ctx->n_nruns;
  return nr;
}

int gci_compute_flags(gci_ctx* ctx, int track, int32_t lo, int32_t hi) {
  Track& t = ctx->track[track];
  ctx->stage_begin(GCI_ST_FLAGS);
  if (ctx->n_tiles) {
    flags_kernel<<<(unsigned)ctx->n_tiles, GCI_TILE_THREADS, 0, ctx->stream>>>(
        t.depth.as<int32_t>(), ctx->d_tile_off.as<int64_t>(), ctx->d_len.as<int64_t>(), ctx->n_contigs,
        t.flags.as<uint32_t>(), lo, hi);
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  t.flags_valid = true;
  t.flags_lo = lo;
  t.flags_hi = hi;
  return GCI_OK;
}

extern "C" {

int gci_depth(gci_ctx* ctx, int32_t track, int32_t flank_len, int32_t lo, int32_t hi) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->filtered) return ctx->fail(GCI_E_ARG, "gci_depth before gci_filter");
  Track& t = ctx->track[track];
  if (!t.allocated) {
    // allocate without the zero fill: every position of the track is written by the tile kernel
    if (ctx->n_contigs <= 0) return ctx->fail(GCI_E_ARG, "gci_set_contigs has not been called");
    GCI_TRY(ctx->ensure(t.depth, sizeof(int32_t) * (size_t)ctx->total_padded));
    GCI_TRY(ctx->ensure(t.flags, sizeof(uint32_t) * (size_t)(ctx->total_padded / 32)));
    GCI_TRY(ctx->ensure(t.sums, sizeof(int64_t) * (size_t)ctx->n_contigs));
    t.allocated = true;
  }
  const int64_t nt = ctx->n_tiles;
  const uint32_t nr = ctx->n_reads;
  if (nt == 0) return GCI_OK;
  if (nt >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "too many tiles");
  // one scratch block: [packed (count, net) u64 x nt | fill cursors u32 x nt] zeroed with one memset
  GCI_TRY(ctx->ensure(ctx->tile_cnt, 12 * (size_t)nt));
  GCI_TRY(ctx->ensure(ctx->tile_evoff, 8 * (size_t)nt));   // exclusive scan of the packed values
  GCI_TRY(ctx->ensure(ctx->events, 2 * 2 * (size_t)std::max<uint32_t>(1, nr)));
  unsigned long long* pack = ctx->tile_cnt.as<unsigned long long>();
  uint32_t* cursor = reinterpret_cast<uint32_t*>(pack + nt);
  unsigned long long* scan = ctx->tile_evoff.as<unsigned long long>();
  ctx->stage_begin(GCI_ST_BUCKET);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(pack, 0, 12 * (size_t)nt, ctx->stream));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  if (nr) {
    bucket_count_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(
        nr, ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(), ctx->surv_end.as<int32_t>(), flank_len,
        ctx->d_len.as<int64_t>(), ctx->d_tile_off.as<int64_t>(), pack, t.sums.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  GCI_TRY(gci_exclusive_scan_u64(ctx, pack, scan, nt, nullptr));
  if (nr) {
    bucket_fill_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(
        nr, ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(), ctx->surv_end.as<int32_t>(), flank_len,
        ctx->d_len.as<int64_t>(), ctx->d_tile_off.as<int64_t>(), scan, cursor, ctx->events.as<uint16_t>());
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  const bool fuse = !(lo == INT32_MIN && hi == INT32_MIN);
  const int32_t lo1 = fuse ? (int32_t)((int64_t)lo + 1 > INT32_MAX ? INT32_MAX : lo + 1) : 0;
  const uint32_t span = (fuse && hi > lo) ? (uint32_t)((int64_t)hi - (int64_t)lo) : 0u;
  ctx->stage_begin(GCI_ST_DEPTH);
  if (fuse) {
    depth_tile_kernel<true><<<(unsigned)nt, GCI_TILE_THREADS, 0, ctx->stream>>>(
        pack, scan, ctx->events.as<uint16_t>(), ctx->d_tile_off.as<int64_t>(), ctx->d_len.as<int64_t>(),
        ctx->n_contigs, t.depth.as<int32_t>(), t.flags.as<uint32_t>(), lo1, span);
  } else {
    depth_tile_kernel<false><<<(unsigned)nt, GCI_TILE_THREADS, 0, ctx->stream>>>(
        pack, scan, ctx->events.as<uint16_t>(), ctx->d_tile_off.as<int64_t>(), ctx->d_len.as<int64_t>(),
        ctx->n_contigs, t.depth.as<int32_t>(), t.flags.as<uint32_t>(), lo1, span);
  }
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  // depth and flags are both un-masked here (the reference writes the single-type .depth.gz before
  // masking, GCI.py:310 vs :993); gci_mask_gaps patches both over the N-runs
  t.flags_valid = fuse;
  t.flags_lo = lo;
  t.flags_hi = hi;
  t.n_intervals = 0;
  t.n_owners = 0;
  return GCI_OK;
}

int gci_mask_gaps(gci_ctx* ctx, int32_t track) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_mask_gaps: track %d holds no depth", track);
  if (ctx->n_nruns == 0) return GCI_OK;
  ctx->stage_begin(GCI_ST_MASK);
  mask_kernel<<<(unsigned)ctx->n_nruns, 256, 0, ctx->stream>>>(make_nruns(ctx), ctx->d_tile_off.as<int64_t>(),
                                                               t.depth.as<int32_t>(), t.flags.as<uint32_t>(),
                                                               t.flags_valid ? 1 : 0, t.flags_lo, t.flags_hi,
                                                               t.sums.as<long long>());
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  return GCI_OK;
}

int gci_merge_max(gci_ctx* ctx, int32_t ta, int32_t tb, int32_t tout, int32_t lo, int32_t hi) {
  if (!ctx || ta < 0 || tb < 0 || tout < 0 || ta >= GCI_MAX_TRACKS || tb >= GCI_MAX_TRACKS || tout >= GCI_MAX_TRACKS)
    return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->track[ta].allocated || !ctx->track[tb].allocated)
    return ctx->fail(GCI_E_ARG, "gci_merge_max: input tracks hold no depth");
  Track& t = ctx->track[tout];
  if (!t.allocated) {
    GCI_TRY(ctx->ensure(t.depth, sizeof(int32_t) * (size_t)ctx->total_padded));
    GCI_TRY(ctx->ensure(t.flags, sizeof(uint32_t) * (size_t)(ctx->total_padded / 32)));
    GCI_TRY(ctx->ensure(t.sums, sizeof(int64_t) * (size_t)ctx->n_contigs));
    t.allocated = true;
  }
  ctx->stage_begin(GCI_ST_MAX);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  const int32_t flo = (lo == INT32_MIN && hi == INT32_MIN) ? -1 : lo;
  const int32_t fhi = (lo == INT32_MIN && hi == INT32_MIN) ? 0 : hi;
  if (ctx->n_tiles) {
    max_kernel<<<(unsigned)ctx->n_tiles, GCI_TILE_THREADS, 0, ctx->stream>>>(
        ctx->track[ta].depth.as<int32_t>(), ctx->track[tb].depth.as<int32_t>(), t.depth.as<int32_t>(),
        ctx->d_tile_off.as<int64_t>(), ctx->d_len.as<int64_t>(), ctx->n_contigs, t.flags.as<uint32_t>(), flo, fhi,
        t.sums.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  t.flags_valid = true;
  t.flags_lo = flo;
  t.flags_hi = fhi;
  t.n_intervals = 0;
  t.n_owners = 0;
  return GCI_OK;
}

int gci_depth_sums(gci_ctx* ctx, int32_t track, int64_t* sums) {
  if (!ctx || !sums || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_sums: track %d holds no depth", track);
  // recompute from the depth array: also right after gci_load_depth
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  if (ctx->n_tiles) {
    sum_kernel<<<(unsigned)ctx->n_tiles, GCI_TILE_THREADS, 0, ctx->stream>>>(
        t.depth.as<int32_t>(), ctx->d_tile_off.as<int64_t>(), ctx->n_contigs, t.sums.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  GCI_TRY(gci_d2h(ctx, sums, t.sums.p, sizeof(int64_t) * (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_depth_text(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count, char* out, int64_t cap,
                   int64_t* n_bytes) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_bytes) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_text: track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || first < 0 || count < 0 ||
      first + count > ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_depth_text: bad range");
  *n_bytes = 0;
  if (count == 0) return GCI_OK;
  const int32_t* d = t.depth.as<int32_t>() + ctx->pos_off[contig] + first;
  DevBuf &lens = ctx->tmp[8], &offs = ctx->tmp[9], &text = ctx->tmp[3];
  GCI_TRY(ctx->ensure(lens, 4 * (size_t)count));
  GCI_TRY(ctx->ensure(offs, 8 * (size_t)(count + 1)));
  ctx->stage_begin(GCI_ST_TEXT);
  text_len_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d, count, lens.as<int32_t>());
  GCI_LAUNCH_CHECK(ctx);
  GCI_TRY(gci_exclusive_scan_i64_from_i32(ctx, lens.as<int32_t>(), offs.as<int64_t>(), count,
                                          offs.as<int64_t>() + count));
  int64_t total = 0;
  GCI_TRY(gci_d2h(ctx, &total, offs.as<int64_t>() + count, 8));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *n_bytes = total;
  if (!out) {
    ctx->stage_end();
    return GCI_OK;
  }
  if (cap < total) {
    ctx->stage_end();
    return ctx->fail(GCI_E_ARG, "text buffer too small (%lld < %lld)", (long long)cap, (long long)total);
  }
  GCI_TRY(ctx->ensure(text, (size_t)total));
  text_write_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(d, count, offs.as<int64_t>(),
                                                                              text.as<char>());
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  ctx->stage_begin(GCI_ST_D2H);
  GCI_TRY(gci_d2h(ctx, out, text.p, (size_t)total));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

}  // extern "C"
