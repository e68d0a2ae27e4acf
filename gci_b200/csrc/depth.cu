// Depth stage: survivors -> per-tile event buckets -> per-base depth tiles (+ fused issue flags),
// N-run masking, two-type max, stand-alone flags, depth text.
// Reference: GCI.py:302-306 (accumulate), :315-329 (mask), :332-353 (max), :110-117 (text).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

int gci_scan_tile_pack(gci_ctx* ctx, const uint32_t* cnt_start, const uint32_t* cnt_end, ulonglong2* tile_ps,
                       int64_t n);   // scan.cu: .x = (events, net) of the tile, .y = exclusive scan of .x

// ================================================================================================
// K5  event buckets
// ================================================================================================
// Every survivor contributes +1 at a = start+fl and -1 at b = end-fl+1 (both normalised like a Python
// slice, GCI.py:304-306).  Instead of a genome-sized delta array (zero-fill + scatter + scan = 12 B per
// base) the events are bucketed per depth tile: tile_pack[t] accumulates (count, net) in one 64-bit
// atomic, an exclusive scan of tile_pack gives every tile its event offset (low word) and the depth
// carried into it (high word: sum of all earlier nets; a contig's nets sum to zero, so the carry
// restarts at 0 on every contig without any segmentation).

__global__ void __launch_bounds__(256)
bucket_count_kernel(uint32_t n_reads, const int32_t* __restrict__ sc, const int32_t* __restrict__ ss,
                    const int32_t* __restrict__ se, BucketArgs bk) {
  __shared__ ContigCache cc;
  contig_cache_load(cc, bk);
  WarpSums ws;
  ws.init();
  long long begin, end;
  cta_range(n_reads, blockDim.x, begin, end);
  for (long long base = begin; base < end; base += blockDim.x) {
    const uint32_t r = (uint32_t)(base + threadIdx.x);
    const bool in = r < n_reads;
    ws.add(bk, in ? sc[r] : -1, in ? ss[r] : 0, in ? se[r] : 0, false);
  }
  ws.flush(bk, nullptr);
}

__global__ void bucket_fill_kernel(uint32_t n_reads, const int32_t* __restrict__ sc, const int32_t* __restrict__ ss,
                                   const int32_t* __restrict__ se, int32_t fl, const int64_t* __restrict__ len,
                                   const int64_t* __restrict__ tile_off,
                                   const ulonglong2* __restrict__ tile_ps, uint32_t* __restrict__ cursor,
                                   uint16_t* __restrict__ events, int32_t n_contigs,
                                   const uint32_t* __restrict__ n_dev /* survivor count on the device, or NULL */) {
  __shared__ ContigCache cc;
  const uint32_t limit = n_dev ? *n_dev : n_reads;
  // contiguous survivor range per CTA: the grid is bounded (a sharded read set sizes n_reads for the slot capacity,
  // world x the survivors that arrive), CTAs behind the count leave at once
  long long begin, end;
  cta_range(limit, blockDim.x, begin, end);
  if (begin >= end) return;
  if (n_contigs <= GCI_SMEM_CONTIGS) {
    for (int i = threadIdx.x; i < n_contigs; i += blockDim.x) cc.len[i] = len[i];
    for (int i = threadIdx.x; i <= n_contigs; i += blockDim.x) cc.tile_off[i] = tile_off[i];
    __syncthreads();
    len = cc.len;
    tile_off = cc.tile_off;
  }
  for (long long at = begin; at < end; at += blockDim.x) {
    const long long r = at + threadIdx.x;
    if (r >= end) break;
    const Slice sl = survivor_slice(sc[r], ss[r], se[r], fl, len, tile_off);
    if (!sl.ok) continue;
    {
      const uint32_t base = (uint32_t)(tile_ps[sl.tile_a].y & 0xffffffffull);
      const uint32_t k = atomicAdd(cursor + sl.tile_a, 1u);
      events[base + k] = (uint16_t)(((uint32_t)(sl.a % GCI_TILE) << 1) | 0u);
    }
    {
      const uint32_t base = (uint32_t)(tile_ps[sl.tile_b].y & 0xffffffffull);
      const uint32_t k = atomicAdd(cursor + sl.tile_b, 1u);
      events[base + k] = (uint16_t)(((uint32_t)(sl.b % GCI_TILE) << 1) | 1u);
    }
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

// 8 flag bits of one lane (positions idx..idx+7) -> 32-bit words held by lanes 0, 4, 8, ...
// (layout of the streaming kernels: a lane owns 8 consecutive positions)
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

__device__ __forceinline__ uint32_t flag4(int a, int b, int c, int d, int lo1, uint32_t span) {
  return ((uint32_t)(a - lo1) < span ? 1u : 0u) | ((uint32_t)(b - lo1) < span ? 2u : 0u) |
         ((uint32_t)(c - lo1) < span ? 4u : 0u) | ((uint32_t)(d - lo1) < span ? 8u : 0u);
}

// Issue-flag words of one round (256 positions): lane l holds the depths of positions [4l, 4l+4) (oa) and
// [128+4l, 128+4l+4) (ob).  Almost every round of an assembly holds no issue at all, so the warp first asks whether
// ANY of its 256 values is inside (lo, hi] — an unsigned minimum per lane and one vote — and only builds the bit
// masks (two flag4 + a 3-shuffle butterfly) for the rounds that need them.
__device__ __forceinline__ void emit_flags(uint32_t* __restrict__ flags, int64_t tile, int it, int lane, const int4& oa,
                                           const int4& ob, int32_t lo1, uint32_t span) {
  const uint32_t m = min(min(min((uint32_t)(oa.x - lo1), (uint32_t)(oa.y - lo1)), min((uint32_t)(oa.z - lo1), (uint32_t)(oa.w - lo1))),
                         min(min((uint32_t)(ob.x - lo1), (uint32_t)(ob.y - lo1)), min((uint32_t)(ob.z - lo1), (uint32_t)(ob.w - lo1))));
  const bool odd = lane & 1;
  uint32_t acc = 0;
  if (__any_sync(0xffffffffu, m < span)) {
    // butterfly: even lanes collect the A-half words, odd lanes the B-half words (3 shuffles for 8 words)
    const uint32_t na = flag4(oa.x, oa.y, oa.z, oa.w, lo1, span), nb = flag4(ob.x, ob.y, ob.z, ob.w, lo1, span);
    uint32_t got = __shfl_xor_sync(0xffffffffu, odd ? na : nb, 1);
    acc = odd ? (got | (nb << 4)) : (na | (got << 4));
    got = __shfl_xor_sync(0xffffffffu, acc, 2);
    acc = (lane & 2) ? (got | (acc << 8)) : (acc | (got << 8));
    got = __shfl_xor_sync(0xffffffffu, acc, 4);
    acc = (lane & 4) ? (got | (acc << 16)) : (acc | (got << 16));
  }
  if ((lane & 6) == 0)   // lanes 8g (A word g) and 8g+1 (B word g)
    flags[tile * (GCI_TILE / 32) + it * 8 + (odd ? 4 : 0) + (lane >> 3)] = acc;
}

// exclusive prefix over lanes of a mostly-zero per-lane value: one shuffle per non-zero lane instead of
// the 5 dependent shuffles of a full warp scan (falls back to the scan when more than 4 lanes are set)
__device__ __forceinline__ int sparse_excl_scan(int v, int lane, int& total) {
  unsigned nz = __ballot_sync(0xffffffffu, v != 0);
  int ex = 0;
  total = 0;
  if (__popc(nz) > 4) {
    const int incl = warp_incl_scan(v, lane);
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
  }
  while (nz) {
    const int j = __ffs(nz) - 1;
    nz &= nz - 1;
    const int x = __shfl_sync(0xffffffffu, v, j);
    ex += lane > j ? x : 0;
    total += x;
  }
  return ex;
}

// Persistent warps, one tile (1024 positions) per warp at a time, no block-level barrier.
// Per round of 256 positions a lane owns two quads: positions [4l, 4l+4) and [128+4l, 128+4l+4), so both
// the shared-memory loads (LDS.128) and the global stores (STG.128) are lane-consecutive: conflict-free and
// full 32-byte sectors.  The tile table entry and the events of the NEXT tile are fetched while the current
// one is expanded; the depth carried into a tile comes from the tile scan, so warps never wait for each
// other.  FLAGS: also emit the issue bit (lo < depth <= hi): lo1 = lo + 1, span = # admissible values.
template <bool FLAGS>
__global__ void __launch_bounds__(GCI_TILE_THREADS)
depth_tile_kernel(const ulonglong2* __restrict__ tile_ps /* (pack, scan) per tile */,
                  const uint16_t* __restrict__ events, int64_t n_tiles, int32_t* __restrict__ depth,
                  uint32_t* __restrict__ flags, int32_t lo1, uint32_t span) {
  __shared__ __align__(16) int s_all[(GCI_TILE_THREADS / 32) * GCI_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* __restrict__ s_delta = s_all + warp * GCI_TILE;
  constexpr int ITERS = GCI_TILE / 256;                          // 4

#pragma unroll
  for (int v = lane; v < GCI_TILE / 4; v += 32) reinterpret_cast<int4*>(s_delta)[v] = make_int4(0, 0, 0, 0);
  __syncwarp();

  const int64_t n_warps = (int64_t)gridDim.x * (GCI_TILE_THREADS / 32);
  int64_t tile = (int64_t)blockIdx.x * (GCI_TILE_THREADS / 32) + warp;
  const ulonglong2 zero2 = make_ulonglong2(0, 0);
  ulonglong2 ps = tile < n_tiles ? tile_ps[tile] : zero2;                       // current tile
  ulonglong2 ps1 = tile + n_warps < n_tiles ? tile_ps[tile + n_warps] : zero2;  // next tile
  uint32_t ev = lane < (uint32_t)(ps.x & 0xffffffffull) ? events[(uint32_t)(ps.y & 0xffffffffull) + lane] : 0u;
  for (; tile < n_tiles; tile += n_warps) {
    const uint32_t n_ev = (uint32_t)(ps.x & 0xffffffffull);
    const uint32_t ev0 = (uint32_t)(ps.y & 0xffffffffull);
    int carry = (int)(uint32_t)(ps.y >> 32);                     // depth carried into the tile
    if (lane < n_ev) atomicAdd(&s_delta[ev >> 1], (ev & 1u) ? -1 : 1);
    for (uint32_t i = lane + 32; i < n_ev; i += 32) {            // more than 32 events in the tile: rare
      const uint32_t e = events[ev0 + i];
      atomicAdd(&s_delta[e >> 1], (e & 1u) ? -1 : 1);
    }
    // software pipeline: events of the next tile, table entry of the one after
    ps = ps1;
    ev = lane < (uint32_t)(ps.x & 0xffffffffull) ? events[(uint32_t)(ps.y & 0xffffffffull) + lane] : 0u;
    ps1 = tile + 2 * n_warps < n_tiles ? tile_ps[tile + 2 * n_warps] : zero2;
    __syncwarp();
    int32_t* __restrict__ out = depth + tile * GCI_TILE;
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int ia = it * 256 + lane * 4, ib = ia + 128;
      const int4 va = *reinterpret_cast<const int4*>(&s_delta[ia]);
      const int4 vb = *reinterpret_cast<const int4*>(&s_delta[ib]);
      if (__ballot_sync(0xffffffffu, (va.x | va.y | va.z | va.w | vb.x | vb.y | vb.z | vb.w) != 0) == 0u) {
        // no read starts or ends in these 256 positions: the depth is flat (nothing to re-zero either)
        const int4 o = make_int4(carry, carry, carry, carry);
        *reinterpret_cast<int4*>(out + ia) = o;
        *reinterpret_cast<int4*>(out + ib) = o;
        if (FLAGS && (lane & 6) == 0)
          flags[tile * (GCI_TILE / 32) + it * 8 + ((lane & 1) ? 4 : 0) + (lane >> 3)] =
              (uint32_t)(carry - lo1) < span ? 0xffffffffu : 0u;
        continue;
      }
      const int a0 = va.x, a1 = a0 + va.y, a2 = a1 + va.z, a3 = a2 + va.w;
      const int b0 = vb.x, b1 = b0 + vb.y, b2 = b1 + vb.z, b3 = b2 + vb.w;
      if ((va.x | va.y | va.z | va.w) != 0) *reinterpret_cast<int4*>(&s_delta[ia]) = make_int4(0, 0, 0, 0);
      if ((vb.x | vb.y | vb.z | vb.w) != 0) *reinterpret_cast<int4*>(&s_delta[ib]) = make_int4(0, 0, 0, 0);
      int tot_a, tot_b;
      const int ra = carry + sparse_excl_scan(a3, lane, tot_a);
      const int rb = carry + tot_a + sparse_excl_scan(b3, lane, tot_b);
      carry += tot_a + tot_b;
      const int4 oa = make_int4(ra + a0, ra + a1, ra + a2, ra + a3);
      const int4 ob = make_int4(rb + b0, rb + b1, rb + b2, rb + b3);
      *reinterpret_cast<int4*>(out + ia) = oa;
      *reinterpret_cast<int4*>(out + ib) = ob;
      if (FLAGS) emit_flags(flags, tile, it, lane, oa, ob, lo1, span);
    }
    __syncwarp();   // re-zeroing stores of all lanes are done before the next tile's events land
  }
}

// The same expansion with TMA bulk stores (verdict r01 item 6): the prefix sums are written back INTO the warp's shared
// tile and one lane hands the finished 4 KB to the copy engine (cp.async.bulk.global.shared::cta -> SASS UBLKCP);
// the warp goes on with its second buffer while the store drains, so no lane ever holds a store in a register or
// waits on the LSU.  Two 4 KB buffers per warp: a buffer is reused two tiles later, after wait_group.read.
__device__ __forceinline__ void bulk_store_4k(void* gdst, const void* ssrc) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;\n"
               "cp.async.bulk.commit_group;" ::"l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc))
               : "memory");
}

template <bool FLAGS>
__global__ void __launch_bounds__(GCI_TILE_THREADS)
depth_tile_tma_kernel(const ulonglong2* __restrict__ tile_ps, const uint16_t* __restrict__ events, int64_t n_tiles,
                      int32_t* __restrict__ depth, uint32_t* __restrict__ flags, int32_t lo1, uint32_t span) {
  extern __shared__ __align__(128) int s_dyn[];                        // [warps][2][GCI_TILE]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* const s_warp = s_dyn + warp * 2 * GCI_TILE;
  constexpr int ITERS = GCI_TILE / 256;                                // 4
  const int64_t n_warps = (int64_t)gridDim.x * (GCI_TILE_THREADS / 32);
  int64_t tile = (int64_t)blockIdx.x * (GCI_TILE_THREADS / 32) + warp;
  const ulonglong2 zero2 = make_ulonglong2(0, 0);
  // software pipeline over the warp's tiles: the tile table entry is requested two tiles ahead, the tile's first 32
  // events one tile ahead (a deeper pipeline measured slower: profiles/r02g)
  auto table = [&](int64_t t) { return t < n_tiles ? tile_ps[t] : zero2; };
  auto first_events = [&](const ulonglong2& p) {
    return lane < (uint32_t)(p.x & 0xffffffffull) ? (uint32_t)events[(uint32_t)(p.y & 0xffffffffull) + lane] : 0u;
  };
  ulonglong2 ps = table(tile), ps1 = table(tile + n_warps);
  uint32_t ev = first_events(ps);
  int cur = 0;
  for (; tile < n_tiles; tile += n_warps, cur ^= 1) {
    int* __restrict__ s_delta = s_warp + cur * GCI_TILE;
    // the bulk store that read this buffer two tiles ago is done reading (only the latest group may be pending)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int v = lane; v < GCI_TILE / 4; v += 32) reinterpret_cast<int4*>(s_delta)[v] = make_int4(0, 0, 0, 0);
    __syncwarp();
    const uint32_t n_ev = (uint32_t)(ps.x & 0xffffffffull);
    const uint32_t ev0 = (uint32_t)(ps.y & 0xffffffffull);
    int carry = (int)(uint32_t)(ps.y >> 32);                     // depth carried into the tile
    if (lane < n_ev) atomicAdd(&s_delta[ev >> 1], (ev & 1u) ? -1 : 1);
    for (uint32_t i = lane + 32; i < n_ev; i += 32) {            // more than 32 events in the tile: rare
      const uint32_t e = events[ev0 + i];
      atomicAdd(&s_delta[e >> 1], (e & 1u) ? -1 : 1);
    }
    ps = ps1;
    ev = first_events(ps);                                       // events of the next tile
    ps1 = table(tile + 2 * n_warps);
    __syncwarp();
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int ia = it * 256 + lane * 4, ib = ia + 128;
      const int4 va = *reinterpret_cast<const int4*>(&s_delta[ia]);
      const int4 vb = *reinterpret_cast<const int4*>(&s_delta[ib]);
      if (__ballot_sync(0xffffffffu, (va.x | va.y | va.z | va.w | vb.x | vb.y | vb.z | vb.w) != 0) == 0u) {
        // no read starts or ends in these 256 positions (about every second round at 30x HiFi): the depth is flat
        const int4 o = make_int4(carry, carry, carry, carry);
        *reinterpret_cast<int4*>(&s_delta[ia]) = o;
        *reinterpret_cast<int4*>(&s_delta[ib]) = o;
        if (FLAGS && (lane & 6) == 0)
          flags[tile * (GCI_TILE / 32) + it * 8 + ((lane & 1) ? 4 : 0) + (lane >> 3)] =
              (uint32_t)(carry - lo1) < span ? 0xffffffffu : 0u;
        continue;
      }
      const int a0 = va.x, a1 = a0 + va.y, a2 = a1 + va.z, a3 = a2 + va.w;
      const int b0 = vb.x, b1 = b0 + vb.y, b2 = b1 + vb.z, b3 = b2 + vb.w;
      int tot_a, tot_b;
      const int ra = carry + sparse_excl_scan(a3, lane, tot_a);
      const int rb = carry + tot_a + sparse_excl_scan(b3, lane, tot_b);
      carry += tot_a + tot_b;
      const int4 oa = make_int4(ra + a0, ra + a1, ra + a2, ra + a3);
      const int4 ob = make_int4(rb + b0, rb + b1, rb + b2, rb + b3);
      *reinterpret_cast<int4*>(&s_delta[ia]) = oa;               // depth replaces the deltas in place
      *reinterpret_cast<int4*>(&s_delta[ib]) = ob;
      if (FLAGS) emit_flags(flags, tile, it, lane, oa, ob, lo1, span);
    }
    // generic-proxy writes of the whole warp -> visible to the async proxy, then one lane starts the copy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) bulk_store_4k(depth + tile * GCI_TILE, s_delta);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the tile must have left before the CTA does
}

// ================================================================================================
// stand-alone flags (resume path / different thresholds), two-type max, N-run mask
// ================================================================================================
// flag bits of 8 consecutive positions per lane, same packing as the depth kernel
__global__ void __launch_bounds__(GCI_TILE_THREADS)
flags_kernel(const int32_t* __restrict__ depth, int64_t total, uint32_t* __restrict__ flags, int32_t lo1, uint32_t span) {
  const int lane = threadIdx.x & 31;
  const int64_t base = (int64_t)blockIdx.x * GCI_CHUNK;
#pragma unroll
  for (int it = 0; it < GCI_CHUNK / (GCI_TILE_THREADS * 8); it++) {
    const int64_t idx = base + it * GCI_TILE_THREADS * 8 + threadIdx.x * 8;
    if (idx >= total) break;                          // total is a multiple of 1024: whole warps leave together
    const int4 a = *reinterpret_cast<const int4*>(&depth[idx]);
    const int4 b = *reinterpret_cast<const int4*>(&depth[idx + 4]);
    const int o[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const uint32_t word = pack_bytes(flag8(o, lo1, span), lane);
    if ((lane & 3) == 0) flags[idx >> 5] = word;
  }
}

// per-tile sums -> per-contig sums: one warp covers one tile per round (a tile never straddles contigs)
__device__ __forceinline__ void add_tile_sum(long long acc, int64_t tile, const int64_t* __restrict__ tile_off,
                                             int32_t n_contigs, long long* __restrict__ sums, int lane) {
  acc = warp_sum_ll(acc);
  if (lane == 0 && acc) {
    const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
    atomicAdd((unsigned long long*)(sums + c), (unsigned long long)acc);
  }
}

__global__ void __launch_bounds__(GCI_TILE_THREADS)
max_kernel(const int32_t* __restrict__ da, const int32_t* __restrict__ db, int32_t* __restrict__ dout, int64_t total,
           const int64_t* __restrict__ tile_off, int32_t n_contigs, uint32_t* __restrict__ flags, int32_t lo1,
           uint32_t span, long long* __restrict__ sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp w of the CTA owns tile (chunk * 8 + w): 4 rounds of 256 positions
  const int64_t tile = (int64_t)blockIdx.x * (GCI_CHUNK / GCI_TILE) + warp;
  if (tile * GCI_TILE >= total) return;
  long long acc = 0;
#pragma unroll
  for (int it = 0; it < GCI_TILE / 256; it++) {
    const int64_t idx = tile * GCI_TILE + it * 256 + lane * 8;
    const int4 a0 = *reinterpret_cast<const int4*>(&da[idx]), a1 = *reinterpret_cast<const int4*>(&da[idx + 4]);
    const int4 b0 = *reinterpret_cast<const int4*>(&db[idx]), b1 = *reinterpret_cast<const int4*>(&db[idx + 4]);
    const int o[8] = {max(a0.x, b0.x), max(a0.y, b0.y), max(a0.z, b0.z), max(a0.w, b0.w),
                      max(a1.x, b1.x), max(a1.y, b1.y), max(a1.z, b1.z), max(a1.w, b1.w)};
    *reinterpret_cast<int4*>(&dout[idx]) = make_int4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<int4*>(&dout[idx + 4]) = make_int4(o[4], o[5], o[6], o[7]);
#pragma unroll
    for (int k = 0; k < 8; k++) acc += o[k];          // padding is 0 in both inputs
    const uint32_t word = pack_bytes(flag8(o, lo1, span), lane);
    if ((lane & 3) == 0) flags[idx >> 5] = word;
  }
  add_tile_sum(acc, tile, tile_off, n_contigs, sums, lane);
}

// per-contig sum of a loaded track (resume path)
__global__ void __launch_bounds__(GCI_TILE_THREADS)
sum_kernel(const int32_t* __restrict__ depth, int64_t total, const int64_t* __restrict__ tile_off, int32_t n_contigs,
           long long* __restrict__ sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * (GCI_CHUNK / GCI_TILE) + warp;
  if (tile * GCI_TILE >= total) return;
  long long acc = 0;
#pragma unroll
  for (int it = 0; it < GCI_TILE / 256; it++) {
    const int64_t idx = tile * GCI_TILE + it * 256 + lane * 8;
    const int4 a = *reinterpret_cast<const int4*>(&depth[idx]), b = *reinterpret_cast<const int4*>(&depth[idx + 4]);
    acc += (long long)a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  }
  add_tile_sum(acc, tile, tile_off, n_contigs, sums, lane);
}

// order-independent 64-bit checksum per contig: sum over positions of (depth + 1) * splitmix64(position) mod 2^64
// (parity gate at sizes where whole arrays do not travel; the oracle's orc_depth_hash is the same sum)
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void __launch_bounds__(GCI_TILE_THREADS)
depth_hash_kernel(const int32_t* __restrict__ depth, int64_t n_tiles, const int64_t* __restrict__ tile_off,
                  const int64_t* __restrict__ len, int32_t n_contigs, unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * (GCI_TILE_THREADS / 32) + warp;
  if (tile >= n_tiles) return;
  const int64_t c = upper_bound_minus1<int64_t>(tile_off, (int64_t)n_contigs + 1, tile);
  const int64_t p0 = (tile - tile_off[c]) * GCI_TILE, L = len[c];
  unsigned long long h = 0;
#pragma unroll
  for (int it = 0; it < GCI_TILE / 128; it++) {
    const int64_t p = p0 + it * 128 + lane * 4;
    const int4 v = *reinterpret_cast<const int4*>(depth + tile * GCI_TILE + it * 128 + lane * 4);
    const int d[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (p + k < L) h += (unsigned long long)((long long)d[k] + 1) * splitmix64((unsigned long long)(p + k));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) h += __shfl_xor_sync(0xffffffffu, h, s);
  if (lane == 0 && h) atomicAdd(out + c, h);
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
// narrow fetch: depth as uint8 / uint16 when every value fits (4x / 2x fewer bytes over PCIe)
// ================================================================================================
template <typename T>
__global__ void narrow_kernel(const int32_t* __restrict__ depth, int64_t n, T* __restrict__ out,
                              unsigned int* __restrict__ overflow) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  constexpr unsigned MAXV = sizeof(T) == 1 ? 0xffu : 0xffffu;
  bool bad = false;
  if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(depth + i) & 15) == 0)) {
    const int4 v = *reinterpret_cast<const int4*>(depth + i);
    bad = ((unsigned)v.x > MAXV) | ((unsigned)v.y > MAXV) | ((unsigned)v.z > MAXV) | ((unsigned)v.w > MAXV);
    out[i] = (T)v.x; out[i + 1] = (T)v.y; out[i + 2] = (T)v.z; out[i + 3] = (T)v.w;
  } else {
    for (int64_t k = i; k < n && k < i + 4; k++) {
      const int v = depth[k];
      bad |= (unsigned)v > MAXV;
      out[k] = (T)v;
    }
  }
  if (__any_sync(__activemask(), bad) && bad) atomicOr(overflow, 1u);
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

int gci_depth_occupancy(gci_ctx* ctx) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, depth_tile_kernel<true>, GCI_TILE_THREADS, 0) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = 4;
  }
  return n;
}

static inline int32_t flag_lo1(int32_t lo) { return lo == INT32_MAX ? INT32_MAX : lo + 1; }
static inline uint32_t flag_span(int32_t lo, int32_t hi) {
  return hi > lo ? (uint32_t)((int64_t)hi - (int64_t)lo) : 0u;
}
static inline unsigned chunk_grid(gci_ctx* ctx) {
  return (unsigned)((ctx->total_padded + GCI_CHUNK - 1) / GCI_CHUNK);
}

int gci_compute_flags(gci_ctx* ctx, int track, int32_t lo, int32_t hi) {
  Track& t = ctx->track[track];
  ctx->stage_begin(GCI_ST_FLAGS);
  if (ctx->n_tiles) {
    flags_kernel<<<chunk_grid(ctx), GCI_TILE_THREADS, 0, ctx->stream>>>(
        t.depth.as<int32_t>(), ctx->total_padded, t.flags.as<uint32_t>(), flag_lo1(lo), flag_span(lo, hi));
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  t.flags_valid = true;
  t.flags_lo = lo;
  t.flags_hi = hi;
  return GCI_OK;
}

// allocate the track and the tile table, zero the table and the depth sums; *bk describes the event counting
// for whoever does it (bucket_count_kernel, or the join kernel in gci_pipeline)
int gci_depth_prepare(gci_ctx* ctx, int32_t track, int32_t flank_len, BucketArgs* bk) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
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
  memset(bk, 0, sizeof *bk);
  if (nt == 0) return GCI_OK;
  if (nt >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "too many tiles");
  // one scratch block: per tile the table entry (pack u64, scan u64: written by the tile scan), then three u32
  // arrays zeroed with one memset: start events, end events, fill cursor
  GCI_TRY(ctx->ensure(ctx->tile_cnt, 28 * (size_t)nt));
  GCI_TRY(ctx->ensure(ctx->events, 2 * 2 * (size_t)std::max<int64_t>(1, ctx->shard.on ? ctx->shard.surv_slots : (int64_t)ctx->n_reads)));
  uint32_t* counts = reinterpret_cast<uint32_t*>(ctx->tile_cnt.as<ulonglong2>() + nt);   // [start | end | cursor]
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(counts, 0, 12 * (size_t)nt, ctx->stream));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  bk->fl = flank_len;
  bk->len = ctx->d_len.as<int64_t>();
  bk->tile_off = ctx->d_tile_off.as<int64_t>();
  bk->cnt_start = counts;
  bk->cnt_end = counts + nt;
  bk->sums = t.sums.as<long long>();
  bk->n_contigs = ctx->n_contigs;
  return GCI_OK;
}

// body of gci_depth; gci_pipeline calls it directly (it must not bump ctx->epoch)
int gci_depth_enqueue(gci_ctx* ctx, int32_t track, int32_t flank_len, int32_t lo, int32_t hi) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->filtered) return ctx->fail(GCI_E_ARG, "gci_depth before gci_filter");
  Track& t = ctx->track[track];
  const int64_t nt = ctx->n_tiles;
  // survivor slots: one per read, or for a sharded read set one per inbox slot (shard.cu)
  const uint32_t nr = ctx->shard.on ? (uint32_t)ctx->shard.surv_slots : ctx->n_reads;
  const bool counted = ctx->counted_track == track && ctx->counted_flank == flank_len;
  ctx->counted_track = -1;
  ctx->stage_begin(GCI_ST_BUCKET);
  if (!counted) {
    BucketArgs bk;
    GCI_TRY(gci_depth_prepare(ctx, track, flank_len, &bk));
    if (nr && nt) {
      bucket_count_kernel<<<(unsigned)std::min<int64_t>(((int64_t)nr + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
          nr, ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(), ctx->surv_end.as<int32_t>(), bk);
      GCI_LAUNCH_CHECK(ctx);
    }
  }
  if (nt == 0) {
    ctx->stage_end();
    return GCI_OK;
  }
  ulonglong2* tile_ps = ctx->tile_cnt.as<ulonglong2>();
  uint32_t* counts = reinterpret_cast<uint32_t*>(tile_ps + nt);
  uint32_t* cursor = counts + 2 * nt;
  GCI_TRY(gci_scan_tile_pack(ctx, counts, counts + nt, tile_ps, nt));
  if (nr) {
    bucket_fill_kernel<<<(unsigned)std::min<int64_t>(((int64_t)nr + 255) / 256, (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        nr, ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(), ctx->surv_end.as<int32_t>(), flank_len,
        ctx->d_len.as<int64_t>(), ctx->d_tile_off.as<int64_t>(), tile_ps, cursor, ctx->events.as<uint16_t>(),
        ctx->n_contigs, ctx->shard.on ? ctx->shard.send_cnt.as<uint32_t>() + GCI_MAX_RANKS : nullptr);
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  if (ctx->depth_ctas_per_sm <= 0) ctx->depth_ctas_per_sm = gci_depth_occupancy(ctx);
  const bool fuse = !(lo == INT32_MIN && hi == INT32_MIN);
  const int32_t lo1 = fuse ? flag_lo1(lo) : 0;
  const uint32_t span = fuse ? flag_span(lo, hi) : 0u;
  ctx->stage_begin(GCI_ST_DEPTH);
  // persistent grid: resident CTAs per SM x SM count, 8 tiles in flight per CTA
  const int64_t ctas_needed = (nt + GCI_TILE_THREADS / 32 - 1) / (GCI_TILE_THREADS / 32);
  const unsigned grid = (unsigned)std::min<int64_t>(ctas_needed, (int64_t)ctx->sm_count * ctx->depth_ctas_per_sm);
  const uint16_t* evs = ctx->events.as<uint16_t>();
  int32_t* dp = t.depth.as<int32_t>();
  uint32_t* fp = t.flags.as<uint32_t>();
  static int use_tma = -1, tma_ctas = 0;
  constexpr size_t TMA_SMEM = sizeof(int) * 2 * GCI_TILE * (GCI_TILE_THREADS / 32);      // 64 KB
  if (use_tma < 0) {
    const char* g = getenv("GCI_DEPTH_TMA");
    use_tma = !(g && g[0] == '0');
    if (use_tma &&
        (cudaFuncSetAttribute(depth_tile_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess ||
         cudaFuncSetAttribute(depth_tile_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess ||
         cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tma_ctas, depth_tile_tma_kernel<true>, GCI_TILE_THREADS, TMA_SMEM) != cudaSuccess ||
         tma_ctas < 1)) {
      cudaGetLastError();
      use_tma = 0;
    }
  }
  if (use_tma) {
    const unsigned tgrid = (unsigned)std::min<int64_t>(ctas_needed, (int64_t)ctx->sm_count * tma_ctas);
    if (fuse) depth_tile_tma_kernel<true><<<tgrid, GCI_TILE_THREADS, TMA_SMEM, ctx->stream>>>(tile_ps, evs, nt, dp, fp, lo1, span);
    else depth_tile_tma_kernel<false><<<tgrid, GCI_TILE_THREADS, TMA_SMEM, ctx->stream>>>(tile_ps, evs, nt, dp, fp, lo1, span);
  } else if (fuse) {
    depth_tile_kernel<true><<<grid, GCI_TILE_THREADS, 0, ctx->stream>>>(tile_ps, evs, nt, dp, fp, lo1, span);
  } else {
    depth_tile_kernel<false><<<grid, GCI_TILE_THREADS, 0, ctx->stream>>>(tile_ps, evs, nt, dp, fp, lo1, span);
  }
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  // depth and flags are both un-masked here (the reference writes the single-type .depth.gz before
  // masking, GCI.py:310 vs :993); gci_mask_gaps patches both over the N-runs
  t.flags_valid = fuse;
  t.sums_valid = true;
  t.flags_lo = lo;
  t.flags_hi = hi;
  t.n_intervals = 0;
  t.n_owners = 0;
  return GCI_OK;
}

extern "C" {

int gci_depth(gci_ctx* ctx, int32_t track, int32_t flank_len, int32_t lo, int32_t hi) {
  if (ctx) ctx->epoch++;
  return gci_depth_enqueue(ctx, track, flank_len, lo, hi);
}

int gci_mask_gaps(gci_ctx* ctx, int32_t track) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  ctx->epoch++;
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
  ctx->epoch++;
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
    max_kernel<<<chunk_grid(ctx), GCI_TILE_THREADS, 0, ctx->stream>>>(
        ctx->track[ta].depth.as<int32_t>(), ctx->track[tb].depth.as<int32_t>(), t.depth.as<int32_t>(),
        ctx->total_padded, ctx->d_tile_off.as<int64_t>(), ctx->n_contigs, t.flags.as<uint32_t>(), flag_lo1(flo),
        flag_span(flo, fhi), t.sums.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  t.flags_valid = true;
  t.sums_valid = true;
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
  // the sums ride along with gci_depth / gci_merge_max / gci_mask_gaps; after gci_load_depth they are
  // recomputed from the depth array
  if (!t.sums_valid) {
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  if (ctx->n_tiles) {
    sum_kernel<<<chunk_grid(ctx), GCI_TILE_THREADS, 0, ctx->stream>>>(
        t.depth.as<int32_t>(), ctx->total_padded, ctx->d_tile_off.as<int64_t>(), ctx->n_contigs,
        t.sums.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  t.sums_valid = true;
  }
  GCI_TRY(gci_d2h(ctx, sums, t.sums.p, sizeof(int64_t) * (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_depth_hash(gci_ctx* ctx, int32_t track, uint64_t* out) {
  if (!ctx || !out || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_depth_hash: track %d holds no depth", track);
  DevBuf& d_out = ctx->misc;
  GCI_TRY(ctx->ensure(d_out, 8 * (size_t)std::max(1, ctx->n_contigs)));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_out.p, 0, 8 * (size_t)std::max(1, ctx->n_contigs), ctx->stream));
  if (ctx->n_tiles) {
    depth_hash_kernel<<<(unsigned)((ctx->n_tiles + 7) / 8), GCI_TILE_THREADS, 0, ctx->stream>>>(
        t.depth.as<int32_t>(), ctx->n_tiles, ctx->d_tile_off.as<int64_t>(), ctx->d_len.as<int64_t>(), ctx->n_contigs,
        d_out.as<unsigned long long>());
    GCI_LAUNCH_CHECK(ctx);
  }
  GCI_TRY(gci_d2h(ctx, out, d_out.p, 8 * (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_fetch_depth_narrow(gci_ctx* ctx, int32_t track, int32_t contig, void* out, int64_t n, int32_t width,
                           int32_t* overflow) {
  if (!ctx || !out || !overflow || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || n != ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_fetch_depth_narrow: contig %d / length %lld mismatch", contig, (long long)n);
  if (width != 1 && width != 2) return ctx->fail(GCI_E_ARG, "width must be 1 or 2 (use gci_fetch_depth for int32)");
  *overflow = 0;
  if (n == 0) return GCI_OK;
  const int32_t* d = t.depth.as<int32_t>() + ctx->pos_off[contig];
  DevBuf& stage = ctx->tmp[8];
  GCI_TRY(ctx->ensure(stage, (size_t)n * width + 16));
  unsigned int* d_ovf = reinterpret_cast<unsigned int*>(ctx->d_err.as<unsigned long long>() + 3);
  GCI_TRY(ctx->ensure(ctx->d_err, 4 * sizeof(unsigned long long)));
  d_ovf = reinterpret_cast<unsigned int*>(ctx->d_err.as<unsigned long long>() + 3);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_ovf, 0, sizeof(unsigned int), ctx->stream));
  const unsigned grid = (unsigned)((n / 4 + 1 + 255) / 256);
  ctx->stage_begin(GCI_ST_D2H);
  if (width == 1) narrow_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(d, n, stage.as<uint8_t>(), d_ovf);
  else narrow_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(d, n, stage.as<uint16_t>(), d_ovf);
  GCI_LAUNCH_CHECK(ctx);
  unsigned int* h = (unsigned int*)ctx->pinned(sizeof(unsigned int));
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_TRY(gci_d2h(ctx, h, d_ovf, sizeof(unsigned int)));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *overflow = (int32_t)*h;
  if (*h == 0) {
    GCI_TRY(gci_d2h(ctx, out, stage.p, (size_t)n * width));
    ctx->stage_end();
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  } else {
    ctx->stage_end();
  }
  return GCI_OK;
}

int gci_depth_text(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count, char* out, int64_t cap,
                   int64_t* n_bytes) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_bytes) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
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
