// Filter stage: CIGAR statistics, per-record gates, last-record-wins dedup, PAF election,
// cross-file same-read join.  Reference: GCI.py:146-169 (read_sam), :211-254 (PAF leg),
// :257-301 (fan-out merge + join).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

#ifndef GCI_USE_TMA
#define GCI_USE_TMA 1
#endif

// ================================================================================================
// K1  cigar_stats: segmented sum of CIGAR op lengths per record, tiled over the op stream
// ================================================================================================
// The packed op stream (uint32 `len << 4 | op`, BAM native) is cut into tiles of CIG_TILE ops and summed
// into five classes per record
//     Mx = M + '=' + X,  I,  D,  N,  S
// (the reference only ever uses M+eq+X as one quantity: GCI.py:164-165; N is needed for reference_end).
// Two kernels share the tiles (split once at upload, cigar_tile_class_kernel): K1a takes the tiles that touch
// many records (HiFi: ~70 records per tile) and gates the records it sees completely on the spot; K1b streams
// the tiles that lie inside one or a few long records (ONT: thousands of ops per record).  Work per tile is
// constant whatever the ops-per-record distribution.
constexpr int CIG_THREADS = 256;
constexpr int CIG_OPT = 8;                       // ops per thread
constexpr int CIG_TILE = CIG_THREADS * CIG_OPT;  // 2048 ops = 8 KB
constexpr int CIG_CAP = 512;                     // records per tile handled through shared memory
constexpr int CST_MAX_LOC = 8;                   // tiles touching at most this many records stream (K1b)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!done);
}

// 1-D TMA bulk copy global -> shared (SASS: UBLKCP); dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// first and last record of every op tile: record r (ops [a,b)) is the first record of the tiles whose
// first op lies in [a,b) and the last record of the tiles whose last op lies in [a,b)
__global__ void cigar_tile_index_kernel(const uint64_t* __restrict__ off, int64_t n, int2* __restrict__ tile_rec,
                                        int64_t n_tiles, int64_t n_ops) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint64_t a = off[r], b = off[r + 1];
  if (b <= a) return;
  const int64_t k0 = (int64_t)((a + CIG_TILE - 1) / CIG_TILE);
  const int64_t k1 = (int64_t)((b - 1) / CIG_TILE);
  for (int64_t k = k0; k <= k1 && k < n_tiles; k++) tile_rec[k].x = (int32_t)r;
  const int64_t l0 = (int64_t)(a / CIG_TILE);
  const int64_t l1 = (int64_t)(b / CIG_TILE) - 1;
  for (int64_t k = l0; k <= l1 && k < n_tiles; k++) tile_rec[k].y = (int32_t)r;
  if ((int64_t)b == n_ops) tile_rec[n_tiles - 1].y = (int32_t)r;
}

// tiles touching more than CST_MAX_LOC records go to the staged kernel (K1a), the others stream (K1b): the
// staged tiles fill the list from the front, the streaming ones from the back.  The order inside either part
// is arbitrary: both kernels only ever ADD u32 partial sums, which commute exactly.
__global__ void cigar_tile_class_kernel(const int2* __restrict__ tile_rec, int64_t n_tiles,
                                        int32_t* __restrict__ list, unsigned int* __restrict__ counts /* [2] */) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int2 tr = tile_rec[t];
  if (tr.y - tr.x + 1 > CST_MAX_LOC) list[atomicAdd(counts, 1u)] = (int32_t)t;
  else list[n_tiles - 1 - (int64_t)atomicAdd(counts + 1, 1u)] = (int32_t)t;
}

struct CigAcc {
  uint32_t tot, i, d, n, s;   // tot = M + '=' + X + I + D
  __device__ __forceinline__ void clear() { tot = i = d = n = s = 0; }
  // branch-free: a branch per op costs more than the five selects
  __device__ __forceinline__ void add(uint32_t op) {
    const uint32_t c = op & 15u, l = op >> 4;
    tot += l * ((0x187u >> c) & 1u);          // op codes 0 (M), 1 (I), 2 (D), 7 (=), 8 (X)
    i += (c == 1u) ? l : 0u;
    d += (c == 2u) ? l : 0u;
    n += (c == 3u) ? l : 0u;
    s += (c == 4u) ? l : 0u;
  }
  __device__ __forceinline__ bool any() const { return (tot | n | s) != 0; }
  __device__ __forceinline__ void warp_reduce() {   // REDUX.SUM: one instruction per counter
    tot = __reduce_add_sync(0xffffffffu, tot);
    i = __reduce_add_sync(0xffffffffu, i);
    d = __reduce_add_sync(0xffffffffu, d);
    n = __reduce_add_sync(0xffffffffu, n);
    s = __reduce_add_sync(0xffffffffu, s);
  }
  // stats layout: [Mx, I, D, N, S]
  __device__ __forceinline__ void flush(uint32_t* p) const {
    const uint32_t mx = tot - i - d;
    if (mx) atomicAdd(p + 0, mx);
    if (i) atomicAdd(p + 1, i);
    if (d) atomicAdd(p + 2, d);
    if (n) atomicAdd(p + 3, n);
    if (s) atomicAdd(p + 4, s);
  }
  // after warp_reduce(): lanes 0..4 add one counter each
  __device__ __forceinline__ void flush_warp(uint32_t* p, int lane) const {
    const uint32_t v = lane == 0 ? tot - i - d : lane == 1 ? i : lane == 2 ? d : lane == 3 ? n : s;
    if (lane < 5 && v) atomicAdd(p + lane, v);
  }
};

// K2  gate + dedup of ONE record from its CIGAR sums (GCI.py:153-168).  Runs inside the staged CIGAR kernel
// for records that lie completely inside one tile (their sums never reach HBM) and in gate_span_kernel for
// the others (records cut by a tile border, records of streaming tiles, records without ops).
// d_err layout: [0] bit mask (1 = NM missing, 2 = zero clip denominator, 4 = zero identity denominator,
//               8 = zero query length in the join, 16 = zero PAF aln length), [1] first offending index
struct GateArgs {
  const int32_t* ref_id;
  const int32_t* ref_start;
  const uint8_t* mapq;
  const uint16_t* flag;
  const int32_t* nm;
  const uint32_t* read_id;
  const uint8_t* selected;
  int32_t n_contigs;
  uint32_t n_reads;
  int32_t map_qual, mq_cutoff;
  double ip, cp;
  int32_t* ref_end;
  long long* win;
  uint8_t* highq;
  unsigned long long* err;
};

__device__ __forceinline__ void gate_one(const GateArgs& g, int64_t r, uint32_t sMx, uint32_t sI, uint32_t sD,
                                         uint32_t sN, uint32_t sS) {
  // every column of the record is requested up front: behind the early returns below the loads would be issued
  // one after the other, six DRAM round trips deep
  const int32_t start = g.ref_start[r];
  const int32_t c = g.ref_id[r];
  const uint32_t f = g.flag[r];
  const int32_t mq = g.mapq[r];
  const int32_t nmv = g.nm[r];
  const uint32_t q = g.read_id[r];
  const long long Mx = sMx, I = sI, D = sD, N = sN, S = sS;
  long long rlen = Mx + D + N;                       // htslib bam_cigar2rlen
  if (rlen == 0) rlen = 1;                           // htslib bam_endpos
  g.ref_end[r] = (int32_t)(start + rlen);
  if (c < 0 || c >= g.n_contigs || !g.selected[c]) return;       // never fetched (GCI.py:151, :202-207)
  if (f & (0x4u | 0x100u | 0x800u)) return;                      // :153-156
  if (mq < g.map_qual) return;                                   // :156
  if (nmv == INT32_MIN) {                                        // KeyError at :163
    atomicOr(g.err, 1ull);
    atomicMin(g.err + 1, (unsigned long long)r);
    return;
  }
  const long long mm = (long long)nmv - (I + D);                 // :164
  const long long d1 = Mx + I + S;
  if (d1 == 0) {
    atomicOr(g.err, 2ull);
    atomicMin(g.err + 1, (unsigned long long)r);
    return;
  }
  if (!ratio_le(S, d1, g.cp)) return;                            // :165, fp64 div.rn like Python int/int
  const long long d2 = Mx + I + D;
  if (d2 == 0) {
    atomicOr(g.err, 4ull);
    atomicMin(g.err + 1, (unsigned long long)r);
    return;
  }
  if (!ratio_ge(Mx - mm, d2, g.ip)) return;
  if (q >= g.n_reads) return;
  // fetch order = contigs in header order, file order inside: the later record wins (:166, :269)
  atomicMax(g.win + q, ((long long)c << 32) | (long long)r);
  if (mq >= g.mq_cutoff) g.highq[q] = 1;                         // :167-168
}

// K1a  staged kernel: tiles holding many records (HiFi: ~70 records per 2048 ops).
// The tile is staged with one TMA bulk copy; every thread sums its 8 consecutive ops into three counters
// (all lengths, I, D, S) and leaves them in shared memory; one thread per record then adds the partial sums of
// the 8-op blocks its record covers and the ragged ends (at most 7 ops at either side) straight from the
// staged ops: no atomics, no search and no divergence on the common path.  N, H, P, B ops are rare: the thread
// that meets one looks up its record and adds it to a small per-record side table.
// GATE: records that lie completely inside the tile are gated right here (gate_one) and their sums are not
// stored; !GATE (gci_fetch_cigar_stats): every record's sums are stored, nothing is gated.
template <bool GATE>
__global__ void __launch_bounds__(CIG_THREADS, 8)
cigar_stats_kernel(const uint32_t* __restrict__ cigar, const uint64_t* __restrict__ off, int64_t n_rec,
                   int64_t n_ops, const int2* __restrict__ tile_rec, const int32_t* __restrict__ tile_list,
                   uint32_t* __restrict__ stats /* [n_rec][8] */, GateArgs gate) {
  __shared__ __align__(128) uint32_t s_ops[CIG_TILE];
  __shared__ int32_t s_off[CIG_CAP + 2];          // record starts relative to the tile, clamped
  __shared__ uint4 s_part[CIG_THREADS];           // (all, I, D, S) of thread t's 8 ops
  __shared__ uint32_t s_rare[CIG_CAP * 2];        // per record: N, other (H, P, B)
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int64_t tile = tile_list ? (int64_t)tile_list[blockIdx.x] : (int64_t)blockIdx.x;
  const int64_t o0 = tile * CIG_TILE;
  const int tile_n = (int)min((int64_t)CIG_TILE, n_ops - o0);
#if GCI_USE_TMA
  // the bulk copy depends on nothing but the tile number: it flies while the record table is read
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    const uint32_t bytes = (uint32_t)((tile_n * 4 + 15) & ~15);
    mbar_expect_tx(&s_bar, bytes);
    tma_load_1d(s_ops, cigar + o0, bytes, &s_bar);
  }
#endif
  const int2 tr = tile_rec[tile];
  const int64_t r_lo = tr.x;                      // first / last record with an op in this tile
  const int n_loc = tr.y - tr.x + 1;

  const int first = tid * CIG_OPT;
  const int nb = min(CIG_OPT, tile_n - first);    // my ops (<= 0: none)
  if (n_loc > CIG_CAP) {
    // pathological tile (more than 512 records in 2048 ops): global atomics, no staging
    if (nb > 0) {
      int64_t rl = upper_bound_minus1<uint64_t>(off + r_lo, n_loc, (uint64_t)(o0 + first));
      CigAcc acc;
      acc.clear();
      for (int k = 0; k < nb; k++) {
        const uint64_t o = (uint64_t)(o0 + first + k);
        if (o >= off[r_lo + rl + 1]) {
          if (acc.any()) acc.flush(stats + (r_lo + rl) * 8);
          acc.clear();
          do { rl++; } while (o >= off[r_lo + rl + 1]);
        }
        acc.add(cigar[o]);
      }
      if (acc.any()) acc.flush(stats + (r_lo + rl) * 8);
    }
#if GCI_USE_TMA
    __syncthreads();                              // s_bar initialised before anyone polls it
    mbar_wait(&s_bar, 0);                         // never leave with a bulk copy in flight into our smem
#endif
    return;
  }

  for (int i = tid; i < n_loc * 2; i += CIG_THREADS) s_rare[i] = 0u;
#if !GCI_USE_TMA
  for (int v = tid; v * 4 < tile_n; v += CIG_THREADS)
    reinterpret_cast<uint4*>(s_ops)[v] = reinterpret_cast<const uint4*>(cigar + o0)[v];
#endif
  for (int i = tid; i <= n_loc; i += CIG_THREADS) {
    const long long rel = (long long)off[r_lo + i] - (long long)o0;
    s_off[i] = (int32_t)max(-1ll, min(rel, (long long)CIG_TILE + 1));
  }
  __syncthreads();                                // s_rare, s_off and the mbarrier are set up
#if GCI_USE_TMA
  mbar_wait(&s_bar, 0);
#endif

  // ---- phase 1: 8 ops per thread -> (all, I, D, S); rare ops -> side table ----
  uint32_t tot = 0, ci = 0, cd = 0, cs = 0;
  if (nb > 0) {
    const uint4 qa = reinterpret_cast<const uint4*>(s_ops)[tid * 2];
    const uint4 qb = reinterpret_cast<const uint4*>(s_ops)[tid * 2 + 1];
    uint32_t ops[CIG_OPT] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    if (nb < CIG_OPT) {
#pragma unroll
      for (int k = 0; k < CIG_OPT; k++)
        if (k >= nb) ops[k] = 0u;                 // beyond the end of the op stream: M of length 0
    }
    uint32_t seen = 0;
#pragma unroll
    for (int k = 0; k < CIG_OPT; k++) {
      const uint32_t c = ops[k] & 15u, l = ops[k] >> 4;
      seen |= __funnelshift_l(0u, 1u, ops[k]);
      tot += l;
      ci += (c == 1u) ? l : 0u;
      cd += (c == 2u) ? l : 0u;
      cs += (c == 4u) ? l : 0u;
    }
    if ((seen | (seen >> 16)) & 0xFE68u) {        // N, H, P, B (or an undefined code)
#pragma unroll 1
      for (int k = 0; k < nb; k++) {              // re-read from shared memory: ops[] stays in registers
        const uint32_t w = s_ops[first + k], c = w & 15u, l = w >> 4;
        if (!((0xFE68u >> c) & 1u) || l == 0u) continue;
        int lo = 0, hi = n_loc;                   // s_off[lo] <= first + k < s_off[hi] (virtual)
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (s_off[mid] <= first + k) lo = mid; else hi = mid;
        }
        atomicAdd(&s_rare[lo * 2 + (c == 3u ? 0 : 1)], l);
      }
    }
  }
  s_part[tid] = make_uint4(tot, ci, cd, cs);
  __syncthreads();
  // ---- phase 2: one thread per record: whole 8-op blocks come from s_part, the ragged ends from s_ops ----
  auto add_ops = [&](int p0, int p1, uint4& v) {
    for (int k = p0; k < p1; k++) {
      const uint32_t w = s_ops[k], c = w & 15u, l = w >> 4;
      v.x += l;
      v.y += (c == 1u) ? l : 0u;
      v.z += (c == 2u) ? l : 0u;
      v.w += (c == 4u) ? l : 0u;
    }
  };
  for (int i = tid; i < n_loc; i += CIG_THREADS) {
    const int a = max(s_off[i], 0), b = min(s_off[i + 1], tile_n);
    if (b <= a) continue;                         // no op of this record here (its row is zero already)
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    const int ta = (a + 7) >> 3, tb = b >> 3;     // threads whose 8 ops lie completely inside [a, b)
    if (ta <= tb) {
      add_ops(a, ta * 8, v);
      for (int t = ta; t < tb; t++) {
        const uint4 q = s_part[t];
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      add_ops(tb * 8, b, v);
    } else {
      add_ops(a, b, v);
    }
    const uint32_t rn = s_rare[i * 2], ro = s_rare[i * 2 + 1];
    const uint32_t vi = v.y, vd = v.z, rs = v.w;
    const uint32_t mx = v.x - vi - vd - rs - rn - ro;
    const bool complete = s_off[i] >= 0 && s_off[i + 1] <= tile_n;
    uint32_t* g = stats + (r_lo + i) * 8;
    if (complete) {
      if (GATE) {
        gate_one(gate, r_lo + i, mx, vi, vd, rn, rs);
      } else {
        *reinterpret_cast<uint4*>(g) = make_uint4(mx, vi, vd, rn);
        g[4] = rs;
      }
    } else {
      if (mx) atomicAdd(g + 0, mx);
      if (vi) atomicAdd(g + 1, vi);
      if (vd) atomicAdd(g + 2, vd);
      if (rn) atomicAdd(g + 3, rn);
      if (rs) atomicAdd(g + 4, rs);
    }
  }
}

// K1a'  the same tiles with ONE LANE PER RECORD: the tile is staged by a TMA bulk copy, then lane i walks the ops of
// record r_lo + i sequentially (consecutive records start ~31 words apart for HiFi: conflict-free LDS), classifying
// every op branch-free into the five counters, and gates the record on the spot when it lies inside the tile.  No
// per-thread partial sums, no ragged ends, no rare-op side table, one block barrier.  It executes about 40 % fewer
// instructions than the block-sum kernel on HiFi records (31 +- 8 ops: a warp runs as long as its longest record);
// a tile that is one long record in a few lanes (mixed files) is still correct, only slower.  GCI_CIGAR_LANE=0
// selects the block-sum kernel.
constexpr int CIGL_THREADS = 128;
template <bool GATE>
__global__ void __launch_bounds__(CIGL_THREADS)
cigar_lane_kernel(const uint32_t* __restrict__ cigar, const uint64_t* __restrict__ off, int64_t n_rec, int64_t n_ops,
                  const int2* __restrict__ tile_rec, const int32_t* __restrict__ tile_list,
                  uint32_t* __restrict__ stats /* [n_rec][8] */, GateArgs gate) {
  __shared__ __align__(128) uint32_t s_ops[CIG_TILE];
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x;
  const int64_t tile = tile_list ? (int64_t)tile_list[blockIdx.x] : (int64_t)blockIdx.x;
  const int64_t o0 = tile * CIG_TILE;
  const int tile_n = (int)min((int64_t)CIG_TILE, n_ops - o0);
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    const uint32_t bytes = (uint32_t)((tile_n * 4 + 15) & ~15);
    mbar_expect_tx(&s_bar, bytes);
    tma_load_1d(s_ops, cigar + o0, bytes, &s_bar);
  }
  const int2 tr = tile_rec[tile];
  const int64_t r_lo = tr.x;
  const int n_loc = tr.y - tr.x + 1;
  // the record borders of my first record, requested while the copy flies
  long long a = 0, b = 0;
  if (tid < n_loc) {
    a = (long long)off[r_lo + tid] - o0;
    b = (long long)off[r_lo + tid + 1] - o0;
  }
  __syncthreads();                                // the mbarrier is initialised before anyone polls it
  mbar_wait(&s_bar, 0);
  for (int i = tid; i < n_loc; i += CIGL_THREADS) {
    if (i != tid) {
      a = (long long)off[r_lo + i] - o0;
      b = (long long)off[r_lo + i + 1] - o0;
    }
    const bool complete = a >= 0 && b <= tile_n;
    const int k0 = (int)max(a, 0ll), k1 = (int)min(b, (long long)tile_n);
    // fast walk: four counters (all lengths, I, D, S) and the set of op codes seen; M/=/X bases are what is left.
    // N, H, P, B (or an undefined code) are rare in DNA alignments: a record holding one is walked again with the
    // five-counter classifier.  (13 instead of 19 instructions per op in the loop that is ~all of this kernel.)
    CigAcc acc;
    {
      uint32_t all = 0, ci = 0, cd = 0, cs = 0, seen = 0;
#pragma unroll 4
      for (int k = k0; k < k1; k++) {
        const uint32_t w = s_ops[k], c = w & 15u, l = w >> 4;
        seen |= __funnelshift_l(0u, 1u, w);          // bit (w & 31): the op code, plus 16 for odd lengths
        all += l;
        ci += (c == 1u) ? l : 0u;
        cd += (c == 2u) ? l : 0u;
        cs += (c == 4u) ? l : 0u;
      }
      if ((seen | (seen >> 16)) & 0xFE68u) {
        acc.clear();
        for (int k = k0; k < k1; k++) acc.add(s_ops[k]);
      } else {
        acc.tot = all - cs;
        acc.i = ci;
        acc.d = cd;
        acc.n = 0u;
        acc.s = cs;
      }
    }
    const uint32_t mx = acc.tot - acc.i - acc.d;
    uint32_t* g = stats + (r_lo + i) * 8;
    // (a tile touching more than CIG_CAP records gates nothing itself: its records are on the span list, like the
    // block-sum kernel's pathological tiles — cigar_record_class_kernel decides who gates what)
    if (complete && k1 > k0 && (!GATE || n_loc <= CIG_CAP)) {
      if (GATE) {
        gate_one(gate, r_lo + i, mx, acc.i, acc.d, acc.n, acc.s);
      } else {
        *reinterpret_cast<uint4*>(g) = make_uint4(mx, acc.i, acc.d, acc.n);
        g[4] = acc.s;
      }
    } else if (k1 > k0) {                         // cut by a tile border: partial sums, gated later from its row
      acc.flush(g);
    }
  }
}

// K1b  streaming kernel: tiles touching at most CST_MAX_LOC records (ONT: thousands of ops per record).
// Persistent warps, one 2048-op tile per warp at a time, no block barrier.  Every warp owns a ring of CST_NST
// shared-memory stages of 512 ops (2 KB) filled by TMA bulk copies (cp.async.bulk -> SASS UBLKCP, one
// mbarrier per stage): three stages = 6 KB per warp are in flight while the fourth is summed, across tile
// borders, and they cost no registers.  The next tile's record table is fetched one tile ahead.  The sums stay
// in registers until a record ends (one REDUX per counter, five atomics): a few times per tile at most.
constexpr int CST_WARPS = 4;          // warps per CTA (32 KB of ring: six CTAs = 24 warps and 192 KB per SM)
constexpr int CST_NST = 4;            // ring stages per warp
constexpr int CST_GROUP = 512;        // ops per stage

struct CstTile {
  int64_t o0;        // first op of the tile
  int tile_n;        // ops in the tile (2048 except for the last tile of the stream)
  int x, n_loc;      // first record with an op in the tile, number of records touching it
};

__global__ void __launch_bounds__(CST_WARPS * 32, 6)
cigar_stream_kernel(const uint32_t* __restrict__ cigar, const uint64_t* __restrict__ off, int64_t n_ops,
                    const int2* __restrict__ tile_rec, const int32_t* __restrict__ tile_list, int64_t n_list,
                    uint32_t* __restrict__ stats) {
  __shared__ __align__(128) uint32_t s_ring[CST_WARPS][CST_NST][CST_GROUP];
  __shared__ __align__(8) uint64_t s_bar[CST_WARPS][CST_NST];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * CST_WARPS;
  int64_t idx = (int64_t)blockIdx.x * CST_WARPS + wp;
  if (idx >= n_list) return;                      // whole warps leave; nothing below is block-wide
  uint32_t (*ring)[CST_GROUP] = s_ring[wp];
  uint64_t* bar = s_bar[wp];
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < CST_NST; st++) mbar_init(&bar[st], 1);
  }
  __syncwarp();
  auto tile_of = [&](int64_t i) -> int64_t { return tile_list ? (int64_t)tile_list[i] : i; };
  auto describe = [&](int64_t tile, int2 tr) {
    CstTile t;
    t.o0 = tile * CIG_TILE;
    t.tile_n = (int)min((int64_t)CIG_TILE, n_ops - t.o0);
    t.x = tr.x;
    t.n_loc = tr.y - tr.x + 1;
    return t;
  };
  // lane i keeps the end of record t.x + i, relative to the tile (the last one may lie beyond it)
  auto load_ends = [&](const CstTile& t) -> int {
    int e = CIG_TILE + 1;
    if (lane < t.n_loc) e = (int)min((long long)off[t.x + lane + 1] - (long long)t.o0, (long long)CIG_TILE + 1);
    return e;
  };

  // ---- producer: lane 0 issues group g of the tile starting at op o0 into stage g.  A tile has as many groups
  // as the ring has stages, so group g always lives in stage g and its barrier flips once per tile. ----
  static_assert(CIG_TILE / CST_GROUP == CST_NST, "one ring stage per group of a tile");
  auto produce = [&](int64_t o0, int tile_n, int g) {
    if (lane != 0) return;
    const int n = max(0, min(tile_n - g * CST_GROUP, CST_GROUP));
    const uint32_t bytes = ((uint32_t)n * 4u + 15u) & ~15u;
    if (bytes) {
      // the stage was read (generic proxy) by the whole warp before the __syncwarp that precedes this call
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar[g], bytes);
      tma_load_1d(ring[g], cigar + o0 + g * CST_GROUP, bytes, &bar[g]);
    } else {                                      // group behind the end of the op stream: complete the phase empty
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[g])) : "memory");
    }
  };

  int64_t tile = tile_of(idx);
  {
    const int64_t o0 = tile * CIG_TILE;
    const int tn = (int)min((int64_t)CIG_TILE, n_ops - o0);
#pragma unroll
    for (int g = 0; g < CST_NST - 1; g++) produce(o0, tn, g);
  }
  CstTile cur = describe(tile, tile_rec[tile]);
  int my_end = load_ends(cur);
  uint32_t phase = 0;                             // parity of the barriers for the current tile
  for (;;) {
    const int64_t idx_n = idx + n_warps;
    const bool more = idx_n < n_list;
    const int64_t tile_n1 = more ? tile_of(idx_n) : 0;
    int2 tr_n = make_int2(0, 0);
    if (more) tr_n = tile_rec[tile_n1];           // needed after group 1: a whole group of work hides it
    const int64_t o0_n = tile_n1 * CIG_TILE;
    const int tn_n = more ? (int)min((int64_t)CIG_TILE, n_ops - o0_n) : 0;
    CstTile nxt_t = cur;
    int my_end_n = CIG_TILE + 1;

    int r = 0;                                    // current record (local index), warp-uniform
    int nxt = __shfl_sync(0xffffffffu, my_end, 0);   // its end
    CigAcc acc;
    acc.clear();
#pragma unroll 1
    for (int g = 0; g < CIG_TILE / CST_GROUP; g++) {
      // three groups ahead: the last group of this tile, then the first three of the next one; the stage it
      // goes to was consumed in the previous round
      if (g == 0) produce(cur.o0, cur.tile_n, CST_NST - 1);
      else if (more) produce(o0_n, tn_n, g - 1);
      const int slot = g;
      mbar_wait(&bar[slot], phase);
      const int gbase = g * CST_GROUP;
      if (gbase < cur.tile_n) {
        uint4 q[4];
#pragma unroll
        for (int j = 0; j < 4; j++) q[j] = reinterpret_cast<const uint4*>(ring[slot])[j * 32 + lane];
        if (gbase + CST_GROUP > cur.tile_n) {     // last tile of the stream: absent ops read as 0 (M of length 0)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int p = gbase + j * 128 + lane * 4;
            if (p >= cur.tile_n) q[j].x = 0u;
            if (p + 1 >= cur.tile_n) q[j].y = 0u;
            if (p + 2 >= cur.tile_n) q[j].z = 0u;
            if (p + 3 >= cur.tile_n) q[j].w = 0u;
          }
        }
        uint32_t seen = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
          seen |= __funnelshift_l(0u, 1u, q[k].x) | __funnelshift_l(0u, 1u, q[k].y) | __funnelshift_l(0u, 1u, q[k].z) |
                  __funnelshift_l(0u, 1u, q[k].w);
        const bool rare = ((seen | (seen >> 16)) & 0xFE68u) != 0;    // N, H, P, B among my 16 ops (S is common)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int ibase = gbase + j * 128;
          const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
          if (nxt >= ibase + 128) {               // the whole round lies inside the current record (warp-uniform)
            if (rare) {
#pragma unroll
              for (int k = 0; k < 4; k++) acc.add(w[k]);
            } else {
              uint32_t all = 0, sc = 0;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const uint32_t c = w[k] & 15u, l = w[k] >> 4;
                all += l;
                acc.i += (c == 1u) ? l : 0u;
                acc.d += (c == 2u) ? l : 0u;
                sc += (c == 4u) ? l : 0u;
              }
              acc.tot += all - sc;
              acc.s += sc;
            }
            continue;
          }
          const int p = ibase + lane * 4;
          int done = ibase;                       // ops before `done` are accounted for
          while (nxt < ibase + 128) {             // a record ends inside this round (warp-uniform)
#pragma unroll
            for (int k = 0; k < 4; k++)
              if (p + k >= done && p + k < nxt) acc.add(w[k]);
            acc.warp_reduce();
            acc.flush_warp(stats + (int64_t)(cur.x + r) * 8, lane);
            acc.clear();
            done = nxt;
            r++;
            nxt = r < cur.n_loc ? __shfl_sync(0xffffffffu, my_end, r) : CIG_TILE + 1;
          }
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (p + k >= done) acc.add(w[k]);
        }
      }
      __syncwarp();                               // every lane is done with the stage before it is refilled
      if (g == 1 && more) nxt_t = describe(tile_n1, tr_n);
      if (g == 2 && more) my_end_n = load_ends(nxt_t);
    }
    if (r < cur.n_loc) {
      acc.warp_reduce();
      acc.flush_warp(stats + (int64_t)(cur.x + r) * 8, lane);
    }
    if (!more) break;
    idx = idx_n;
    cur = nxt_t;
    my_end = my_end_n;
    phase ^= 1u;
  }
}

// ================================================================================================
// K2  gate of the records the staged kernel did not see completely (GCI.py:153-168)
// ================================================================================================
// span_list (built at upload): records without ops, records cut by a tile border, records of streaming tiles
// and of pathological tiles.  list == NULL: all n records (every row of `stats` is valid then).
__global__ void gate_span_kernel(const int32_t* __restrict__ list, int64_t n, const uint32_t* __restrict__ stats,
                                 GateArgs gate) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int64_t r = list ? (int64_t)list[k] : k;
  const uint4 st = *reinterpret_cast<const uint4*>(stats + r * 8);
  gate_one(gate, r, st.x, st.y, st.z, st.w, stats[r * 8 + 4]);
}

__global__ void zero_rows_kernel(const int32_t* __restrict__ list, int64_t n, uint32_t* __restrict__ stats) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;      // two threads per 32-byte row
  if (k >= 2 * n) return;
  reinterpret_cast<uint4*>(stats + (int64_t)list[k >> 1] * 8)[k & 1] = make_uint4(0u, 0u, 0u, 0u);
}

// which records does the staged kernel gate itself?  Those with at least one op, all of them inside one tile,
// that tile being a staged one handled through shared memory.  Everything else goes to span_list.
__global__ void cigar_record_class_kernel(const uint64_t* __restrict__ off, int64_t n, const int2* __restrict__ tile_rec,
                                          int32_t* __restrict__ span_list, unsigned int* __restrict__ n_span) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint64_t a = off[r], b = off[r + 1];
  bool local = false;
  if (b > a) {
    const int64_t t0 = (int64_t)(a / CIG_TILE), t1 = (int64_t)((b - 1) / CIG_TILE);
    if (t0 == t1) {
      const int2 tr = tile_rec[t0];
      const int n_loc = tr.y - tr.x + 1;
      local = n_loc > CST_MAX_LOC && n_loc <= CIG_CAP;
    }
  }
  if (!local) span_list[atomicAdd(n_span, 1u)] = (int32_t)r;
}

__global__ void table_win_kernel(int64_t n, const uint32_t* __restrict__ read_id, uint32_t n_reads,
                                 const uint8_t* __restrict__ hq_in, long long* __restrict__ win,
                                 uint8_t* __restrict__ highq) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t q = read_id[i];
  if (q >= n_reads) return;
  atomicMax(win + q, (long long)i);
  if (hq_in && hq_in[i]) highq[q] = 1;
}

// ================================================================================================
// K4  cross-file join (GCI.py:272-301): one thread per read, files in join order
// ================================================================================================
struct JoinFile {
  const long long* win;
  const int32_t* ref_id;
  const int32_t* start;
  const int32_t* end;
  const int32_t* qlen;
};
struct JoinArgs {
  JoinFile f[GCI_MAX_FILES];
  int n_files;
};

// persistent grid: a warp walks 32 consecutive reads per round, so its depth-sum / survivor-count partials stay in
// registers (WarpSums) until the contig changes.
// NF = 1, 2: the common shapes (one file; BAM + PAF or two BAMs).  The kernel is a chain of dependent gathers
// (winner -> record -> fields), so a thread works on JOIN_ILP reads at once and issues every load of one level for
// all of them before it looks at any: the chain is paid once per JOIN_ILP reads instead of once per read.
// NF = 0: any number of files, one read at a time.
constexpr int JOIN_ILP = 4;
template <int NF>
__global__ void __launch_bounds__(256)
join_kernel(JoinArgs a, uint32_t n_reads, const uint8_t* __restrict__ highq, double op,
            int32_t* __restrict__ s_contig, int32_t* __restrict__ s_start, int32_t* __restrict__ s_end,
            unsigned long long* __restrict__ count, unsigned long long* __restrict__ err, BucketArgs bk) {
  __shared__ ContigCache cc;
  contig_cache_load(cc, bk);
  WarpSums ws;
  ws.init();
  if (NF == 0) {
    long long begin, end;
    cta_range(n_reads, blockDim.x, begin, end);
    for (long long base = begin; base < end; base += blockDim.x) {
      const uint32_t r = (uint32_t)(base + threadIdx.x);
      bool have = false;
      int32_t c = -1, s = 0, e = 0;
      if (r < n_reads) {
        long long k[GCI_MAX_FILES];
        for (int f = 0; f < a.n_files; f++) k[f] = a.f[f].win[r];
        have = join_one(a.n_files, k, [&](int f) {
          const uint32_t i = (uint32_t)(k[f] & 0xffffffffll);
          return JoinEntry{a.f[f].ref_id[i], a.f[f].start[i], a.f[f].end[i], f ? a.f[f].qlen[i] : 0};
        }, a.n_files > 1 && highq[r] != 0, op, r, err, c, s, e);
        s_contig[r] = have ? c : -1;
        s_start[r] = s;
        s_end[r] = e;
      }
      ws.add(bk, have ? c : -1, s, e, have);
    }
  } else {
    const uint32_t span = blockDim.x * JOIN_ILP;
    long long begin, end;
    cta_range(n_reads, span, begin, end);
    for (long long base = begin; base < end; base += span) {
      // read j of this thread: base + j * blockDim + tid (a warp still covers 32 consecutive reads per j)
      constexpr int NFX = NF > 0 ? NF : 1;
      long long k[JOIN_ILP][NFX];
      uint8_t hq[JOIN_ILP];
#pragma unroll
      for (int j = 0; j < JOIN_ILP; j++) {
        const uint32_t r = (uint32_t)(base + j * blockDim.x + threadIdx.x);
        const bool in = r < n_reads;
#pragma unroll
        for (int f = 0; f < NF; f++) k[j][f] = in ? a.f[f].win[r] : -1;
        hq[j] = (NF > 1 && in) ? highq[r] : 0;
      }
      JoinEntry x[JOIN_ILP][NFX];
#pragma unroll
      for (int j = 0; j < JOIN_ILP; j++)
#pragma unroll
        for (int f = 0; f < NF; f++) {
          const uint32_t i = k[j][f] >= 0 ? (uint32_t)(k[j][f] & 0xffffffffll) : 0u;   // absent: any valid entry
          x[j][f] = JoinEntry{a.f[f].ref_id[i], a.f[f].start[i], a.f[f].end[i], f ? a.f[f].qlen[i] : 0};
        }
#pragma unroll
      for (int j = 0; j < JOIN_ILP; j++) {
        const uint32_t r = (uint32_t)(base + j * blockDim.x + threadIdx.x);
        int32_t c, s, e;
        const bool have = join_one(NF, k[j], [&](int f) { return x[j][f]; }, hq[j] != 0, op, r, err, c, s, e);
        if (r < n_reads) {
          s_contig[r] = have ? c : -1;
          s_start[r] = s;
          s_end[r] = e;
        }
        ws.add(bk, have ? c : -1, s, e, have);
      }
    }
  }
  ws.flush(bk, count);
}

// ================================================================================================
// host drivers
// ================================================================================================
// one D2H of [error mask, first index, survivor count] through pinned memory, one sync
static int check_err(gci_ctx* ctx, const char* where, unsigned long long* count) {
  unsigned long long* h = (unsigned long long*)ctx->pinned(4 * sizeof(unsigned long long));
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_err.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                    ctx->stream));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (count) *count = h[2];
  if (h[0] == 0) return GCI_OK;
  const char* what = (h[0] & 1)   ? "record without NM tag (KeyError at GCI.py:163)"
                     : (h[0] & 2) ? "ZeroDivisionError at GCI.py:165 (clip ratio: no M/=/X/I/S bases)"
                     : (h[0] & 4) ? "ZeroDivisionError at GCI.py:165 (identity: no M/=/X/I/D bases)"
                     : (h[0] & 8) ? "ZeroDivisionError at GCI.py:292 (query_length 0 in the join)"
                     : (h[0] & 16) ? "ZeroDivisionError at GCI.py:231 (PAF alignment length 0)"
                                   : "ZeroDivisionError at GCI.py:247 (PAF query length 0)";
  if (h[0] & (64 | 128))
    return ctx->fail(GCI_E_CUDA, "%s: read-set exchange failed (%s)", where,
                     (h[0] & 64) ? "a peer rank did not arrive in time" : "an inbox overflowed");
  return ctx->fail(GCI_E_REFERENCE_RAISES, "%s: the reference raises here: %s [first index %llu]", where, what, h[1]);
}

static int reset_err(gci_ctx* ctx) {
  // every filter run starts from the marks that came with uploaded tables: the high-quality set depends on
  // the run's own --mq-cutoff, so marks of an earlier run on the same read set must not survive
  if (ctx->n_reads)
    GCI_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->highq.p, ctx->highq_base.p, ctx->n_reads, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
  if (ctx->shard.on && ctx->shard.n_home)     // marks of the home reads: PAF election and inbox merge set them
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->shard.hq_home.p, 0, ctx->shard.n_home, ctx->stream));
  GCI_TRY(ctx->ensure(ctx->d_err, 4 * sizeof(unsigned long long)));
  // [0] error mask = 0, [1] first offending index = ~0, [2] survivor count = 0
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_err.p, 0, 4 * sizeof(unsigned long long), ctx->stream));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_err.as<unsigned long long>() + 1, 0xff, sizeof(unsigned long long),
                                    ctx->stream));
  return GCI_OK;
}

int gci_index_bam(gci_ctx* ctx, BamFile& b) {
  b.n_dense = 0;
  b.n_span = b.n;                                 // no op stream: every record is gated from its (zero) row
  if (b.n == 0 || b.n_ops == 0) return GCI_OK;
  const int64_t n_tiles = (b.n_ops + CIG_TILE - 1) / CIG_TILE;
  GCI_TRY(ctx->ensure(b.tile_rec, sizeof(int2) * (size_t)(n_tiles + 1)));
  cigar_tile_index_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, ctx->stream>>>(
      b.cigar_off.as<uint64_t>(), b.n, b.tile_rec.as<int2>(), n_tiles, b.n_ops);
  GCI_LAUNCH_CHECK(ctx);
  // split the tiles between the staged (many records per tile) and the streaming kernel; the count of
  // staged tiles reaches the host with the synchronisation that ends the upload (gci_index_bam_finish)
  if (n_tiles >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "too many CIGAR op tiles");
  GCI_TRY(ctx->ensure(b.dense_list, sizeof(int32_t) * (size_t)n_tiles + 16));
  unsigned int* d_cnt = reinterpret_cast<unsigned int*>(b.dense_list.as<int32_t>() + n_tiles);   // [dense, stream, span, -]
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned int), ctx->stream));
  cigar_tile_class_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, ctx->stream>>>(
      b.tile_rec.as<int2>(), n_tiles, b.dense_list.as<int32_t>(), d_cnt);
  GCI_LAUNCH_CHECK(ctx);
  // records the staged kernel cannot gate on its own (see cigar_record_class_kernel)
  GCI_TRY(ctx->ensure(b.span_list, sizeof(int32_t) * (size_t)b.n));
  cigar_record_class_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, ctx->stream>>>(
      b.cigar_off.as<uint64_t>(), b.n, b.tile_rec.as<int2>(), b.span_list.as<int32_t>(), d_cnt + 2);
  GCI_LAUNCH_CHECK(ctx);
  unsigned int* h = (unsigned int*)ctx->pinned(4 * sizeof(unsigned int));
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_TRY(gci_d2h(ctx, h, d_cnt, 4 * sizeof(unsigned int)));
  b.n_dense = -1;                                 // pending
  return GCI_OK;
}

// after the stream synchronisation that follows gci_index_bam
void gci_index_bam_finish(gci_ctx* ctx, BamFile& b) {
  if (b.n_dense == -1) {
    const unsigned int* h = (const unsigned int*)ctx->pinned_scratch;
    b.n_dense = (int64_t)h[0];
    b.n_span = (int64_t)h[2];
  }
}

// the CIGAR kernels of one BAM upload.  gate != NULL: the staged kernel gates the records it sees completely
// and stores nothing for them; gate == NULL: every record's sums are stored (gci_fetch_cigar_stats).
static int run_cigar_kernels(gci_ctx* ctx, BamFile& b, const GateArgs* gate) {
  if (b.n_ops <= 0) return GCI_OK;
  const int64_t n_tiles = (b.n_ops + CIG_TILE - 1) / CIG_TILE;
  if (b.n_dense < 0 || b.n_dense > n_tiles) return ctx->fail(GCI_E_ARG, "internal: CIGAR tile index is not built");
  if (b.n_dense > 0) {
    const int32_t* list = b.n_dense == n_tiles ? nullptr : b.dense_list.as<int32_t>();
    static int lane_kernel = -1;
    if (lane_kernel < 0) {
      const char* e = getenv("GCI_CIGAR_LANE");
      lane_kernel = !(e && e[0] == '0');
    }
    if (lane_kernel) {
      if (gate) {
        cigar_lane_kernel<true><<<(unsigned)b.n_dense, CIGL_THREADS, 0, ctx->stream>>>(
            b.cigar.as<uint32_t>(), b.cigar_off.as<uint64_t>(), b.n, b.n_ops, b.tile_rec.as<int2>(), list,
            b.stats.as<uint32_t>(), *gate);
      } else {
        GateArgs none;
        memset(&none, 0, sizeof none);
        cigar_lane_kernel<false><<<(unsigned)b.n_dense, CIGL_THREADS, 0, ctx->stream>>>(
            b.cigar.as<uint32_t>(), b.cigar_off.as<uint64_t>(), b.n, b.n_ops, b.tile_rec.as<int2>(), list,
            b.stats.as<uint32_t>(), none);
      }
    } else if (gate) {
      cigar_stats_kernel<true><<<(unsigned)b.n_dense, CIG_THREADS, 0, ctx->stream>>>(
          b.cigar.as<uint32_t>(), b.cigar_off.as<uint64_t>(), b.n, b.n_ops, b.tile_rec.as<int2>(), list,
          b.stats.as<uint32_t>(), *gate);
    } else {
      GateArgs none;
      memset(&none, 0, sizeof none);
      cigar_stats_kernel<false><<<(unsigned)b.n_dense, CIG_THREADS, 0, ctx->stream>>>(
          b.cigar.as<uint32_t>(), b.cigar_off.as<uint64_t>(), b.n, b.n_ops, b.tile_rec.as<int2>(), list,
          b.stats.as<uint32_t>(), none);
    }
    GCI_LAUNCH_CHECK(ctx);
  }
  if (b.n_dense < n_tiles) {
    // persistent warps: resident CTAs per SM x SM count (fewer when the stream is short)
    static int per_sm = 0;
    if (per_sm <= 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cigar_stream_kernel, CST_WARPS * 32, 0) !=
                            cudaSuccess || per_sm < 1)) {
      cudaGetLastError();
      per_sm = 2;
    }
    const int64_t n_stream = n_tiles - b.n_dense;
    const int64_t grid = std::min<int64_t>((n_stream + CST_WARPS - 1) / CST_WARPS, (int64_t)ctx->sm_count * per_sm);
    cigar_stream_kernel<<<(unsigned)grid, CST_WARPS * 32, 0, ctx->stream>>>(
        b.cigar.as<uint32_t>(), b.cigar_off.as<uint64_t>(), b.n_ops, b.tile_rec.as<int2>(),
        b.n_dense == 0 ? nullptr : b.dense_list.as<int32_t>() + b.n_dense, n_stream, b.stats.as<uint32_t>());
    GCI_LAUNCH_CHECK(ctx);
  }
  return GCI_OK;
}

int gci_run_bam_leg(gci_ctx* ctx, int file_idx, int bam_idx, int32_t mq, int32_t mq_cutoff, double ip, double cp) {
  BamFile& b = ctx->bam[bam_idx];
  FileTable& ft = ctx->files[file_idx];
  const int64_t n = b.n;
  GCI_TRY(ctx->ensure(ft.win, sizeof(long long) * (size_t)std::max<uint32_t>(1, ctx->n_reads)));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ft.win.p, 0xff, sizeof(long long) * (size_t)ctx->n_reads, ctx->stream));
  GCI_TRY(ctx->ensure(b.stats, 32 * (size_t)std::max<int64_t>(1, n)));
  GCI_TRY(ctx->ensure(b.ref_end, 4 * (size_t)std::max<int64_t>(1, n)));
  if (n == 0) return GCI_OK;
  if (b.n_span < 0 || b.n_span > n) return ctx->fail(GCI_E_ARG, "internal: CIGAR record index is not built");
  GateArgs g;
  g.ref_id = b.ref_id.as<int32_t>(); g.ref_start = b.ref_start.as<int32_t>(); g.mapq = b.mapq.as<uint8_t>();
  g.flag = b.flag.as<uint16_t>(); g.nm = b.nm.as<int32_t>(); g.read_id = b.read_id.as<uint32_t>();
  g.selected = ctx->d_gate_sel.as<uint8_t>(); g.n_contigs = ctx->n_contigs; g.n_reads = ctx->n_reads;
  g.map_qual = mq; g.mq_cutoff = mq_cutoff; g.ip = ip; g.cp = cp;
  g.ref_end = b.ref_end.as<int32_t>(); g.win = ft.win.as<long long>(); g.highq = ctx->highq.as<uint8_t>();
  g.err = ctx->d_err.as<unsigned long long>();
  ctx->stage_begin(GCI_ST_CIGAR);
  // only the rows of the records gated from HBM (span_list) are ever read: zero those (all of them when they
  // are a large share, as for ONT)
  const bool all_rows = b.n_span * 4 > n;
  if (all_rows) {
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(b.stats.p, 0, 32 * (size_t)n, ctx->stream));
  } else if (b.n_span > 0) {
    zero_rows_kernel<<<(unsigned)((2 * b.n_span + 255) / 256), 256, 0, ctx->stream>>>(
        b.span_list.as<int32_t>(), b.n_span, b.stats.as<uint32_t>());
    GCI_LAUNCH_CHECK(ctx);
  }
  GCI_TRY(run_cigar_kernels(ctx, b, &g));
  ctx->stage_end();
  ctx->stage_begin(GCI_ST_GATE);
  if (b.n_span > 0) {
    gate_span_kernel<<<(unsigned)((b.n_span + 255) / 256), 256, 0, ctx->stream>>>(
        b.n_span == n ? nullptr : b.span_list.as<int32_t>(), b.n_span, b.stats.as<uint32_t>(), g);
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  return GCI_OK;
}

static int install_table(gci_ctx* ctx, FileTable& f, int64_t n, const uint32_t* read_id, const int32_t* ref_id,
                         const int32_t* start, const int32_t* end, const int32_t* qlen, const uint8_t* highq) {
  f.kind = 1;
  f.paf = -1;
  f.n = n;
  DevBuf &d_read = ctx->tmp[4], &d_hq = ctx->tmp[5];
  ctx->stage_begin(GCI_ST_H2D);
  GCI_TRY(gci_h2d(ctx, d_read, read_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, f.ref_id, ref_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, f.start, start, 4 * n));
  GCI_TRY(gci_h2d(ctx, f.end, end, 4 * n));
  GCI_TRY(gci_h2d(ctx, f.qlen, qlen, 4 * n));
  if (highq) GCI_TRY(gci_h2d(ctx, d_hq, highq, n));
  ctx->stage_end();
  GCI_TRY(ctx->ensure(f.win, sizeof(long long) * (size_t)std::max<uint32_t>(1, ctx->n_reads)));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(f.win.p, 0xff, sizeof(long long) * (size_t)ctx->n_reads, ctx->stream));
  if (n) {
    table_win_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        n, d_read.as<uint32_t>(), ctx->n_reads, highq ? d_hq.as<uint8_t>() : nullptr, f.win.as<long long>(),
        ctx->highq_base.as<uint8_t>());
    GCI_LAUNCH_CHECK(ctx);
  }
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

// ================================================================================================
// K3  PAF leg on the GPU (GCI.py:211-254)
// ================================================================================================
// Lines of all PAF files of the read set share one index space (global line id = lines of the earlier files + line
// number), which is also their arrival order.  Per line: contig selected, identity = nmatch / alnlen (fp64), keep iff
// mapq >= -mq and identity >= -ip.  Nothing is compacted: the kept lines are grouped by read as a CSR over line ids
// (count -> scan -> fill) and the election reads the original columns.  A read with ONE kept line (almost all of
// them) is its own result; the others are elected like the reference does.  The reference's `synteny` dict is
// created once per filter() call (:214), so the table of PAF file f is elected over the lines of files 0..f.
struct PafCols {
  const uint32_t* read_id;
  const int32_t *qlen, *qstart, *qend, *ref_id, *tstart, *tend, *nmatch, *alnlen, *mapq;
};
struct PafSet {
  PafCols f[GCI_MAX_FILES];
  const int4* rows[GCI_MAX_FILES];    // packed lines of every file (PafFile::rows)
  int64_t off[GCI_MAX_FILES + 1];     // first global line id of every file
  int n;
};
struct PafLine { int32_t ref, qlen, q0, q1, t0, t1; double ident; };

__device__ __forceinline__ PafLine paf_line(const PafSet& ps, int32_t g) {
  int k = 0;
  while (k + 1 < ps.n && (int64_t)g >= ps.off[k + 1]) k++;
  const int4* row = ps.rows[k] + 2 * ((int64_t)g - ps.off[k]);
  const int4 a = row[0], b = row[1];                                   // one 32-byte sector
  PafLine l;
  l.ref = a.x; l.qlen = a.y; l.q0 = a.z; l.q1 = a.w; l.t0 = b.x; l.t1 = b.y;
  l.ident = (double)b.z / (double)b.w;                                 // :231 (kept lines have alnlen != 0)
  return l;
}

__global__ void paf_pack_rows_kernel(int64_t n, PafCols p, int4* __restrict__ rows) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  rows[2 * i] = make_int4(p.ref_id[i], p.qlen[i], p.qstart[i], p.qend[i]);
  rows[2 * i + 1] = make_int4(p.tstart[i], p.tend[i], p.nmatch[i], p.alnlen[i]);
}

int gci_pack_paf_rows(gci_ctx* ctx, PafFile& pf) {
  GCI_TRY(ctx->ensure(pf.rows, 32 * (size_t)std::max<int64_t>(1, pf.n)));
  if (pf.n == 0) return GCI_OK;
  const PafCols pc{pf.read_id.as<uint32_t>(), pf.qlen.as<int32_t>(), pf.qstart.as<int32_t>(), pf.qend.as<int32_t>(),
                   pf.ref_id.as<int32_t>(), pf.tstart.as<int32_t>(), pf.tend.as<int32_t>(), pf.nmatch.as<int32_t>(),
                   pf.alnlen.as<int32_t>(), pf.mapq.as<int32_t>()};
  paf_pack_rows_kernel<<<(unsigned)((pf.n + 255) / 256), 256, 0, ctx->stream>>>(pf.n, pc, pf.rows.as<int4>());
  GCI_LAUNCH_CHECK(ctx);
  return GCI_OK;
}

__global__ void paf_mark_kernel(int64_t n, PafCols p, int64_t line0, const uint8_t* __restrict__ selected,
                                int32_t n_contigs, uint32_t n_reads, int32_t map_qual, int32_t mq_cutoff, double ip,
                                uint8_t* __restrict__ keep, int32_t* __restrict__ cnt, uint8_t* __restrict__ highq,
                                unsigned long long* __restrict__ err) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t k = 0;
  const int32_t t = p.ref_id[i];
  const int32_t al = p.alnlen[i], nm = p.nmatch[i], mq = p.mapq[i];
  const uint32_t q = p.read_id[i];
  if (t >= 0 && t < n_contigs && selected[t]) {                         // :220
    if (al == 0) {                                                      // ZeroDivisionError :231
      atomicOr(err, 16ull);
      atomicMin(err + 1, (unsigned long long)i);
    } else {
      const double ident = (double)nm / (double)al;                    // :231
      if (mq >= map_qual && ident >= ip && q < n_reads) {               // :232
        k = 1;
        atomicAdd(&cnt[q], 1);
        if (mq >= mq_cutoff) highq[q] = 1;                              // :238
      }
    }
  }
  keep[line0 + i] = k;
}

__global__ void paf_fill_kernel(int64_t n, const uint32_t* __restrict__ read_id, int64_t line0,
                                const uint8_t* __restrict__ keep, const int32_t* __restrict__ off,
                                int32_t* __restrict__ cur, int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !keep[line0 + i]) return;
  const uint32_t q = read_id[i];
  idx[off[q] + atomicAdd(&cur[q], 1)] = (int32_t)(line0 + i);
}

// GCI.py:64-96 over the lines get(0..g) that sit on contig `ref`: walk the (lo, hi) pairs in sorted
// order by repeated selection of the next smallest (lo, hi, position) triple (groups are tiny), merge
// touching / overlapping blocks, return the total merged length and the longest block (first on ties).
struct Blocks { long long total; int32_t lo, hi; };
template <bool TARGET, class Get>
__device__ Blocks merged_blocks(const Get& get, int g, int32_t ref) {
  Blocks r{0, 0, 0};
  long long best_len = -1;
  bool have_cur = false, have_last = false;
  int32_t cur_lo = 0, cur_hi = 0, last_lo = 0, last_hi = 0;
  int last_pos = -1;
  for (;;) {
    // next pair in (lo, hi, position) order after (last_lo, last_hi, last_pos)
    int pick = -1;
    int32_t p_lo = 0, p_hi = 0;
    for (int j = 0; j < g; j++) {
      const PafLine l = get(j);
      if (l.ref != ref) continue;
      const int32_t lo = TARGET ? l.t0 : l.q0, hi = TARGET ? l.t1 : l.q1;
      if (have_last) {
        const bool after = lo > last_lo || (lo == last_lo && (hi > last_hi || (hi == last_hi && j > last_pos)));
        if (!after) continue;
      }
      if (pick < 0 || lo < p_lo || (lo == p_lo && (hi < p_hi))) { pick = j; p_lo = lo; p_hi = hi; }
    }
    if (pick < 0) break;
    have_last = true; last_lo = p_lo; last_hi = p_hi; last_pos = pick;
    if (!have_cur) { have_cur = true; cur_lo = p_lo; cur_hi = p_hi; continue; }
    if (cur_hi >= p_lo) {
      if (cur_hi < p_hi) cur_hi = p_hi;
    } else {
      const long long len = (long long)cur_hi - cur_lo;
      r.total += len;
      if (len > best_len) { best_len = len; r.lo = cur_lo; r.hi = cur_hi; }
      cur_lo = p_lo; cur_hi = p_hi;
    }
  }
  if (have_cur) {
    const long long len = (long long)cur_hi - cur_lo;
    r.total += len;
    if (len > best_len) { r.lo = cur_lo; r.hi = cur_hi; }
  }
  return r;
}

struct PafTableOut {
  int32_t *ref, *start, *end, *qlen;
  long long* win;
};

// the election of one read over its g >= 2 kept lines, get(j) = j-th line in arrival order (GCI.py:241-254)
template <class Get>
__device__ void paf_elect_read(uint32_t r, int g, const Get& get, const int32_t* __restrict__ name_rank,
                               const PafTableOut& o, unsigned long long* __restrict__ err) {
  bool have = false;
  double best_score = 0;
  int32_t best_rank = 0, b_ref = 0, b_s = 0, b_e = 0, b_q = 0;
  for (int i = 0; i < g; i++) {
    const PafLine li = get(i);
    const int32_t ref = li.ref;
    bool seen = false;
    for (int j = 0; j < i && !seen; j++) seen = get(j).ref == ref;
    if (seen) continue;
    double sum = 0;                          // Python sum() starts from int 0; 0 + x is exact
    int cnt = 0;
    for (int j = i; j < g; j++) {
      const PafLine lj = get(j);
      if (lj.ref == ref) { sum = sum + lj.ident; cnt++; }
    }
    const int32_t qlen = li.qlen;            // alns[0][0], :246
    if (qlen == 0) {                         // ZeroDivisionError :247
      atomicOr(err, 32ull);
      atomicMin(err + 1, (unsigned long long)r);
      o.win[r] = -1;
      return;
    }
    const Blocks bq = merged_blocks<false>(get, g, ref);
    const Blocks bt = merged_blocks<true>(get, g, ref);
    const double rate = (double)bq.total / (double)qlen;               // :247
    const double score = (sum / (double)cnt) * rate;                   // :248-249
    const int32_t rank = name_rank[ref];
    if (!have || score > best_score || (score == best_score && rank > best_rank)) {   // :252
      have = true;
      best_score = score; best_rank = rank;
      b_ref = ref; b_s = bt.lo; b_e = bt.hi; b_q = qlen;
    }
  }
  o.ref[r] = b_ref; o.start[r] = b_s; o.end[r] = b_e; o.qlen[r] = b_q;
  o.win[r] = (long long)r;                   // table entries are indexed by read id
}

// one thread per read: no kept line -> absent; ONE kept line (almost every read) -> that line is the result;
// more -> the read is left for paf_elect_multi_kernel (its slot of `multi` says so), so no warp waits for a lane
// walking a group
constexpr int PAF_ELECT_THREADS = 128;
__global__ void __launch_bounds__(PAF_ELECT_THREADS)
paf_elect_kernel(uint32_t n_reads, PafSet ps, const int32_t* __restrict__ off, const int32_t* __restrict__ idx,
                 PafTableOut o, uint8_t* __restrict__ multi, unsigned long long* __restrict__ err) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const int32_t a0 = off[r], g = off[r + 1] - a0;
  multi[r] = g >= 2;
  if (g == 0) o.win[r] = -1;
  if (g != 1) return;
  const int32_t line = idx[a0];
  int k = 0;
  while (k + 1 < ps.n && (int64_t)line >= ps.off[k + 1]) k++;
  const int4* row = ps.rows[k] + 2 * ((int64_t)line - ps.off[k]);
  const int4 a = row[0], b = row[1];         // (ref, qlen, q0, q1) (t0, t1, nmatch, alnlen): one sector
  if (a.y == 0) {                            // ZeroDivisionError :247
    atomicOr(err, 32ull);
    atomicMin(err + 1, (unsigned long long)r);
    o.win[r] = -1;
    return;
  }
  o.ref[r] = a.x; o.start[r] = b.x; o.end[r] = b.y; o.qlen[r] = a.y;
  o.win[r] = (long long)r;
}

// the reads with several kept lines: a warp scans 32 marks at a time and its lanes take the marked reads one each
// (compacted inside the warp: no lane idles while its neighbour walks a group)
constexpr int PAF_LOCAL = 8;                 // lines of one read kept in registers / local memory
constexpr int PAF_MULTI_CHUNK = 256;         // reads per warp round of paf_elect_multi_kernel
__global__ void __launch_bounds__(128)
paf_elect_multi_kernel(uint32_t n_reads, PafSet ps, const int32_t* __restrict__ off, int32_t* __restrict__ idx,
                       const int32_t* __restrict__ name_rank, PafTableOut o, const uint8_t* __restrict__ multi,
                       unsigned long long* __restrict__ err) {
  __shared__ uint32_t s_list[4][PAF_MULTI_CHUNK];
  const int lane = threadIdx.x & 31;
  const uint32_t n_warps = gridDim.x * (blockDim.x >> 5);
  constexpr uint32_t chunk = PAF_MULTI_CHUNK; // reads per warp round: 8 marks per lane
  constexpr int PER_LANE = PAF_MULTI_CHUNK / 32;
  // every warp walks one contiguous range of reads, `chunk` reads per round
  const uint32_t wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint32_t per_warp = ((n_reads + n_warps - 1) / n_warps + chunk - 1) / chunk * chunk;
  const uint32_t w_begin = (uint32_t)min((unsigned long long)n_reads, (unsigned long long)wid * per_warp);
  const uint32_t w_end = (uint32_t)min((unsigned long long)n_reads, (unsigned long long)w_begin + per_warp);
  for (uint32_t base = w_begin; base < w_end; base += chunk) {
    // lane l looks at reads base + 8 l .. + 7 (one 8-byte load when the range is complete)
    uint32_t mask = 0;
    const uint32_t r0 = base + lane * PER_LANE;
    static_assert(PER_LANE == 8, "one 8-byte load per lane");
    if (r0 + PER_LANE <= n_reads) {
      const unsigned long long w = *reinterpret_cast<const unsigned long long*>(multi + r0);
#pragma unroll
      for (int b8 = 0; b8 < PER_LANE; b8++) mask |= (uint32_t)((w >> (8 * b8)) & 1ull) << b8;
    } else {
      for (int k = 0; k < PER_LANE; k++)
        if (r0 + k < n_reads && multi[r0 + k]) mask |= 1u << k;
    }
    // hand the marked reads out: the lanes list them in the warp's shared buffer, then take one each per round
    const int cnt = __popc(mask);
    const int incl = warp_incl_scan(cnt, lane);
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t* list = s_list[threadIdx.x >> 5];
    int pos = incl - cnt;
    while (mask) {
      list[pos++] = r0 + (__ffs(mask) - 1);
      mask &= mask - 1;
    }
    __syncwarp();
    for (int t = lane; t < total; t += 32) {
      const uint32_t r = list[t];
      const int32_t a0 = off[r], g = off[r + 1] - a0;
      int32_t* seg = idx + a0;
      for (int i = 1; i < g; i++) {          // insertion sort by line id = arrival order (file order, :237)
        const int32_t e = seg[i];
        int j = i - 1;
        while (j >= 0 && seg[j] > e) { seg[j + 1] = seg[j]; j--; }
        seg[j + 1] = e;
      }
      if (g <= PAF_LOCAL) {
        PafLine L[PAF_LOCAL];
        for (int j = 0; j < g; j++) L[j] = paf_line(ps, seg[j]);
        paf_elect_read(r, g, [&L](int j) { return L[j]; }, name_rank, o, err);
      } else {
        paf_elect_read(r, g, [&ps, seg](int j) { return paf_line(ps, seg[j]); }, name_rank, o, err);
      }
    }
    __syncwarp();                            // everybody is done with the list before the next round rewrites it
  }
}

int gci_run_paf_legs(gci_ctx* ctx, int32_t mq, int32_t mq_cutoff, double ip) {
  bool any = false;
  for (size_t fi = 0; fi < ctx->n_files; fi++) any |= ctx->files[fi].paf >= 0;
  if (!any) return GCI_OK;
  if ((int32_t)ctx->name_rank.size() != ctx->n_contigs)
    return ctx->fail(GCI_E_ARG, "gci_set_name_rank must be called before filtering PAF files");
  PafSet ps;
  memset(&ps, 0, sizeof ps);
  std::vector<int> file_of;                  // FileTable index of PAF file k (join order)
  for (size_t fi = 0; fi < ctx->n_files; fi++) {
    FileTable& ft = ctx->files[fi];
    if (ft.paf < 0) continue;
    const PafFile& pf = ctx->paf[ft.paf];
    const int k = ps.n++;
    ps.f[k] = PafCols{pf.read_id.as<uint32_t>(), pf.qlen.as<int32_t>(), pf.qstart.as<int32_t>(), pf.qend.as<int32_t>(),
                      pf.ref_id.as<int32_t>(), pf.tstart.as<int32_t>(), pf.tend.as<int32_t>(), pf.nmatch.as<int32_t>(),
                      pf.alnlen.as<int32_t>(), pf.mapq.as<int32_t>()};
    ps.rows[k] = pf.rows.as<int4>();
    ps.off[k + 1] = ps.off[k] + pf.n;
    file_of.push_back((int)fi);
  }
  const int64_t total_lines = ps.off[ps.n];
  if (total_lines >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "more than 2^31 PAF lines in one read set");
  const size_t cap = (size_t)std::max<int64_t>(1, total_lines);
  // sharded read sets upload PAF lines with home-local read ids (id / world): the tables are indexed by those
  const uint32_t n_reads = ctx->shard.on ? ctx->shard.n_home : ctx->n_reads;
  uint8_t* highq = ctx->shard.on ? ctx->shard.hq_home.as<uint8_t>() : ctx->highq.as<uint8_t>();
  const size_t nr = std::max<uint32_t>(1, n_reads);
  DevBuf &cnt = ctx->tmp[6], &off = ctx->tmp[7], &idx = ctx->tmp[8], &keep = ctx->paf_keep, &multi_buf = ctx->tmp[9];
  GCI_TRY(ctx->ensure(multi_buf, 4 * (nr + 1)));
  GCI_TRY(ctx->ensure(cnt, 4 * (nr + 1) * 2));     // counts | fill cursors
  GCI_TRY(ctx->ensure(off, 4 * (nr + 1)));
  GCI_TRY(ctx->ensure(idx, 4 * cap));
  GCI_TRY(ctx->ensure(keep, cap));
  int32_t* d_cnt = cnt.as<int32_t>();
  int32_t* d_cur = d_cnt + (nr + 1);
  ctx->stage_begin(GCI_ST_PAF);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_cnt, 0, 4 * (nr + 1), ctx->stream));
  for (int k = 0; k < ps.n; k++) {
    FileTable& ft = ctx->files[file_of[k]];
    const int64_t n = ps.off[k + 1] - ps.off[k];
    if (n) {
      paf_mark_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
          n, ps.f[k], ps.off[k], ctx->d_gate_sel.as<uint8_t>(), ctx->n_contigs, n_reads, mq, mq_cutoff, ip,
          keep.as<uint8_t>(), d_cnt, highq, ctx->d_err.as<unsigned long long>());
      GCI_LAUNCH_CHECK(ctx);
    }
    // group the kept lines of files 0..k by read: CSR over read ids (the counts accumulate across the files)
    GCI_TRY(gci_exclusive_scan_i32(ctx, d_cnt, off.as<int32_t>(), (int64_t)n_reads + 1));
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_cur, 0, 4 * (nr + 1), ctx->stream));
    for (int j = 0; j <= k; j++) {
      const int64_t nj = ps.off[j + 1] - ps.off[j];
      if (!nj) continue;
      paf_fill_kernel<<<(unsigned)((nj + 255) / 256), 256, 0, ctx->stream>>>(
          nj, ps.f[j].read_id, ps.off[j], keep.as<uint8_t>(), off.as<int32_t>(), d_cur, idx.as<int32_t>());
      GCI_LAUNCH_CHECK(ctx);
    }
    // the file's table: one entry per read id
    ft.kind = 1;
    ft.n = n_reads;
    GCI_TRY(ctx->ensure(ft.ref_id, 4 * nr)); GCI_TRY(ctx->ensure(ft.start, 4 * nr));
    GCI_TRY(ctx->ensure(ft.end, 4 * nr)); GCI_TRY(ctx->ensure(ft.qlen, 4 * nr));
    GCI_TRY(ctx->ensure(ft.win, 8 * nr));
    if (n_reads) {
      PafSet upto = ps;
      upto.n = k + 1;
      PafTableOut o{ft.ref_id.as<int32_t>(), ft.start.as<int32_t>(), ft.end.as<int32_t>(), ft.qlen.as<int32_t>(),
                    ft.win.as<long long>()};
      uint8_t* multi = multi_buf.as<uint8_t>();
      paf_elect_kernel<<<(n_reads + PAF_ELECT_THREADS - 1) / PAF_ELECT_THREADS, PAF_ELECT_THREADS, 0, ctx->stream>>>(
          n_reads, upto, off.as<int32_t>(), idx.as<int32_t>(), o, multi, ctx->d_err.as<unsigned long long>());
      GCI_LAUNCH_CHECK(ctx);
      const unsigned mgrid = (unsigned)std::min<int64_t>(((int64_t)n_reads + 4 * PAF_MULTI_CHUNK - 1) / (4 * PAF_MULTI_CHUNK),
                                                         (int64_t)ctx->sm_count * 16);
      paf_elect_multi_kernel<<<mgrid, 128, 0, ctx->stream>>>(
          n_reads, upto, off.as<int32_t>(), idx.as<int32_t>(), ctx->d_name_rank.as<int32_t>(), o, multi,
          ctx->d_err.as<unsigned long long>());
      GCI_LAUNCH_CHECK(ctx);
    }
  }
  ctx->stage_end();
  return GCI_OK;
}

int gci_run_join(gci_ctx* ctx, double op) { return gci_run_join_counting(ctx, op, -1, 0); }

// track >= 0: also count the depth events of (track, flank_len) into the tile table (gci_depth_enqueue then
// skips its own counting pass)
int gci_run_join_counting(gci_ctx* ctx, double op, int32_t track, int32_t flank_len) {
  ctx->counted_track = -1;
  BucketArgs bk;
  memset(&bk, 0, sizeof bk);
  if (track >= 0) GCI_TRY(gci_depth_prepare(ctx, track, flank_len, &bk));
  JoinArgs a;
  memset(&a, 0, sizeof a);
  a.n_files = (int)ctx->n_files;
  for (int i = 0; i < a.n_files; i++) {
    FileTable& f = ctx->files[i];
    a.f[i].win = f.win.as<long long>();
    if (f.kind == 0) {
      BamFile& b = ctx->bam[f.src];
      a.f[i].ref_id = b.ref_id.as<int32_t>();
      a.f[i].start = b.ref_start.as<int32_t>();
      a.f[i].end = b.ref_end.as<int32_t>();
      a.f[i].qlen = b.qlen.as<int32_t>();
    } else {
      a.f[i].ref_id = f.ref_id.as<int32_t>();
      a.f[i].start = f.start.as<int32_t>();
      a.f[i].end = f.end.as<int32_t>();
      a.f[i].qlen = f.qlen.as<int32_t>();
    }
  }
  const size_t nr = std::max<uint32_t>(1, ctx->n_reads);
  GCI_TRY(ctx->ensure(ctx->surv_contig, 4 * nr));
  GCI_TRY(ctx->ensure(ctx->surv_start, 4 * nr));
  GCI_TRY(ctx->ensure(ctx->surv_end, 4 * nr));
  unsigned long long* cnt = ctx->d_err.as<unsigned long long>() + 2;
  ctx->stage_begin(GCI_ST_JOIN);
  if (ctx->n_reads) {
    // an empty file table (no rows) still has a win array; entry 0 of its columns must be readable for the
    // load-everything-first path, so files without rows take the generic kernel
    bool small = a.n_files <= 2;
    for (int i = 0; i < a.n_files; i++) small = small && ctx->files[i].n > 0;
    const int64_t per_cta = small ? 256 * JOIN_ILP : 256;
    const unsigned grid = (unsigned)std::min<int64_t>(((int64_t)ctx->n_reads + per_cta - 1) / per_cta, (int64_t)ctx->sm_count * 8);
    int32_t *sc = ctx->surv_contig.as<int32_t>(), *ss = ctx->surv_start.as<int32_t>(), *se = ctx->surv_end.as<int32_t>();
    unsigned long long* d_err = ctx->d_err.as<unsigned long long>();
    const uint8_t* hq = ctx->highq.as<uint8_t>();
    if (small && a.n_files == 1) join_kernel<1><<<grid, 256, 0, ctx->stream>>>(a, ctx->n_reads, hq, op, sc, ss, se, cnt, d_err, bk);
    else if (small) join_kernel<2><<<grid, 256, 0, ctx->stream>>>(a, ctx->n_reads, hq, op, sc, ss, se, cnt, d_err, bk);
    else join_kernel<0><<<grid, 256, 0, ctx->stream>>>(a, ctx->n_reads, hq, op, sc, ss, se, cnt, d_err, bk);
    GCI_LAUNCH_CHECK(ctx);
  }
  ctx->stage_end();
  if (track >= 0) {
    ctx->counted_track = track;
    ctx->counted_flank = flank_len;
  }
  return GCI_OK;
}

// ordered compaction helpers (fetch paths; not on the timed path)
__global__ void mark_survivors_kernel(uint32_t n, const int32_t* __restrict__ contig, int32_t* __restrict__ mark) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) mark[r] = contig[r] >= 0 ? 1 : 0;
}

__global__ void mark_present_kernel(uint32_t n, const long long* __restrict__ win, int32_t* __restrict__ mark) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) mark[r] = win[r] >= 0 ? 1 : 0;
}

__global__ void gather_survivors_kernel(uint32_t n, const int32_t* __restrict__ mark, const int32_t* __restrict__ pos,
                                        const int32_t* __restrict__ c, const int32_t* __restrict__ s,
                                        const int32_t* __restrict__ e, uint32_t* __restrict__ o_read,
                                        int32_t* __restrict__ o_c, int32_t* __restrict__ o_s, int32_t* __restrict__ o_e) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || !mark[r]) return;
  const int32_t p = pos[r];
  o_read[p] = r; o_c[p] = c[r]; o_s[p] = s[r]; o_e[p] = e[r];
}

__global__ void gather_table_kernel(uint32_t n, const int32_t* __restrict__ mark, const int32_t* __restrict__ pos,
                                    JoinFile f, const uint8_t* __restrict__ highq, uint32_t* __restrict__ o_read,
                                    int32_t* __restrict__ o_c, int32_t* __restrict__ o_s, int32_t* __restrict__ o_e,
                                    int32_t* __restrict__ o_q, uint8_t* __restrict__ o_h) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || !mark[r]) return;
  const int32_t p = pos[r];
  const uint32_t i = (uint32_t)(f.win[r] & 0xffffffffll);
  o_read[p] = r; o_c[p] = f.ref_id[i]; o_s[p] = f.start[i]; o_e[p] = f.end[i]; o_q[p] = f.qlen[i];
  o_h[p] = highq[r];
}

extern "C" {

int gci_upload_table(gci_ctx* ctx, int64_t n, const uint32_t* read_id, const int32_t* ref_id, const int32_t* start,
                     const int32_t* end, const int32_t* qlen, const uint8_t* highq) {
  if (!ctx || n < 0) return GCI_E_ARG;
  if (n && (!read_id || !ref_id || !start || !end || !qlen)) return GCI_E_ARG;
  if ((int)ctx->n_files >= GCI_MAX_FILES) return ctx->fail(GCI_E_ARG, "too many files");
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  if (ctx->n_files == ctx->files.size()) ctx->files.emplace_back();
  ctx->filtered = false;
  return install_table(ctx, ctx->files[ctx->n_files++], n, read_id, ref_id, start, end, qlen, highq);
}

int gci_filter(gci_ctx* ctx, int32_t map_qual, int32_t mq_cutoff, double iden_percent, double clip_percent,
               double ovlp_percent, int64_t* n_survivors) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  if (ctx->n_files == 0) return ctx->fail(GCI_E_ARG, "gci_filter: no files uploaded");
  if (ctx->shard.on) return ctx->fail(GCI_E_ARG, "gci_filter: a sharded read set runs through gci_pipeline");
  GCI_TRY(reset_err(ctx));
  // PAF election first (files join in upload order); tables uploaded by the caller are already final
  GCI_TRY(gci_run_paf_legs(ctx, map_qual, mq_cutoff, iden_percent));
  for (size_t i = 0; i < ctx->n_files; i++)
    if (ctx->files[i].kind == 0)
      GCI_TRY(gci_run_bam_leg(ctx, (int)i, ctx->files[i].src, map_qual, mq_cutoff, iden_percent, clip_percent));
  GCI_TRY(gci_run_join(ctx, ovlp_percent));
  unsigned long long cnt = 0;
  GCI_TRY(check_err(ctx, "gci_filter", &cnt));
  ctx->n_survivors = (int64_t)cnt;
  ctx->filtered = true;
  if (n_survivors) *n_survivors = ctx->n_survivors;
  return GCI_OK;
}

int gci_fetch_cigar_stats(gci_ctx* ctx, int32_t bam, int64_t n_records, uint32_t* stats, int32_t* ref_end) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->filtered) return ctx->fail(GCI_E_ARG, "gci_fetch_cigar_stats before gci_filter");
  if (bam < 0 || (size_t)bam >= ctx->n_bam) return ctx->fail(GCI_E_ARG, "no BAM upload %d", bam);
  BamFile& b = ctx->bam[bam];
  if (n_records != b.n) return ctx->fail(GCI_E_ARG, "BAM upload %d holds %lld records", bam, (long long)b.n);
  if (b.n == 0) return GCI_OK;
  if (stats) {
    // the filter keeps the sums of most records on chip: run the CIGAR kernels once more, storing every row
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(b.stats.p, 0, 32 * (size_t)b.n, ctx->stream));
    GCI_TRY(run_cigar_kernels(ctx, b, nullptr));
    // device rows are 8 words wide (one 32-byte sector per record); the caller gets the 5 used ones
    uint32_t* h = (uint32_t*)ctx->pinned(32 * (size_t)b.n);
    if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
    GCI_TRY(gci_d2h(ctx, h, b.stats.p, 32 * (size_t)b.n));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < b.n; r++)
      for (int k = 0; k < 5; k++) stats[r * 5 + k] = h[r * 8 + k];
  }
  if (ref_end) {
    GCI_TRY(gci_d2h(ctx, ref_end, b.ref_end.p, 4 * (size_t)b.n));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return GCI_OK;
}

int gci_fetch_survivors(gci_ctx* ctx, int64_t cap, uint32_t* read_id, int32_t* contig, int32_t* start, int32_t* end,
                        int64_t* n) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->filtered) return ctx->fail(GCI_E_ARG, "gci_fetch_survivors before gci_filter");
  if (ctx->shard.on) return ctx->fail(GCI_E_ARG, "gci_fetch_survivors: not available for a sharded read set");
  if (n) *n = ctx->n_survivors;
  if (!read_id && !contig && !start && !end) return GCI_OK;
  if (cap < ctx->n_survivors) return ctx->fail(GCI_E_ARG, "survivor buffer too small");
  const uint32_t nr = ctx->n_reads;
  const int64_t ns = ctx->n_survivors;
  if (ns == 0) return GCI_OK;
  DevBuf mark, pos, o_r, o_c, o_s, o_e;
  GCI_TRY(ctx->ensure(mark, 4 * (size_t)nr));
  GCI_TRY(ctx->ensure(pos, 4 * (size_t)nr));
  for (DevBuf* d : {&o_r, &o_c, &o_s, &o_e}) GCI_TRY(ctx->ensure(*d, 4 * (size_t)ns));
  mark_survivors_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(nr, ctx->surv_contig.as<int32_t>(), mark.as<int32_t>());
  GCI_LAUNCH_CHECK(ctx);
  GCI_TRY(gci_exclusive_scan_i32(ctx, mark.as<int32_t>(), pos.as<int32_t>(), nr));
  gather_survivors_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(
      nr, mark.as<int32_t>(), pos.as<int32_t>(), ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(),
      ctx->surv_end.as<int32_t>(), o_r.as<uint32_t>(), o_c.as<int32_t>(), o_s.as<int32_t>(), o_e.as<int32_t>());
  GCI_LAUNCH_CHECK(ctx);
  if (read_id) GCI_TRY(gci_d2h(ctx, read_id, o_r.p, 4 * ns));
  if (contig) GCI_TRY(gci_d2h(ctx, contig, o_c.p, 4 * ns));
  if (start) GCI_TRY(gci_d2h(ctx, start, o_s.p, 4 * ns));
  if (end) GCI_TRY(gci_d2h(ctx, end, o_e.p, 4 * ns));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (DevBuf* d : {&mark, &pos, &o_r, &o_c, &o_s, &o_e}) ctx->release(*d);
  return GCI_OK;
}

int gci_fetch_file_table(gci_ctx* ctx, int32_t file, int64_t cap, uint32_t* read_id, int32_t* contig, int32_t* start,
                         int32_t* end, int32_t* qlen, uint8_t* highq, int64_t* n) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (!ctx->filtered) return ctx->fail(GCI_E_ARG, "gci_fetch_file_table before gci_filter");
  if (file < 0 || file >= (int)ctx->n_files) return ctx->fail(GCI_E_ARG, "bad file index %d", file);
  FileTable& f = ctx->files[file];
  const uint32_t nr = ctx->n_reads;
  DevBuf mark, pos;
  GCI_TRY(ctx->ensure(mark, 4 * (size_t)(nr + 1)));
  GCI_TRY(ctx->ensure(pos, 4 * (size_t)(nr + 1)));
  int64_t cnt = 0;
  if (nr) {
    mark_present_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(nr, f.win.as<long long>(), mark.as<int32_t>());
    GCI_LAUNCH_CHECK(ctx);
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(mark.as<int32_t>() + nr, 0, 4, ctx->stream));
    GCI_TRY(gci_exclusive_scan_i32(ctx, mark.as<int32_t>(), pos.as<int32_t>(), (int64_t)nr + 1));
    int32_t c32 = 0;
    GCI_TRY(gci_d2h(ctx, &c32, pos.as<int32_t>() + nr, 4));
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cnt = c32;
  }
  if (n) *n = cnt;
  int rc = GCI_OK;
  if (read_id || contig || start || end || qlen || highq) {
    if (cap < cnt) {
      rc = ctx->fail(GCI_E_ARG, "table buffer too small");
    } else if (cnt) {
      JoinFile jf;
      jf.win = f.win.as<long long>();
      if (f.kind == 0) {
        BamFile& b = ctx->bam[f.src];
        jf.ref_id = b.ref_id.as<int32_t>(); jf.start = b.ref_start.as<int32_t>();
        jf.end = b.ref_end.as<int32_t>(); jf.qlen = b.qlen.as<int32_t>();
      } else {
        jf.ref_id = f.ref_id.as<int32_t>(); jf.start = f.start.as<int32_t>();
        jf.end = f.end.as<int32_t>(); jf.qlen = f.qlen.as<int32_t>();
      }
      DevBuf o_r, o_c, o_s, o_e, o_q, o_h;
      for (DevBuf* d : {&o_r, &o_c, &o_s, &o_e, &o_q}) GCI_TRY(ctx->ensure(*d, 4 * (size_t)cnt));
      GCI_TRY(ctx->ensure(o_h, (size_t)cnt));
      gather_table_kernel<<<(nr + 255) / 256, 256, 0, ctx->stream>>>(
          nr, mark.as<int32_t>(), pos.as<int32_t>(), jf, ctx->highq.as<uint8_t>(), o_r.as<uint32_t>(),
          o_c.as<int32_t>(), o_s.as<int32_t>(), o_e.as<int32_t>(), o_q.as<int32_t>(), o_h.as<uint8_t>());
      GCI_LAUNCH_CHECK(ctx);
      if (read_id) GCI_TRY(gci_d2h(ctx, read_id, o_r.p, 4 * cnt));
      if (contig) GCI_TRY(gci_d2h(ctx, contig, o_c.p, 4 * cnt));
      if (start) GCI_TRY(gci_d2h(ctx, start, o_s.p, 4 * cnt));
      if (end) GCI_TRY(gci_d2h(ctx, end, o_e.p, 4 * cnt));
      if (qlen) GCI_TRY(gci_d2h(ctx, qlen, o_q.p, 4 * cnt));
      if (highq) GCI_TRY(gci_d2h(ctx, highq, o_h.p, cnt));
      GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      for (DevBuf* d : {&o_r, &o_c, &o_s, &o_e, &o_q, &o_h}) ctx->release(*d);
    }
  }
  ctx->release(mark);
  ctx->release(pos);
  return rc;
}


}  // extern "C"

// ---- gci_pipeline internals -----------------------------------------------------------------------
struct PipeArgs {
  int32_t track, map_qual, mq_cutoff, flank_len, lo, hi;
  double ip, cp, op, dp;
  bool with_rows;
  int64_t sum_len, cap, row_n, n_sel;
  int world;
};
struct PipeOut {           // where the step's results land in the persistent pinned block
  unsigned long long* h_err;
  int64_t *h_off, *h_res, *h_rows;
  int64_t no, n_slots;
};

// pinned block: [err 4 x u64 | owner_off n_sel+3 | score result | genome rows]
static void pipeline_layout(gci_ctx* ctx, const PipeArgs& a, const Track& t, int64_t no, int64_t n_slots, PipeOut* o) {
  const int64_t res_bound = 3 * a.n_sel + 1 + std::max<int64_t>(t.iv_cap, 4096) + a.n_sel;
  o->h_err = (unsigned long long*)ctx->pipe_pin;
  o->h_off = (int64_t*)(o->h_err + 4);
  o->h_res = o->h_off + (a.n_sel + 3);
  o->h_rows = o->h_res + res_bound;
  o->no = no;
  o->n_slots = n_slots;
}

// everything the enqueued work depends on besides device memory contents
static std::vector<char> pipeline_signature(gci_ctx* ctx, const PipeArgs& a) {
  std::vector<char> sig;
  auto put = [&sig](const void* p, size_t n) { sig.insert(sig.end(), (const char*)p, (const char*)p + n); };
  auto put64 = [&put](int64_t v) { put(&v, sizeof v); };
  put64(a.track); put64(a.map_qual); put64(a.mq_cutoff); put64(a.flank_len); put64(a.lo); put64(a.hi);
  put(&a.ip, 8); put(&a.cp, 8); put(&a.op, 8); put(&a.dp, 8);
  put64(a.with_rows); put64(a.sum_len); put64(a.cap); put64(a.world); put64(a.n_sel);
  put64((int64_t)ctx->epoch); put64((int64_t)ctx->alloc_gen); put64((int64_t)(intptr_t)ctx->stream);
  put64((int64_t)(intptr_t)ctx->nccl_comm); put64(ctx->p2p_ok); put64(ctx->n_reads); put64((int64_t)ctx->n_files);
  put64((int64_t)ctx->n_bam); put64((int64_t)ctx->n_paf);
  put64(ctx->shard.on); put64(ctx->shard.rank); put64(ctx->shard.world); put64(ctx->shard.cap1); put64(ctx->shard.opened);
  for (size_t i = 0; i < ctx->n_files; i++) {
    const FileTable& f = ctx->files[i];
    put64(f.paf >= 0 ? 2 : f.kind); put64(f.src); put64(f.paf); put64(f.paf >= 0 ? 0 : f.n);
  }
  for (size_t i = 0; i < ctx->n_bam; i++) {
    const BamFile& b = ctx->bam[i];
    put64(b.n); put64(b.n_ops); put64(b.n_dense); put64(b.n_span);   // grids and list lengths of the CIGAR / gate kernels
  }
  for (size_t i = 0; i < ctx->n_paf; i++) put64(ctx->paf[i].n);
  return sig;
}

static int pipeline_enqueue(gci_ctx* ctx, const PipeArgs& a, PipeOut* out) {
  GCI_TRY(reset_err(ctx));
  // sharded read set: the BAM winners leave for their homes first, the PAF legs (home-local) run while they travel.
  // (Running the PAF legs on a second stream beside the BAM legs changed nothing, 3.541 against 3.548 ms per step,
  // profiles/r02af: every kernel of both legs brings far more CTAs than the GPU holds, so the block scheduler drains
  // one grid after the other and the two legs' warps never share an SM.)
  if (!ctx->shard.on) GCI_TRY(gci_run_paf_legs(ctx, a.map_qual, a.mq_cutoff, a.ip));
  for (size_t i = 0; i < ctx->n_files; i++)
    if (ctx->files[i].kind == 0)
      GCI_TRY(gci_run_bam_leg(ctx, (int)i, ctx->files[i].src, a.map_qual, a.mq_cutoff, a.ip, a.cp));
  if (ctx->shard.on) {
    GCI_TRY(gci_shard_dispatch_enqueue(ctx));
    GCI_TRY(gci_run_paf_legs(ctx, a.map_qual, a.mq_cutoff, a.ip));
    GCI_TRY(gci_shard_exchange_enqueue(ctx, a.op, a.track, a.flank_len));   // join at the read homes
  } else {
    GCI_TRY(gci_run_join_counting(ctx, a.op, a.track, a.flank_len));
  }
  ctx->filtered = true;
  GCI_TRY(gci_depth_enqueue(ctx, a.track, a.flank_len, a.lo, a.hi));
  Track& t = ctx->track[a.track];
  const int64_t res_bound = 3 * a.n_sel + 1 + std::max<int64_t>(t.iv_cap, 4096) + a.n_sel;
  const size_t need = 8 * (size_t)(4 + (a.n_sel + 3) + res_bound + a.row_n * a.world);
  if (ctx->pipe_pin_cap < need) {
    ctx->alloc_gen++;
    if (ctx->capturing) ctx->capture_abort = true;
    if (ctx->pipe_pin) cudaFreeHost(ctx->pipe_pin);
    ctx->pipe_pin = nullptr;
    ctx->pipe_pin_cap = 0;
    if (cudaHostAlloc(&ctx->pipe_pin, need, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return ctx->fail(GCI_E_NOMEM, "pinned allocation failed");
    }
    ctx->pipe_pin_cap = need;
  }
  pipeline_layout(ctx, a, t, 0, 0, out);
  GCI_TRY(gci_scan_enqueue(ctx, a.track, a.lo, a.hi, a.flank_len, out->h_off));
  ctx->stage_begin(GCI_ST_SCORE);
  GCI_TRY(gci_launch_score_kernels(ctx, t, a.dp, a.flank_len, &out->no, &out->n_slots, true));
  if (a.with_rows) GCI_TRY(gci_enqueue_genome_row(ctx, t, out->no, a.sum_len, a.cap, out->h_rows));
  GCI_TRY(gci_d2h(ctx, out->h_res, ctx->tmp[1].p, 8 * (size_t)(3 * out->no + 1 + out->n_slots)));
  GCI_TRY(gci_d2h(ctx, out->h_err, ctx->d_err.p, 4 * sizeof(unsigned long long)));
  ctx->stage_end();
  return GCI_OK;
}

extern "C" {

// filter -> depth -> scan -> score terms with ONE host synchronisation: nothing between the stages needs the
// host (the interval buffers keep their capacity from earlier scans; if a run produces more intervals than
// fit, the scan + score part is redone through the synchronous entry points).
int gci_pipeline(gci_ctx* ctx, int32_t track, int32_t map_qual, int32_t mq_cutoff, double iden_percent,
                 double clip_percent, double ovlp_percent, int32_t flank_len, int32_t lo, int32_t hi,
                 double dist_percent, int64_t* n_survivors, int64_t* n_intervals, int64_t* n50, int64_t* n_ctg,
                 int64_t* depth_sums) {
  return gci_pipeline_row(ctx, track, map_qual, mq_cutoff, iden_percent, clip_percent, ovlp_percent, flank_len, lo, hi,
                          dist_percent, n_survivors, n_intervals, n50, n_ctg, depth_sums, 0, 0, nullptr);
}

// the same with the multi-GPU genome row in the same synchronisation: rows != NULL adds one ncclAllGather
// (gci_comm_init must have been called) between the score kernels and the device->host copy
int gci_pipeline_row(gci_ctx* ctx, int32_t track, int32_t map_qual, int32_t mq_cutoff, double iden_percent,
                     double clip_percent, double ovlp_percent, int32_t flank_len, int32_t lo, int32_t hi,
                     double dist_percent, int64_t* n_survivors, int64_t* n_intervals, int64_t* n50, int64_t* n_ctg,
                     int64_t* depth_sums, int64_t sum_len, int64_t cap, int64_t* rows) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  if (rows && (!ctx->nccl_comm || cap < 1)) return ctx->fail(GCI_E_ARG, "gci_pipeline_row: gci_comm_init / cap missing");
  cudaSetDevice(ctx->device);
  if (ctx->n_files == 0) return ctx->fail(GCI_E_ARG, "gci_pipeline: no files uploaded");
  PipeArgs a;
  a.track = track; a.map_qual = map_qual; a.mq_cutoff = mq_cutoff; a.flank_len = flank_len; a.lo = lo; a.hi = hi;
  a.ip = iden_percent; a.cp = clip_percent; a.op = ovlp_percent; a.dp = dist_percent;
  a.with_rows = rows != nullptr; a.sum_len = sum_len; a.cap = cap;
  a.row_n = rows ? 4 + cap : 0;
  a.world = rows ? ctx->comm_world : 0;
  a.n_sel = 0;
  for (int c = 0; c < ctx->n_contigs; c++) a.n_sel += ctx->selected[c] ? 1 : 0;
  if (a.n_sel == 0) return ctx->fail(GCI_E_ARG, "gci_pipeline: no contigs");
  Track& t = ctx->track[track];

  // ---- enqueue: eagerly, or as one graph launch when this exact step has been seen before ----
  std::vector<char> sig;
  const bool graphs = !ctx->timing && ctx->graph_ok;
  if (graphs) sig = pipeline_signature(ctx, a);
  PipeOut out;
  if (graphs && ctx->pipe_exec && sig == ctx->pipe_sig) {
    GCI_CUDA_TRY(ctx, cudaGraphLaunch(ctx->pipe_exec, ctx->stream));
    ctx->graph_replays++;
    ctx->launches += ctx->pipe_launches;
    ctx->filtered = true;
    pipeline_layout(ctx, a, t, ctx->pipe_no, ctx->pipe_slots, &out);
  } else {
    ctx->drop_graph();
    bool done = false;
    if (graphs && sig == ctx->pipe_sig_seen && sig != ctx->pipe_sig_bad) {
      // second identical step: record it.  Everything it allocates exists since the first one; if a buffer
      // moves or a host->device copy shows up all the same, the capture is thrown away.
      const int64_t launches0 = ctx->launches;
      const uint64_t gen0 = ctx->alloc_gen;
      ctx->capture_abort = false;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t exec = nullptr;
      bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
      if (ok) {
        ctx->capturing = true;
        const int rc = pipeline_enqueue(ctx, a, &out);
        ctx->capturing = false;
        ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr;
        ok = ok && rc == GCI_OK && !ctx->capture_abort && gen0 == ctx->alloc_gen;
        ok = ok && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
      }
      if (ok && cudaGraphLaunch(exec, ctx->stream) == cudaSuccess) {
        ctx->pipe_exec = exec;
        ctx->pipe_sig = sig;
        ctx->pipe_launches = ctx->launches - launches0;
        ctx->pipe_no = out.no;
        ctx->pipe_slots = out.n_slots;
        ctx->graph_replays++;
        done = true;
      } else {
        if (exec) cudaGraphExecDestroy(exec);
        cudaGetLastError();
        ctx->launches = launches0;
        ctx->pipe_sig_bad = sig;          // do not try again for this step shape
        // nothing recorded has run: a table upload that was captured (it is what aborts a capture) never
        // reached the device, so the host-side "already uploaded" marks must not survive
        ctx->lay_cache.clear();
        ctx->ob_cache.clear();
        ctx->owner_of_stale = true;
      }
    }
    if (!done) GCI_TRY(pipeline_enqueue(ctx, a, &out));
    if (graphs) ctx->pipe_sig_seen = graphs && !done ? pipeline_signature(ctx, a) : sig;
  }
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));

  // ---- finish on the host ----
  const int64_t no = out.no;
  if (out.h_err[0] != 0) {
    ctx->filtered = false;
    unsigned long long dummy;
    return check_err(ctx, "gci_pipeline", &dummy);
  }
  ctx->n_survivors = (int64_t)out.h_err[2];
  if (n_survivors) *n_survivors = ctx->n_survivors;
  bool overflow = false;
  GCI_TRY(gci_scan_finish(ctx, track, out.h_off, &overflow));
  if (overflow) {
    // rare: redo the scan with grown buffers and the score terms through the synchronous entry points
    ctx->drop_graph();
    GCI_TRY(gci_scan(ctx, track, lo, hi, flank_len, n_intervals));
    if (rows) return gci_genome_row(ctx, track, dist_percent, flank_len, sum_len, cap, n50, n_ctg, depth_sums, rows);
    return gci_score_terms_sums(ctx, track, dist_percent, flank_len, n50, n_ctg, 0, nullptr, nullptr, depth_sums);
  }
  if (rows) {
    memcpy(rows, out.h_rows, 8 * (size_t)a.row_n * a.world);
    for (int r = 0; r < a.world; r++)
      if (rows[(size_t)r * a.row_n + 3] < 0) {
        ctx->p2p_ok = false;
        ctx->drop_graph();
        return ctx->fail(GCI_E_CUDA, "gci_pipeline_row: rank %d did not deliver its row over peer memory in time", r);
      }
    for (int r = 0; r < a.world; r++)
      if (rows[(size_t)r * a.row_n + 3] > cap)
        return ctx->fail(GCI_E_ARG, "gci_pipeline_row: rank %d has more curated lengths than cap", r);
  }
  if (n_intervals) *n_intervals = t.n_intervals;
  long long all_c = 0, all_d = 0;
  for (int64_t o = 0; o < no; o++) {
    if (n_ctg) n_ctg[o] = out.h_res[no + 1 + o];
    if (depth_sums) depth_sums[o] = out.h_res[2 * no + 1 + o];
    all_c += out.h_res[no + 1 + o];
    all_d += out.h_res[2 * no + 1 + o];
  }
  if (n50) memcpy(n50, out.h_res, 8 * (size_t)(no + 1));
  if (n_ctg) n_ctg[no] = all_c;
  if (depth_sums) depth_sums[no] = all_d;
  return GCI_OK;
}

}  // extern "C"
