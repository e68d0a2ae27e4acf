// Device-wide scans, issue-run extraction from the flag bitmask (GCI.py:356-390) and score terms
// (GCI.py:422-519).
#include <algorithm>

#include "common.cuh"

// ================================================================================================
// generic exclusive scan: per-block reduce -> (recursive) scan of block sums -> per-block scan
// ================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;   // 2048 items per block

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const TI* __restrict__ in, int64_t n, TO* __restrict__ block_sums) {
  __shared__ TO s_w[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
  TO v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) v += (TO)in[i];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    TO t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}

// exclusive scan of one block's items (thread-blocked order), adding block_off[blockIdx]
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n, const TO* __restrict__ block_off,
                  TO* __restrict__ total /* nullable: written by the last block */) {
  __shared__ TO s_w[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  TO item[SCAN_ITEMS];
  TO sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int64_t i = base + k;
    item[k] = i < n ? (TO)in[i] : (TO)0;
    sum += item[k];
  }
  // warp inclusive scan of thread sums
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  TO incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    TO t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  TO woff = 0;
  for (int j = 0; j < w; j++) woff += s_w[j];
  TO run = (block_off ? block_off[blockIdx.x] : (TO)0) + woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int64_t i = base + k;
    if (i < n) out[i] = run;
    run += item[k];
  }
  if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total = run;
}

// one block walks the whole array with a running carry: one launch for arrays up to ~10^5 items
// (tile tables, run chunks) instead of reduce + scan + apply
constexpr int SCAN1_THREADS = 1024;
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN1_THREADS)
scan_single_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n, TO* __restrict__ total) {
  __shared__ TO s_w[SCAN1_THREADS / 32];
  __shared__ TO s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += SCAN1_THREADS) {
    const int64_t i = base + threadIdx.x;
    const TO v = i < n ? (TO)in[i] : (TO)0;
    TO incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      TO t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    TO woff = 0, tot = 0;
    for (int j = 0; j < SCAN1_THREADS / 32; j++) {
      const TO x = s_w[j];
      if (j < w) woff += x;
      tot += x;
    }
    const TO carry = s_carry;
    if (i < n) out[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (total && threadIdx.x == 0) *total = s_carry;
}

template <typename TI, typename TO>
static int exclusive_scan(gci_ctx* ctx, const TI* in, TO* out, int64_t n, TO* total_dev, int depth = 0) {
  if (n <= 0) {
    if (total_dev) GCI_CUDA_TRY(ctx, cudaMemsetAsync(total_dev, 0, sizeof(TO), ctx->stream));
    return GCI_OK;
  }
  const int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (n <= 8 * 1024) {
    scan_single_kernel<TI, TO><<<1, SCAN1_THREADS, 0, ctx->stream>>>(in, out, n, total_dev);
    GCI_LAUNCH_CHECK(ctx);
    return GCI_OK;
  }
  // block sums / offsets live in per-level context scratch (no allocation in steady state)
  if (depth >= 4) return ctx->fail(GCI_E_ARG, "scan recursion too deep");
  DevBuf& sums = ctx->scan_lvl[2 * depth];
  DevBuf& offs = ctx->scan_lvl[2 * depth + 1];
  GCI_TRY(ctx->ensure(sums, sizeof(TO) * (size_t)nb));
  GCI_TRY(ctx->ensure(offs, sizeof(TO) * (size_t)nb));
  scan_reduce_kernel<TI, TO><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums.as<TO>());
  GCI_LAUNCH_CHECK(ctx);
  GCI_TRY((exclusive_scan<TO, TO>(ctx, sums.as<TO>(), offs.as<TO>(), nb, nullptr, depth + 1)));
  scan_apply_kernel<TI, TO><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, offs.as<TO>(), total_dev);
  GCI_LAUNCH_CHECK(ctx);
  return GCI_OK;
}

int gci_exclusive_scan_i64_from_i32(gci_ctx* ctx, const int32_t* in, int64_t* out, int64_t n, int64_t* total_dev) {
  return exclusive_scan<int32_t, long long>(ctx, in, reinterpret_cast<long long*>(out), n,
                                            reinterpret_cast<long long*>(total_dev));
}

int gci_exclusive_scan_i32(gci_ctx* ctx, const int32_t* in, int32_t* out, int64_t n) {
  return exclusive_scan<int32_t, int32_t>(ctx, in, out, n, nullptr);
}

// tile table from the per-tile event counts: .x = (events, net) packed = (starts + ends) | (starts - ends) << 32,
// .y = exclusive scan of .x over all tiles (the two halves never carry into each other: a contig's nets sum to 0 and
// the event total stays below 2^32)
__device__ __forceinline__ unsigned long long tile_pack(const uint32_t* __restrict__ cs, const uint32_t* __restrict__ ce,
                                                       int64_t i) {
  const uint32_t s = cs[i], e = ce[i];
  return (unsigned long long)(s + e) | ((unsigned long long)(uint32_t)(s - e) << 32);
}

__global__ void __launch_bounds__(SCAN_THREADS)
tile_reduce_kernel(const uint32_t* __restrict__ cs, const uint32_t* __restrict__ ce, int64_t n,
                   unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long s_w[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK;
  unsigned long long v = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    const int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) v += tile_pack(cs, ce, i);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; w++) t += s_w[w];
    block_sums[blockIdx.x] = t;
  }
}

// block b first sums the block totals before it (at most a few thousand values: 10 Gbp = 4 768 blocks), so
// no separate scan of the block totals is needed between the reduce and the apply pass
__global__ void __launch_bounds__(SCAN_THREADS)
tile_apply_kernel(const uint32_t* __restrict__ cs, const uint32_t* __restrict__ ce, ulonglong2* __restrict__ ps,
                  int64_t n, const unsigned long long* __restrict__ block_sums) {
  __shared__ unsigned long long s_w[SCAN_THREADS / 32];
  __shared__ unsigned long long s_b[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned long long before = 0;
  if (block_sums) {
    for (int64_t i = threadIdx.x; i < (int64_t)blockIdx.x; i += SCAN_THREADS) before += block_sums[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) before += __shfl_xor_sync(0xffffffffu, before, d);
    if (lane == 0) s_b[w] = before;
  }
  const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + (int64_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long item[SCAN_ITEMS], sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    item[k] = base + k < n ? tile_pack(cs, ce, base + k) : 0ull;
    sum += item[k];
  }
  unsigned long long incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  unsigned long long woff = 0;
  for (int j = 0; j < w; j++) woff += s_w[j];
  if (block_sums)
    for (int j = 0; j < SCAN_THREADS / 32; j++) woff += s_b[j];
  unsigned long long run = woff + incl - sum;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) ps[base + k] = make_ulonglong2(item[k], run);
    run += item[k];
  }
}

int gci_scan_tile_pack(gci_ctx* ctx, const uint32_t* cnt_start, const uint32_t* cnt_end, ulonglong2* tile_ps, int64_t n) {
  if (n <= 0) return GCI_OK;
  const int64_t nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (nb == 1) {
    tile_apply_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(cnt_start, cnt_end, tile_ps, n, nullptr);
    GCI_LAUNCH_CHECK(ctx);
    return GCI_OK;
  }
  DevBuf& sums = ctx->scan_lvl[6];
  GCI_TRY(ctx->ensure(sums, 8 * (size_t)nb));
  tile_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(cnt_start, cnt_end, n, sums.as<unsigned long long>());
  GCI_LAUNCH_CHECK(ctx);
  tile_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(cnt_start, cnt_end, tile_ps, n,
                                                                   sums.as<unsigned long long>());
  GCI_LAUNCH_CHECK(ctx);
  return GCI_OK;
}

// ================================================================================================
// K7  run extraction from the flag bitmask
// ================================================================================================
// A scan "owner" is a window [lo, hi) of one contig (whole-genome scan: one window per selected
// contig, lo from the drop rule below, hi = L - fl; regions scan: the regions themselves).  Bits
// outside the window read as 0, so every run start has a matching end at or before position hi.

__device__ __forceinline__ bool flag_bit(const uint32_t* __restrict__ flags, int64_t gpos) {
  return (flags[gpos >> 5] >> (gpos & 31)) & 1u;
}

// Whole-genome windows with the reference's drop rule (GCI.py:385 `if i > flank_len`): a run whose
// end is <= 2*fl is dropped unless it touches the view end L - fl (GCI.py:380-382).  All dropped runs
// sit in [fl, 2*fl), so the rule is equivalent to moving the window start to
//   h = start of the run covering 2*fl-1   if that run continues past 2*fl,   else 2*fl
// (for contigs with L - fl <= 2*fl: the start of the run touching L - fl, else the empty window).
__global__ void genome_windows_kernel(int32_t n_contigs, const int64_t* __restrict__ len,
                                      const uint8_t* __restrict__ selected, const int64_t* __restrict__ tile_off,
                                      const uint32_t* __restrict__ flags, int32_t fl, int32_t* __restrict__ w_contig,
                                      int64_t* __restrict__ w_lo, int64_t* __restrict__ w_hi,
                                      const int32_t* __restrict__ owner_of_contig) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_contigs || !selected[c]) return;
  const int o = owner_of_contig[c];
  const int64_t L = len[c];
  const int64_t g0 = tile_off[c] * GCI_TILE;
  int64_t lo = 0, hi = 0;
  if (fl >= 0 && L - 2 * (int64_t)fl > 0) {
    hi = L - fl;
    if (fl == 0) {
      lo = 0;
    } else {
      const int64_t two = 2 * (int64_t)fl;
      int64_t probe;      // last position whose run may survive
      bool keep;
      if (hi <= two) {
        probe = hi - 1;
        keep = flag_bit(flags, g0 + probe);
        lo = hi;
      } else {
        probe = two - 1;
        keep = flag_bit(flags, g0 + probe) && flag_bit(flags, g0 + two);
        lo = two;
      }
      if (keep) {
        int64_t p = probe;
        while (p > fl && flag_bit(flags, g0 + p - 1)) p--;
        lo = p;
      }
    }
  }
  w_contig[o] = c;
  w_lo[o] = lo;
  w_hi[o] = hi;
}

struct RunWin {
  int64_t g0;       // global padded position of the contig start
  int64_t lo, hi;   // window in contig coordinates
  int64_t w0;       // first word (contig-relative) of the chunk layout, <= lo >> 5
  int64_t n_words;  // words of the layout: covers [w0 * 32, hi] inclusive
};

__device__ __forceinline__ RunWin load_win(int64_t o, const int64_t* w_lo, const int64_t* w_hi, const int64_t* lay_w0,
                                           const int64_t* lay_g0) {
  RunWin w;
  w.g0 = lay_g0[o];
  w.lo = w_lo[o];
  w.hi = w_hi[o];
  w.w0 = lay_w0[o];
  w.n_words = w.hi > w.lo ? (w.hi >> 5) - w.w0 + 1 : 0;
  return w;
}

// flag word `j` (layout-relative) restricted to the window
__device__ __forceinline__ uint32_t win_word(const uint32_t* __restrict__ flags, const RunWin& w, int64_t j) {
  if (j < 0 || j >= w.n_words) return 0u;
  const int64_t cw = w.w0 + j;                  // contig-relative word
  const int64_t p0 = cw << 5;
  if (p0 + 32 <= w.lo) return 0u;
  uint32_t m = flags[(w.g0 >> 5) + cw];
  if (p0 < w.lo) m &= 0xffffffffu << (w.lo - p0);
  if (p0 + 32 > w.hi) {
    const int64_t keep = w.hi - p0;             // bits [0, keep) stay
    m = keep <= 0 ? 0u : (keep >= 32 ? m : (m & ((1u << keep) - 1u)));
  }
  return m;
}

constexpr int RUN_THREADS = 256;   // (CTAs of 1024 threads over 262 144 positions measured 0.38 instead of 0.21 ms, r02o)
constexpr int RUN_WPT = GCI_RUN_CHUNK_WORDS / RUN_THREADS;   // words per thread (8)
// Run starts / ends a chunk may leave in its staging slots during the COUNT pass.  The positions are known there
// already, only their place in the interval arrays is not: staging them spares the second walk over the flag words
// of every chunk that has a run border (17 M warp instructions, 62 us for configs[2] under ncu, profiles/r02ac).
constexpr int RUN_STAGE = 16;

// chunk -> owner through a host-built table (the chunk layout only depends on the contig table / regions):
// one load instead of a binary search, and the contig's first flag word comes with the layout, so a CTA sees
// two dependent global loads before its flag words instead of eight.
// One chunk by one CTA (block-collective; every return is taken by the whole CTA).  !WRITE: count, and stage the
// positions of a chunk with few of them.  WRITE: a chunk with more than RUN_STAGE starts or ends, straight to the
// interval arrays at the offsets the scan produced.
struct RunArgs {
  const uint32_t* flags;
  const int32_t* chunk_owner;
  const int64_t *chunk_off, *w_lo, *w_hi, *lay_w0, *lay_g0;
  int2* cnt;
  const longlong2* off;
  int32_t *iv_start, *iv_end;
  int64_t cap;
  int32_t* stage;                 // [n_chunks][2][RUN_STAGE]
  int64_t n_chunks;
};

template <bool WRITE>
__device__ __forceinline__ void runs_chunk(const RunArgs& a, const int64_t chunk) {
  const uint32_t* __restrict__ flags = a.flags;
  const int32_t* __restrict__ chunk_owner = a.chunk_owner;
  const int64_t* __restrict__ chunk_off = a.chunk_off;
  const int64_t *__restrict__ w_lo = a.w_lo, *__restrict__ w_hi = a.w_hi, *__restrict__ lay_w0 = a.lay_w0,
                *__restrict__ lay_g0 = a.lay_g0;
  int2* __restrict__ cnt = a.cnt;
  const longlong2* __restrict__ off = a.off;
  int32_t *__restrict__ iv_start = a.iv_start, *__restrict__ iv_end = a.iv_end;
  int64_t cap = a.cap;
  int32_t* __restrict__ stage = a.stage;
  __shared__ int s_ws[RUN_THREADS / 32], s_we[RUN_THREADS / 32];
  const int64_t o = chunk_owner[chunk];
  const RunWin w = load_win(o, w_lo, w_hi, lay_w0, lay_g0);
  const int64_t j0 = (chunk - chunk_off[o]) * GCI_RUN_CHUNK_WORDS + (int64_t)threadIdx.x * RUN_WPT;
  uint32_t m[RUN_WPT];
  int ns = 0, ne = 0;
  uint32_t prev;
  // my 8 words and the one before them lie completely inside the window: two 16-byte loads, no masking
  const int64_t p0w = (w.w0 + j0) << 5;
  const int64_t gw = (w.g0 >> 5) + w.w0 + j0;
  bool quiet = false;
  if (j0 >= 1 && j0 + RUN_WPT <= w.n_words && p0w - 32 >= w.lo && p0w + RUN_WPT * 32 <= w.hi && (gw & 3) == 0) {
    const uint4 a = *reinterpret_cast<const uint4*>(flags + gw), b = *reinterpret_cast<const uint4*>(flags + gw + 4);
    prev = flags[gw - 1];
    m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
    quiet = ((a.x | a.y | a.z | a.w | b.x | b.y | b.z | b.w) == 0u) && (prev >> 31) == 0u;
  } else {
    prev = win_word(flags, w, j0 - 1);
    uint32_t any = prev >> 31;
#pragma unroll
    for (int k = 0; k < RUN_WPT; k++) {
      m[k] = win_word(flags, w, j0 + k);
      any |= m[k];
    }
    quiet = any == 0u;
  }
  // almost every stretch of an assembly is issue-free: no bit set in the CTA's chunk and none just before its
  // threads' words -> nothing starts or ends here
  if (__syncthreads_and(quiet)) {
    if (!WRITE && threadIdx.x == 0) cnt[chunk] = make_int2(0, 0);
    return;
  }
  // run starts / ends of word k (recomputed where they are written: the masks of eight words cost 16 registers, and
  // the count pass lives on eight resident CTAs per SM)
  auto borders = [&](int k, uint32_t& st, uint32_t& en) {
    const uint32_t before = k ? m[k - 1] : prev;
    const uint32_t sh = (m[k] << 1) | (before >> 31);
    st = m[k] & ~sh;
    en = ~m[k] & sh;
    // an end can only be reported on a word that exists in the layout (positions <= hi)
    if (j0 + k >= w.n_words) en = 0u;
  };
#pragma unroll
  for (int k = 0; k < RUN_WPT; k++) {
    uint32_t st, en;
    borders(k, st, en);
    ns += __popc(st);
    ne += __popc(en);
  }
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int is = warp_incl_scan(ns, lane), ie = warp_incl_scan(ne, lane);
  if (lane == 31) { s_ws[wp] = is; s_we[wp] = ie; }
  __syncthreads();
  int bs = 0, be = 0, ts = 0, te = 0;
  for (int j = 0; j < RUN_THREADS / 32; j++) {
    if (j < wp) { bs += s_ws[j]; be += s_we[j]; }
    ts += s_ws[j]; te += s_we[j];
  }
  if (!WRITE) {
    if (threadIdx.x == 0) cnt[chunk] = make_int2(ts, te);
    if (ts > RUN_STAGE || te > RUN_STAGE) return;           // the write pass walks this chunk again
  }
  // count pass: the chunk's few positions go to its staging slots (runs_unstage_kernel moves them once the offsets
  // are known); write pass (a chunk with more than RUN_STAGE of either kind): straight to the interval arrays
  int32_t* out_s = WRITE ? iv_start : stage + chunk * (2 * RUN_STAGE);
  int32_t* out_e = WRITE ? iv_end : stage + chunk * (2 * RUN_STAGE) + RUN_STAGE;
  if (!WRITE) cap = RUN_STAGE;
  longlong2 base = make_longlong2(0, 0);
  if (WRITE) base = off[chunk];
  int64_t ps = base.x + bs + is - ns;
  int64_t pe = base.y + be + ie - ne;
#pragma unroll
  for (int k = 0; k < RUN_WPT; k++) {
    const int64_t p0 = (w.w0 + j0 + k) << 5;
    uint32_t st, en;
    borders(k, st, en);
    while (st) {
      const int b = __ffs(st) - 1;
      st &= st - 1;
      if (ps < cap) out_s[ps] = (int32_t)(p0 + b);
      ps++;
    }
    while (en) {
      const int b = __ffs(en) - 1;
      en &= en - 1;
      if (pe < cap) out_e[pe] = (int32_t)(p0 + b);
      pe++;
    }
  }
}

// count pass: one CTA per chunk
__global__ void __launch_bounds__(RUN_THREADS, 8)   // 32 registers: eight CTAs per SM keep enough flag loads in flight
runs_count_kernel(RunArgs a) { runs_chunk<false>(a, blockIdx.x); }

// write pass for the chunks the count pass could not stage (normally none): a bounded grid looks at 32 counts per
// CTA and round, and walks the few chunks that need it (a CTA per chunk was 28 us of launches for nothing)
__global__ void __launch_bounds__(RUN_THREADS)
runs_write_kernel(RunArgs a) {
  __shared__ int s_list[32];
  __shared__ int s_n;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < a.n_chunks; base += (int64_t)gridDim.x * 32) {
    if (threadIdx.x < 32) {
      const int64_t chunk = base + threadIdx.x;
      bool need = false;
      if (chunk < a.n_chunks) {
        const int2 c = a.cnt[chunk];
        need = c.x > RUN_STAGE || c.y > RUN_STAGE;
      }
      const unsigned m = __ballot_sync(0xffffffffu, need);
      if (need) s_list[__popc(m & ((1u << threadIdx.x) - 1u))] = (int)threadIdx.x;
      if (threadIdx.x == 0) s_n = __popc(m);
    }
    __syncthreads();
    const int n = s_n;
    for (int i = 0; i < n; i++) {
      runs_chunk<true>(a, base + s_list[i]);
      __syncthreads();                                      // the chunk's shared scratch is free again
    }
    __syncthreads();                                        // s_list / s_n are rewritten in the next round
  }
}

// staged positions -> interval arrays, one thread per chunk (almost every chunk has nothing)
__global__ void __launch_bounds__(256)
runs_unstage_kernel(const int2* __restrict__ cnt, const longlong2* __restrict__ off, int64_t n_chunks,
                    const int32_t* __restrict__ stage, int32_t* __restrict__ iv_start, int32_t* __restrict__ iv_end,
                    int64_t cap) {
  const int64_t chunk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (chunk >= n_chunks) return;
  const int2 c = cnt[chunk];
  if ((c.x | c.y) == 0 || c.x > RUN_STAGE || c.y > RUN_STAGE) return;
  const longlong2 base = off[chunk];
  const int32_t* sp = stage + chunk * (2 * RUN_STAGE);
  for (int j = 0; j < c.x; j++)
    if (base.x + j < cap) iv_start[base.x + j] = sp[j];
  for (int j = 0; j < c.y; j++)
    if (base.y + j < cap) iv_end[base.y + j] = sp[RUN_STAGE + j];
}

// one block: exclusive scan of the (starts, ends) chunk counts, totals, and the per-owner interval offsets.
// Every WARP owns a contiguous strip of chunks and walks it 32 chunks at a time with coalesced loads: strip sums ->
// the 32 strip totals through shared memory -> second walk with a warp scan per 32 chunks, the next 32 counts
// already requested.  (A strip per THREAD made every load of a warp touch 32 different lines, one after the other:
// 55 us for 47 578 chunks under ncu, profiles/r02ac, with the rest of the GPU idle.)
__global__ void __launch_bounds__(1024)
runs_scan_kernel(const int2* __restrict__ cnt, int64_t n_chunks, longlong2* __restrict__ off, int64_t n_owners,
                 const int64_t* __restrict__ chunk_off, int64_t* __restrict__ owner_off /* [n_owners+1 | total_s | total_e] */) {
  __shared__ long long s_ws[32], s_we[32];
  __shared__ long long s_ts, s_te;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t strip = (((n_chunks + 31) / 32) + 31) & ~int64_t(31);
  const int64_t i0 = min(n_chunks, (int64_t)w * strip), i1 = min(n_chunks, i0 + strip);
  long long ss = 0, se = 0;
#pragma unroll 4
  for (int64_t i = i0 + lane; i < i1; i += 32) {
    const int2 v = cnt[i];
    ss += v.x;
    se += v.y;
  }
  ss = warp_sum_ll(ss);
  se = warp_sum_ll(se);
  if (lane == 0) { s_ws[w] = ss; s_we[w] = se; }
  __syncthreads();
  long long os = 0, oe = 0, ts = 0, te = 0;
  for (int j = 0; j < 32; j++) {
    const long long a = s_ws[j], b = s_we[j];
    if (j < w) { os += a; oe += b; }
    ts += a; te += b;
  }
  // a chunk holds at most 2^15 runs: the scan of 32 chunks fits 32 bits
  int2 v = (i0 + lane < i1) ? cnt[i0 + lane] : make_int2(0, 0);
  for (int64_t base = i0; base < i1; base += 32) {
    const int2 nxt = (base + 32 + lane < i1) ? cnt[base + 32 + lane] : make_int2(0, 0);
    const int xs = warp_incl_scan(v.x, lane), xe = warp_incl_scan(v.y, lane);
    if (base + lane < i1) off[base + lane] = make_longlong2(os + xs - v.x, oe + xe - v.y);
    os += __shfl_sync(0xffffffffu, xs, 31);
    oe += __shfl_sync(0xffffffffu, xe, 31);
    v = nxt;
  }
  if (threadIdx.x == 0) { s_ts = ts; s_te = te; }
  __syncthreads();                                          // off[] of the whole block is written (same block reads it)
  for (int64_t o = threadIdx.x; o <= n_owners; o += 1024) {
    const int64_t ch = chunk_off[o];
    owner_off[o] = (o == n_owners || ch >= n_chunks) ? s_ts : off[ch].x;
  }
  if (threadIdx.x == 0) { owner_off[n_owners + 1] = s_ts; owner_off[n_owners + 2] = s_te; }
}

// shared tail of gci_scan / gci_scan_windows.  The chunk layout comes from host-known bounds
// (lay_lo[o] <= real window start, hi[o] = real window end); the real windows are on the device.
// defer_pin != NULL: enqueue only — the interval offsets / totals are copied to defer_pin[n_owners + 3] and the
// caller synchronises later and calls finish_runs()
static int extract_runs(gci_ctx* ctx, Track& t, const std::vector<int64_t>& lay_lo, const std::vector<int64_t>& h_hi,
                        int64_t* n_intervals, int64_t* defer_pin = nullptr) {
  const int64_t n_owners = t.n_owners;
  // layout, int64 words: [chunk_off (n+1) | w0 (n) | g0 (n) | chunk_owner (int32 x n_chunks)]
  std::vector<int64_t> lay(3 * (n_owners + 1), 0);
  {
    int64_t* chunk_off = lay.data();
    int64_t* w0 = lay.data() + n_owners + 1;
    int64_t* g0 = w0 + n_owners;
    for (int64_t o = 0; o < n_owners; o++) {
      w0[o] = lay_lo[o] >> 5;
      g0[o] = ctx->pos_off[t.owner_contig[o]];
      const int64_t nw = h_hi[o] > lay_lo[o] ? (h_hi[o] >> 5) - w0[o] + 1 : 0;
      chunk_off[o + 1] = chunk_off[o] + (nw + GCI_RUN_CHUNK_WORDS - 1) / GCI_RUN_CHUNK_WORDS;
    }
  }
  const int64_t n_chunks = lay[n_owners];
  const size_t lay_head = lay.size();
  if (n_chunks > 0 && n_chunks < (int64_t(1) << 31)) {
    lay.resize(lay_head + (size_t)(n_chunks + 1) / 2, 0);
    int32_t* owner = reinterpret_cast<int32_t*>(lay.data() + lay_head);
    for (int64_t o = 0; o < n_owners; o++)
      for (int64_t c = lay[o]; c < lay[o + 1]; c++) owner[c] = (int32_t)o;
  }
  t.h_owner_off.assign(n_owners + 1, 0);
  t.n_intervals = 0;
  if (n_intervals) *n_intervals = 0;
  GCI_TRY(ctx->ensure(t.owner_off, sizeof(int64_t) * (size_t)(n_owners + 3)));
  if (n_chunks == 0) {
    GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.owner_off.p, 0, sizeof(int64_t) * (n_owners + 3), ctx->stream));
    if (defer_pin) memset(defer_pin, 0, sizeof(int64_t) * (size_t)(n_owners + 3));
    return GCI_OK;
  }
  if (n_chunks >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "too many run chunks");
  if (lay != ctx->lay_cache) {   // the chunk layout only changes with the contig table / flank / regions
    GCI_TRY(gci_h2d(ctx, ctx->chunk_off, lay.data(), sizeof(int64_t) * lay.size()));
    ctx->lay_cache = lay;
  }
  const int64_t* d_chunk_off = ctx->chunk_off.as<int64_t>();
  const int64_t* d_w0 = d_chunk_off + n_owners + 1;
  const int64_t* d_g0 = d_w0 + n_owners;
  const int32_t* d_chunk_owner = reinterpret_cast<const int32_t*>(d_chunk_off + lay_head);
  GCI_TRY(ctx->ensure(ctx->chunk_cnt, sizeof(int2) * (size_t)n_chunks));
  GCI_TRY(ctx->ensure(ctx->scan_tmp, sizeof(longlong2) * (size_t)n_chunks));
  GCI_TRY(ctx->ensure(ctx->run_stage, sizeof(int32_t) * 2 * RUN_STAGE * (size_t)n_chunks));
  int32_t* stage = ctx->run_stage.as<int32_t>();
  int2* cnt = ctx->chunk_cnt.as<int2>();
  longlong2* off = ctx->scan_tmp.as<longlong2>();
  if (t.iv_cap < 4096) {
    GCI_TRY(ctx->ensure(t.iv_start, sizeof(int32_t) * 4096));
    GCI_TRY(ctx->ensure(t.iv_end, sizeof(int32_t) * 4096));
    t.iv_cap = 4096;
  }
  RunArgs ra;
  ra.flags = t.flags.as<uint32_t>();
  ra.chunk_owner = d_chunk_owner;
  ra.chunk_off = d_chunk_off;
  ra.w_lo = t.win_lo.as<int64_t>();
  ra.w_hi = t.win_hi.as<int64_t>();
  ra.lay_w0 = d_w0;
  ra.lay_g0 = d_g0;
  ra.cnt = cnt;
  ra.off = off;
  ra.iv_start = nullptr;
  ra.iv_end = nullptr;
  ra.cap = 0;
  ra.stage = stage;
  ra.n_chunks = n_chunks;
  runs_count_kernel<<<(unsigned)n_chunks, RUN_THREADS, 0, ctx->stream>>>(ra);
  GCI_LAUNCH_CHECK(ctx);
  runs_scan_kernel<<<1, 1024, 0, ctx->stream>>>(cnt, n_chunks, off, n_owners, d_chunk_off, t.owner_off.as<int64_t>());
  GCI_LAUNCH_CHECK(ctx);
  auto emit = [&]() -> int {
    runs_unstage_kernel<<<(unsigned)((n_chunks + 255) / 256), 256, 0, ctx->stream>>>(
        cnt, off, n_chunks, stage, t.iv_start.as<int32_t>(), t.iv_end.as<int32_t>(), t.iv_cap);
    GCI_LAUNCH_CHECK(ctx);
    ra.iv_start = t.iv_start.as<int32_t>();
    ra.iv_end = t.iv_end.as<int32_t>();
    ra.cap = t.iv_cap;
    runs_write_kernel<<<(unsigned)std::min<int64_t>((n_chunks + 31) / 32, (int64_t)ctx->sm_count * 8), RUN_THREADS, 0, ctx->stream>>>(ra);
    GCI_LAUNCH_CHECK(ctx);
    return GCI_OK;
  };
  if (defer_pin) {
    GCI_TRY(emit());
    GCI_TRY(gci_d2h(ctx, defer_pin, t.owner_off.p, sizeof(int64_t) * (size_t)(n_owners + 3)));
    return GCI_OK;
  }
  int64_t* h = (int64_t*)ctx->pinned(sizeof(int64_t) * (size_t)(n_owners + 3));
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  for (int attempt = 0; attempt < 2; attempt++) {
    GCI_TRY(emit());
    if (attempt == 0) {
      GCI_TRY(gci_d2h(ctx, h, t.owner_off.p, sizeof(int64_t) * (size_t)(n_owners + 3)));
    }
    GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[n_owners + 1] != h[n_owners + 2])
      return ctx->fail(GCI_E_CUDA, "internal: run starts (%lld) != run ends (%lld)", (long long)h[n_owners + 1],
                       (long long)h[n_owners + 2]);
    if (h[n_owners + 1] <= t.iv_cap) break;
    // more intervals than the buffers hold: grow once and write again
    const int64_t want = h[n_owners + 1] + h[n_owners + 1] / 4;
    GCI_TRY(ctx->ensure(t.iv_start, sizeof(int32_t) * (size_t)want));
    GCI_TRY(ctx->ensure(t.iv_end, sizeof(int32_t) * (size_t)want));
    t.iv_cap = want;
  }
  t.n_intervals = h[n_owners + 1];
  t.h_owner_off.assign(h, h + n_owners + 1);
  if (n_intervals) *n_intervals = t.n_intervals;
  return GCI_OK;
}

// ================================================================================================
// K8  score terms: complement lengths, -dp merge count, N50
// ================================================================================================
// For owner o with window [S, E) and sorted disjoint intervals (s_i, e_i):
//   gaps g_0 = s_0 - S, g_i = s_i - e_{i-1}, tail = E - e_last            (GCI.py:446-458)
//   complement lengths = {g_i > 0} + {tail > 0}            or  [E - S] if there is no interval (:460)
//   -dp merge (GCI.py:509-518) joins neighbours across every gap <= dist, starting from the sentinel
//   (S,S) and ending at E, so the merged complement keeps exactly the gaps that are > dist (and > 0).
//   N50 (GCI.py:473-479) = largest length v with  2 * sum{len >= v} >= total.
// Result buffer (int64): [ n50 (owners+1) | n_ctg (owners) | depth sum (owners) | gap slots (n_intervals + owners) ];
// owner o uses gap slots [owner_off[o] + o, owner_off[o+1] + o + 1).
struct OwnerBounds { long long S, E; double dist; };

__global__ void __launch_bounds__(256)
complement_kernel(int64_t n_owners, const int64_t* __restrict__ owner_off, const int32_t* __restrict__ iv_start,
                  const int32_t* __restrict__ iv_end, const OwnerBounds* __restrict__ ob, int64_t* __restrict__ res,
                  const long long* __restrict__ sums, const int32_t* __restrict__ owner_contig, int64_t iv_cap) {
  __shared__ long long s_red[8];
  const int64_t o = blockIdx.x;
  if (owner_off[n_owners] > iv_cap) return;   // more intervals than were stored: the host redoes the scan
  const int64_t a = owner_off[o], b = owner_off[o + 1];
  int64_t* g = res + (3 * n_owners + 1) + a + o;
  const int64_t n = b - a;
  if (threadIdx.x == 0) res[2 * n_owners + 1 + o] = sums ? sums[owner_contig[o]] : 0;   // depth sum of the contig
  const long long S = ob[o].S, E = ob[o].E;
  const double d = ob[o].dist;
  long long my_ctg = 0;
  if (n == 0) {
    if (threadIdx.x == 0) {
      g[0] = E - S;                                         // appended unconditionally (:460)
      // merged = [(S,S)] then the tail rule: (E - S) <= dist -> one segment (S,E) -> empty complement
      my_ctg = ((double)(E - S) <= d) ? 0 : (E > S ? 1 : 0);
    }
  } else {
    for (int64_t i = threadIdx.x; i <= n; i += blockDim.x) {
      long long v;
      if (i < n) {
        const long long prev = i == 0 ? S : (long long)iv_end[a + i - 1];
        v = (long long)iv_start[a + i] - prev;
      } else {
        v = E - (long long)iv_end[b - 1];
      }
      const bool emit = v > 0;
      g[i] = emit ? v : 0;
      my_ctg += (emit && (double)v > d);
    }
  }
  my_ctg = warp_sum_ll(my_ctg);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = my_ctg;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tsum = 0;
    for (int j = 0; j < 8; j++) tsum += s_red[j];
    res[n_owners + 1 + o] = tsum;
  }
}

// N50 of the positive entries of one owner's gap slots (block o < n_owners) or of all slots (block n_owners)
constexpr int N50_SMALL = 1024;
constexpr int N50_REG = 32;        // values per thread of the register-resident selection (256 threads: 8192 values)
__global__ void __launch_bounds__(256)
n50_kernel(int64_t n_owners, const int64_t* __restrict__ owner_off, int64_t iv_cap, int64_t* __restrict__ res) {
  __shared__ long long s_red[8];
  if (owner_off[n_owners] > iv_cap) return;
  const int64_t n_slots = owner_off[n_owners] + n_owners;
  __shared__ long long s_bcast;
  __shared__ long long s_val[N50_SMALL];
  const int64_t o = blockIdx.x;
  const int64_t* vals = res + (3 * n_owners + 1);
  const int64_t lo = o < n_owners ? owner_off[o] + o : 0;
  const int64_t hi = o < n_owners ? owner_off[o + 1] + o + 1 : n_slots;
  const int64_t n = hi - lo;
  auto block_sum = [&](long long v) -> long long {
    v = warp_sum_ll(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int j = 0; j < 8; j++) t += s_red[j];
      s_bcast = t;
    }
    __syncthreads();
    return s_bcast;
  };
  auto block_max = [&](long long v) -> long long {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int j = 0; j < 8; j++) t = max(t, s_red[j]);
      s_bcast = t;
    }
    __syncthreads();
    return s_bcast;
  };
  long long best = 0;
  if (n <= N50_SMALL) {
    // small list: every thread sums the lengths >= its own (shared-memory broadcast reads)
    long long part = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const long long v = vals[lo + i] > 0 ? vals[lo + i] : 0;
      s_val[i] = v;
      part += v;
    }
    const long long total = block_sum(part);
    long long cand = 0;
    if (total > 0) {
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const long long v = s_val[i];
        if (v <= cand) continue;
        long long s = 0;
        for (int64_t j = 0; j < n; j++) s += s_val[j] >= v ? s_val[j] : 0;
        if (2 * s >= total) cand = v;
      }
    }
    best = block_max(cand);
  } else if (n <= (int64_t)N50_REG * 256 && blockDim.x == 256) {
    // medium list (the genome row of a few thousand gaps): every thread keeps its values in registers (lengths are
    // below 2^31), the 32 rounds of the bitwise selection then cost one barrier each (partials double-buffered)
    // instead of a walk over global memory and three barriers (34 us in the launch list for 4 564 slots before;
    // the score stage went from 0.059 to 0.041 ms, profiles/r02bj)
    __shared__ long long s_fast[2][8];
    uint32_t rv[N50_REG];
    long long part = 0;
#pragma unroll
    for (int k = 0; k < N50_REG; k++) {
      const int64_t i = lo + threadIdx.x + (int64_t)k * 256;
      const long long v = i < hi ? vals[i] : 0;
      rv[k] = v > 0 ? (uint32_t)v : 0u;
      part += rv[k];
    }
    const long long total = block_sum(part);
    if (total > 0) {
      uint32_t b = 0;
      for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = b | (1u << bit);
        long long p = 0;
#pragma unroll
        for (int k = 0; k < N50_REG; k++) p += rv[k] >= cand ? rv[k] : 0u;
        p = warp_sum_ll(p);
        if ((threadIdx.x & 31) == 0) s_fast[bit & 1][threadIdx.x >> 5] = p;
        __syncthreads();
        long long sum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) sum += s_fast[bit & 1][j];
        if (2 * sum >= total) b = cand;
      }
      best = (long long)b;
    }
  } else {
    long long part = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) part += vals[i] > 0 ? vals[i] : 0;
    const long long total = block_sum(part);
    if (total > 0) {
      for (int bit = 31; bit >= 0; bit--) {
        const long long cand = best | (1ll << bit);
        part = 0;
        for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) part += vals[i] >= cand ? vals[i] : 0;
        const long long s = block_sum(part);
        if (2 * s >= total) best = cand;
      }
    }
  }
  if (threadIdx.x == 0) res[o] = best;
}

// score kernels of one track's last scan; the result buffer stays on the device in ctx->tmp[1].
// pending: the scan's interval count has not reached the host yet -> sizes are bounded by the capacity.
int gci_launch_score_kernels(gci_ctx* ctx, Track& t, double dist_percent, int32_t flank_len, int64_t* n_owners,
                             int64_t* n_slots_out, bool pending) {
  const int64_t no = t.n_owners;
  if (no <= 0 || (!pending && (int64_t)t.h_owner_off.size() != no + 1))
    return ctx->fail(GCI_E_ARG, "score terms: no scan on this track");
  const int64_t n_slots = (pending ? t.iv_cap : t.n_intervals) + no;
  const int64_t n_res = 3 * no + 1 + n_slots;
  // per-owner [S, E) and dist (GCI.py:505-508, :629-634); uploaded only when they change
  std::vector<OwnerBounds> ob((size_t)no);
  for (int64_t o = 0; o < no; o++) {
    if (t.owners_are_windows) {
      // complement / merge use the region bounds as given (GCI.py:629-634), only the depth slice is
      // normalised (:627)
      ob[o].S = t.raw_lo[o];
      ob[o].E = t.raw_hi[o];
      ob[o].dist = (double)(ob[o].E - ob[o].S) * dist_percent;      // targets_length = {target: exp_n50}
    } else {
      const int64_t L = ctx->len[t.owner_contig[o]];
      ob[o].S = flank_len;
      ob[o].E = L - flank_len;
      ob[o].dist = (double)L * dist_percent;
    }
  }
  DevBuf &d_ob = ctx->tmp[0], &d_res = ctx->tmp[1];
  const size_t ob_bytes = sizeof(OwnerBounds) * (size_t)no;
  if (ctx->ob_cache.size() != ob_bytes || memcmp(ctx->ob_cache.data(), ob.data(), ob_bytes) != 0) {
    GCI_TRY(gci_h2d(ctx, d_ob, ob.data(), ob_bytes));
    GCI_CUDA_TRY(ctx, cudaEventRecord(ctx->h2d_done, ctx->stream));
    GCI_CUDA_TRY(ctx, cudaEventSynchronize(ctx->h2d_done));   // `ob` is a local: the copy must have left the host
    ctx->ob_cache.assign((const char*)ob.data(), (const char*)ob.data() + ob_bytes);
  }
  GCI_TRY(ctx->ensure(d_res, sizeof(int64_t) * (size_t)n_res));
  // pending: the copy back is sized by the capacity, not by the (still unknown) interval count
  if (pending) GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_res.p, 0, sizeof(int64_t) * (size_t)n_res, ctx->stream));
  complement_kernel<<<(unsigned)no, 256, 0, ctx->stream>>>(
      no, t.owner_off.as<int64_t>(), t.iv_start.as<int32_t>(), t.iv_end.as<int32_t>(), d_ob.as<OwnerBounds>(),
      d_res.as<int64_t>(), (!t.owners_are_windows && t.sums_valid) ? t.sums.as<long long>() : nullptr,
      t.win_contig.as<int32_t>(), t.iv_cap);
  GCI_LAUNCH_CHECK(ctx);
  n50_kernel<<<(unsigned)(no + 1), 256, 0, ctx->stream>>>(no, t.owner_off.as<int64_t>(), t.iv_cap, d_res.as<int64_t>());
  GCI_LAUNCH_CHECK(ctx);
  *n_owners = no;
  *n_slots_out = n_slots;
  return GCI_OK;
}

// whole-genome scan, enqueue only (gci_pipeline): windows kernel + run extraction, results into `pin`
int gci_scan_enqueue(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* pin);
int gci_scan_finish(gci_ctx* ctx, int32_t track, const int64_t* pin, bool* overflow);

extern "C" {

static int scan_genome(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* n_intervals,
                       int64_t* defer_pin);

int gci_scan(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* n_intervals) {
  if (ctx) ctx->epoch++;
  return scan_genome(ctx, track, lo, hi, flank_len, n_intervals, nullptr);
}

}  // extern "C"

int gci_scan_enqueue(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* pin) {
  return scan_genome(ctx, track, lo, hi, flank_len, nullptr, pin);
}

int gci_scan_finish(gci_ctx* ctx, int32_t track, const int64_t* pin, bool* overflow) {
  Track& t = ctx->track[track];
  const int64_t no = t.n_owners;
  *overflow = false;
  if (pin[no + 1] != pin[no + 2])
    return ctx->fail(GCI_E_CUDA, "internal: run starts (%lld) != run ends (%lld)", (long long)pin[no + 1],
                     (long long)pin[no + 2]);
  if (pin[no + 1] > t.iv_cap) {
    *overflow = true;
    return GCI_OK;
  }
  t.n_intervals = pin[no + 1];
  t.h_owner_off.assign(pin, pin + no + 1);
  return GCI_OK;
}

static int scan_genome(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* n_intervals,
                       int64_t* defer_pin) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_scan: track %d holds no depth", track);
  if (!t.flags_valid || t.flags_lo != lo || t.flags_hi != hi) GCI_TRY(gci_compute_flags(ctx, track, lo, hi));
  ctx->stage_begin(GCI_ST_RUNS);
  // owners = selected contigs in header order; the owner map is cached on the device per contig table
  int64_t n_owners = 0;
  for (int c = 0; c < ctx->n_contigs; c++) n_owners += ctx->selected[c] ? 1 : 0;
  if (ctx->d_owner_of.cap == 0 || ctx->owner_of_stale) {
    std::vector<int32_t> owner_of(std::max(1, ctx->n_contigs), -1);
    int32_t k = 0;
    for (int c = 0; c < ctx->n_contigs; c++)
      if (ctx->selected[c]) owner_of[c] = k++;
    GCI_TRY(gci_h2d(ctx, ctx->d_owner_of, owner_of.data(), sizeof(int32_t) * owner_of.size()));
    ctx->owner_of_stale = false;
  }
  t.n_owners = n_owners;
  t.owners_are_windows = false;
  t.scan_flank = flank_len;
  GCI_TRY(ctx->ensure(t.win_contig, sizeof(int32_t) * (size_t)std::max<int64_t>(1, n_owners)));
  GCI_TRY(ctx->ensure(t.win_lo, sizeof(int64_t) * (size_t)std::max<int64_t>(1, n_owners)));
  GCI_TRY(ctx->ensure(t.win_hi, sizeof(int64_t) * (size_t)std::max<int64_t>(1, n_owners)));
  // host-known layout bounds: the real window start is in [fl, 2*fl], its end is L - fl
  std::vector<int64_t> lay_lo(n_owners), h_hi(n_owners);
  t.owner_contig.assign(n_owners, 0);
  {
    int64_t o = 0;
    for (int c = 0; c < ctx->n_contigs; c++) {
      if (!ctx->selected[c]) continue;
      const int64_t L = ctx->len[c];
      const bool ok = flank_len >= 0 && L - 2 * (int64_t)flank_len > 0;
      lay_lo[o] = ok ? flank_len : 0;
      h_hi[o] = ok ? L - flank_len : 0;
      t.owner_contig[o] = c;
      o++;
    }
  }
  int rc = GCI_OK;
  if (n_owners) {
    genome_windows_kernel<<<(ctx->n_contigs + 127) / 128, 128, 0, ctx->stream>>>(
        ctx->n_contigs, ctx->d_len.as<int64_t>(), ctx->d_selected.as<uint8_t>(), ctx->d_tile_off.as<int64_t>(),
        t.flags.as<uint32_t>(), flank_len, t.win_contig.as<int32_t>(), t.win_lo.as<int64_t>(), t.win_hi.as<int64_t>(),
        ctx->d_owner_of.as<int32_t>());
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = ctx->fail(GCI_E_CUDA, "genome_windows_kernel launch failed");
  }
  if (rc == GCI_OK) rc = extract_runs(ctx, t, lay_lo, h_hi, n_intervals, defer_pin);
  ctx->stage_end();
  return rc;
}

extern "C" {

int gci_scan_windows(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int64_t n_windows, const int32_t* contig,
                     const int64_t* start, const int64_t* end, int64_t* n_intervals) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || n_windows < 0) return GCI_E_ARG;
  if (n_windows && (!contig || !start || !end)) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_scan_windows: track %d holds no depth", track);
  if (!t.flags_valid || t.flags_lo != lo || t.flags_hi != hi) GCI_TRY(gci_compute_flags(ctx, track, lo, hi));
  // depth[start:end] is a Python slice (GCI.py:627): normalise against the contig length
  std::vector<int32_t> c2(n_windows);
  std::vector<int64_t> lo2(n_windows), hi2(n_windows);
  for (int64_t i = 0; i < n_windows; i++) {
    const int32_t c = contig[i];
    if (c < 0 || c >= ctx->n_contigs || !ctx->selected[c])
      return ctx->fail(GCI_E_ARG, "window %lld: contig %d is not selected", (long long)i, c);
    const int64_t L = ctx->len[c];
    auto norm = [L](int64_t v) {
      if (v < 0) { v += L; if (v < 0) v = 0; } else if (v >= L) v = L;
      return v;
    };
    c2[i] = c;
    lo2[i] = norm(start[i]);
    hi2[i] = std::max(lo2[i], norm(end[i]));
  }
  t.raw_lo.assign(start, start + n_windows);
  t.raw_hi.assign(end, end + n_windows);
  t.owner_contig.assign(c2.begin(), c2.end());
  ctx->stage_begin(GCI_ST_RUNS);
  t.n_owners = n_windows;
  t.owners_are_windows = true;
  t.scan_flank = 0;
  GCI_TRY(gci_h2d(ctx, t.win_contig, c2.data(), sizeof(int32_t) * n_windows));
  GCI_TRY(gci_h2d(ctx, t.win_lo, lo2.data(), sizeof(int64_t) * n_windows));
  GCI_TRY(gci_h2d(ctx, t.win_hi, hi2.data(), sizeof(int64_t) * n_windows));
  int rc = extract_runs(ctx, t, lo2, hi2, n_intervals);
  ctx->stage_end();
  return rc;
}

int gci_load_intervals(gci_ctx* ctx, int32_t track, int64_t n_owners, const int32_t* contig, const int64_t* owner_off,
                       const int32_t* start, const int32_t* end) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || n_owners < 0 || !owner_off) return GCI_E_ARG;
  if (n_owners && !contig) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  Track& t = ctx->track[track];
  const int64_t n = owner_off[n_owners];
  if (n && (!start || !end)) return GCI_E_ARG;
  for (int64_t o = 0; o < n_owners; o++)
    if (contig[o] < 0 || contig[o] >= ctx->n_contigs) return ctx->fail(GCI_E_ARG, "bad contig in gci_load_intervals");
  t.n_owners = n_owners;
  t.n_intervals = n;
  t.owners_are_windows = false;
  t.owner_contig.assign(contig, contig + n_owners);
  t.h_owner_off.assign(owner_off, owner_off + n_owners + 1);
  GCI_TRY(ctx->ensure(t.owner_off, sizeof(int64_t) * (size_t)(n_owners + 3)));
  GCI_TRY(gci_h2d(ctx, t.win_contig, contig, sizeof(int32_t) * n_owners));
  GCI_CUDA_TRY(ctx, cudaMemcpyAsync(t.owner_off.p, owner_off, sizeof(int64_t) * (n_owners + 1), cudaMemcpyHostToDevice,
                                    ctx->stream));
  GCI_TRY(gci_h2d(ctx, t.iv_start, start, sizeof(int32_t) * n));
  GCI_TRY(gci_h2d(ctx, t.iv_end, end, sizeof(int32_t) * n));
  t.iv_cap = std::max<int64_t>(t.iv_cap, n);
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_score_terms(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len, int64_t* n50,
                    int64_t* n_ctg, int64_t cap_lengths, int64_t* lengths, int64_t* lengths_off) {
  return gci_score_terms_sums(ctx, track, dist_percent, flank_len, n50, n_ctg, cap_lengths, lengths, lengths_off,
                              nullptr);
}

int gci_score_terms_sums(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len, int64_t* n50,
                         int64_t* n_ctg, int64_t cap_lengths, int64_t* lengths, int64_t* lengths_off,
                         int64_t* depth_sums) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  Track& t = ctx->track[track];
  int64_t no = 0, n_slots = 0;
  ctx->stage_begin(GCI_ST_SCORE);
  GCI_TRY(gci_launch_score_kernels(ctx, t, dist_percent, flank_len, &no, &n_slots, false));
  const int64_t n_res = 3 * no + 1 + n_slots;
  int64_t* h_res = (int64_t*)ctx->pinned(sizeof(int64_t) * (size_t)n_res);
  if (!h_res) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_TRY(gci_d2h(ctx, h_res, ctx->tmp[1].p, sizeof(int64_t) * (size_t)n_res));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (n50) memcpy(n50, h_res, sizeof(int64_t) * (size_t)(no + 1));
  if (n_ctg) {
    int64_t all = 0;
    for (int64_t o = 0; o < no; o++) { n_ctg[o] = h_res[no + 1 + o]; all += n_ctg[o]; }
    n_ctg[no] = all;
  }
  if (depth_sums) {
    if (t.owners_are_windows || !t.sums_valid)
      return ctx->fail(GCI_E_ARG, "depth sums ride along only for a whole-contig scan of a track with valid sums");
    int64_t all = 0;
    for (int64_t o = 0; o < no; o++) { depth_sums[o] = h_res[2 * no + 1 + o]; all += depth_sums[o]; }
    depth_sums[no] = all;
  }
  // compact the emitted lengths per owner, in position order (zero slots were not emitted, except the
  // unconditional [E - S] entry of an owner without intervals)
  int rc = GCI_OK;
  if (lengths_off || lengths) {
    const int64_t* h_gaps = h_res + 3 * no + 1;
    const std::vector<int64_t>& h_off = t.h_owner_off;
    int64_t k = 0;
    for (int64_t o = 0; o < no && rc == GCI_OK; o++) {
      if (lengths_off) lengths_off[o] = k;
      const int64_t a = h_off[o] + o, b = h_off[o + 1] + o + 1;
      const bool empty = (h_off[o + 1] == h_off[o]);
      for (int64_t i = a; i < b; i++) {
        if (h_gaps[i] > 0 || empty) {
          if (lengths) {
            if (k >= cap_lengths) { rc = ctx->fail(GCI_E_ARG, "lengths buffer too small"); break; }
            lengths[k] = h_gaps[i];
          }
          k++;
        }
      }
    }
    if (lengths_off && rc == GCI_OK) lengths_off[no] = k;
  }
  return rc;
}

}  // extern "C"

// ================================================================================================
// plot feed: sliding_window_average_depth (GCI.py:660-705) on the device
// ================================================================================================
// The reference walks one contig (or region) base by base: every zero-depth base is a point of its own and closes the
// open window; non-zero bases fill windows of `window_size`; a window still open at a zero or at the end is averaged
// over what it holds.  Here: the zero runs of the region come from the run extraction above (bit = depth == 0), the
// region is then a sequence of SEGMENTS nz0 z0 nz1 z1 ... nz_last; a zero segment of length m emits m points, a
// non-zero one floor(m / ws) full windows + one short window for the rest.  Points are (index, sum, count, kind);
// the division, the clip at max_depth and the Mbp positions stay on the host (Python float semantics).
struct SwArgs {
  const int32_t* iv_start;   // zero runs [start, end) in contig coordinates, sorted
  const int32_t* iv_end;
  int64_t n_iv;
  int64_t lo, hi;            // region [lo, hi) in contig coordinates
  int64_t ws;
};

// segment k: even = non-zero stretch before zero run k/2 (or the tail), odd = zero run (k-1)/2
__device__ __forceinline__ void sw_segment(const SwArgs& a, int64_t k, int64_t& s, int64_t& e) {
  const int64_t j = k >> 1;
  if (k & 1) { s = a.iv_start[j]; e = a.iv_end[j]; return; }
  s = j == 0 ? a.lo : (int64_t)a.iv_end[j - 1];
  e = j == a.n_iv ? a.hi : (int64_t)a.iv_start[j];
}

__global__ void sw_count_kernel(SwArgs a, int64_t n_seg, int32_t* __restrict__ cnt) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n_seg) return;
  int64_t s, e;
  sw_segment(a, k, s, e);
  const int64_t m = max((int64_t)0, e - s);
  cnt[k] = (int32_t)((k & 1) ? m : m / a.ws + (m % a.ws ? 1 : 0));
}

__global__ void sw_emit_kernel(SwArgs a, int64_t n_seg, const int64_t* __restrict__ seg_off, int64_t n_points,
                               const int64_t* __restrict__ csum /* exclusive prefix of depth over [lo, hi) */,
                               int64_t* __restrict__ o_idx, int64_t* __restrict__ o_num, int64_t* __restrict__ o_den,
                               uint8_t* __restrict__ o_kind) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  const int64_t k = upper_bound_minus1<int64_t>(seg_off, n_seg, p);
  // (segments without points share their offset with the next one: step to the last segment starting at or before p
  // that really holds it)
  int64_t s, e;
  sw_segment(a, k, s, e);
  const int64_t r = p - seg_off[k];
  if (k & 1) {                                     // a zero base: a point of its own, value 0
    o_idx[p] = s + r - a.lo; o_num[p] = 0; o_den[p] = 1; o_kind[p] = 2;
    return;
  }
  const int64_t w0 = s + r * a.ws, w1 = min(e, w0 + a.ws);
  o_idx[p] = w1 - 1 - a.lo;                        // reported at the window's last base
  o_num[p] = csum[w1 - a.lo] - csum[w0 - a.lo];
  o_den[p] = w1 - w0;
  o_kind[p] = (w1 - w0 == a.ws) ? 0 : 1;
}

extern "C" int gci_sliding_window(gci_ctx* ctx, int32_t track, int32_t contig, int64_t start, int64_t end,
                                  int64_t window_size, int64_t cap, int64_t* idx, int64_t* num, int64_t* den,
                                  uint8_t* kind, int64_t* n_points) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS || !n_points || window_size < 1) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "gci_sliding_window: track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || start < 0 || end < start || end > ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_sliding_window: bad region");
  *n_points = 0;
  const int64_t n = end - start;
  if (n == 0) return GCI_OK;
  // zero runs of the region: depth == 0  <=>  -1 < depth <= 0 for the non-negative depths of a track
  int64_t n_iv = 0;
  const int64_t cs = start, ce = end;
  GCI_TRY(gci_scan_windows(ctx, track, -1, 0, 1, &contig, &cs, &ce, &n_iv));
  const int64_t n_seg = 2 * n_iv + 1;
  DevBuf &cnt = ctx->tmp[6], &off = ctx->tmp[7], &csum = ctx->sw_csum, &out = ctx->sw_out;
  GCI_TRY(ctx->ensure(cnt, 4 * (size_t)(n_seg + 1)));
  GCI_TRY(ctx->ensure(off, 8 * (size_t)(n_seg + 2)));
  GCI_TRY(ctx->ensure(csum, 8 * (size_t)(n + 1)));
  SwArgs a{t.iv_start.as<int32_t>(), t.iv_end.as<int32_t>(), n_iv, start, end, window_size};
  sw_count_kernel<<<(unsigned)((n_seg + 255) / 256), 256, 0, ctx->stream>>>(a, n_seg, cnt.as<int32_t>());
  GCI_LAUNCH_CHECK(ctx);
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(cnt.as<int32_t>() + n_seg, 0, 4, ctx->stream));
  GCI_TRY(gci_exclusive_scan_i64_from_i32(ctx, cnt.as<int32_t>(), off.as<int64_t>(), n_seg + 1, nullptr));
  int64_t* h = (int64_t*)ctx->pinned(8);
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_TRY(gci_d2h(ctx, h, off.as<int64_t>() + n_seg, 8));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t P = *h;
  *n_points = P;
  if (!idx && !num && !den && !kind) return GCI_OK;
  if (cap < P) return ctx->fail(GCI_E_ARG, "gci_sliding_window: %lld points, buffers hold %lld", (long long)P, (long long)cap);
  if (P == 0) return GCI_OK;
  const int32_t* d = t.depth.as<int32_t>() + ctx->pos_off[contig] + start;
  GCI_TRY(gci_exclusive_scan_i64_from_i32(ctx, d, csum.as<int64_t>(), n, csum.as<int64_t>() + n));
  GCI_TRY(ctx->ensure(out, 25 * (size_t)P));
  int64_t* o_idx = out.as<int64_t>();
  int64_t* o_num = o_idx + P;
  int64_t* o_den = o_num + P;
  uint8_t* o_kind = reinterpret_cast<uint8_t*>(o_den + P);
  sw_emit_kernel<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(a, n_seg, off.as<int64_t>(), P, csum.as<int64_t>(),
                                                                      o_idx, o_num, o_den, o_kind);
  GCI_LAUNCH_CHECK(ctx);
  if (idx) GCI_TRY(gci_d2h(ctx, idx, o_idx, 8 * (size_t)P));
  if (num) GCI_TRY(gci_d2h(ctx, num, o_num, 8 * (size_t)P));
  if (den) GCI_TRY(gci_d2h(ctx, den, o_den, 8 * (size_t)P));
  if (kind) GCI_TRY(gci_d2h(ctx, kind, o_kind, (size_t)P));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

