// Read-set sharding over the GPUs of one NVLink domain (the reference's analogue is the Pool fan-out
// plus dict merge of GCI.py:257-301).
//
// Contigs are owned by ranks (depth, scan and score are per contig), but the cross-file join (GCI.py:272-301) is
// keyed by READ: a read's records may sit on contigs of different owners.  Instead of gathering every table on every
// rank, every read has a HOME rank (block-cyclic in the read id, home_rank()) and the step moves each piece of data
// exactly once:
//
//   owner of the contig      BAM gates + last-record-wins dedup on the records it holds
//        |   dispatch 1      every per-file winner row (contig, start, end, qlen, high-quality mark) is stored
//        v                   straight into the inbox of the read's home over NVLink peer memory
//   home of the read         PAF election for its reads (the host deals PAF lines by read) while the rows travel, then
//        |                   the join of all files for its reads (of the rows several owners sent for one read the
//        |                   highest contig wins, :268-270)
//        |   dispatch 2      every survivor (contig, start, end) is stored into the inbox of the contig's owner
//        v
//   owner of the contig      depth events -> depth tiles -> scan -> score, as on one GPU
//
// Both dispatches are all-to-all exchanges written as plain kernels over peer pointers (CUDA IPC between processes):
// the row of a read goes to a fixed slot (source rank, home-local read id) and the step's epoch, stored beside it,
// is its validity tag, so nothing is counted, claimed or cleared; every rank then raises an epoch flag in the peers'
// headers and waits for theirs.  Two receive areas alternate by epoch parity (a rank can be at most one step ahead of a
// peer), the epoch lives in device memory, so the step replays inside a CUDA graph.
#include <algorithm>

#include "common.cuh"

// 16 B: a file's winner for one read; the high-quality mark of the read rides in bit 31 of the query length.  One
// 16-byte store per lane: a warp whose reads share a home writes 512 contiguous bytes there.
struct XRow1 { int32_t contig, start, end; uint32_t qlen_hq; };
static_assert(sizeof(XRow1) == 16, "winner rows are stored and loaded as int4");
struct XRow2 { int32_t contig, start, end, pad; };                                // 16 B: a survivor
// Winner slots are DENSE: the row of home read h (= home_local(read id)) sent by rank `src` lives at [src][h], so a
// sender needs no slot counter and the receiver no merge pass; a row is valid when its tag (tags1, an array of its
// own) equals the step's epoch (the areas are never cleared: a stale row carries an older epoch).  Survivor rows are
// COMPACT per (destination, source): the owner of a contig must not scan world x home-reads slots to find the few
// that are his, so the home claims slots per destination (one atomic per CTA and destination) and publishes the
// counts with its flag.

// geometry of one rank's exchange area (identical on every rank)
struct XLayout {
  int world, files;
  long long cap1, cap2;          // home reads per rank (rows per source)
  // header, in 8-byte words: [epoch | flag[2 phases][2 parities][world] | cnt2[2 parities][world]]
  __host__ __device__ long long flag_word(int phase, int par, int src) const { return 1 + ((phase * 2 + par) * world + src); }
  __host__ __device__ long long cnt2_word(int par, int src) const { return 1 + 4LL * world + par * world + src; }
  __host__ __device__ long long rows1_base() const { return ((1 + 6LL * world) * 8 + 255) & ~255LL; }
  __host__ __device__ long long rows1_off(int par, int f, int src) const {
    return rows1_base() + (long long)sizeof(XRow1) * cap1 * ((par * files + f) * (long long)world + src);
  }
  __host__ __device__ long long rows2_base() const { return (rows1_off(2, 0, 0) + 255) & ~255LL; }
  __host__ __device__ long long rows2_off(int par, int src) const {
    return rows2_base() + (long long)sizeof(XRow2) * cap2 * (par * (long long)world + src);
  }
  // validity tags of the winner rows again, on their own ([par][file][src][cap1] u32): the home looks at world tags per
  // read and file but only one or two rows carry the step's epoch, so it reads 4 bytes per source instead of a
  // 32-byte sector of every source's row array
  __host__ __device__ long long tags1_base() const { return (rows2_off(2, 0) + 255) & ~255LL; }
  __host__ __device__ long long tags1_off(int par, int f, int src) const {
    return tags1_base() + 4LL * ((cap1 + 63) & ~63LL) * ((par * files + f) * (long long)world + src);
  }
  __host__ __device__ long long total() const { return tags1_off(2, 0, 0); }
};

struct XPeers { char* area[GCI_MAX_RANKS]; };
constexpr unsigned long long XCHG_TIMEOUT_NS = 8000000000ull;

// ---- step begin: advance the epoch ---------------------------------------------------------------------------
__global__ void xchg_begin_kernel(char* mine, uint32_t* send_cnt, int world) {
  if (threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(mine) += 1;
  if ((int)threadIdx.x < world) send_cnt[threadIdx.x] = 0;
}

// slot for one row per destination rank: rows of one CTA are counted in shared memory, one global atomic per
// (CTA, destination) claims the range.  Block-collective.
__device__ __forceinline__ uint32_t claim_slot(uint32_t* cursor /* [world] */, int dst, bool active, int world) {
  __shared__ uint32_t s_cnt[GCI_MAX_RANKS], s_base[GCI_MAX_RANKS];
  if (threadIdx.x < GCI_MAX_RANKS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  uint32_t local = 0;
  if (active) local = atomicAdd(&s_cnt[dst], 1u);
  __syncthreads();
  if ((int)threadIdx.x < world && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
  return active ? s_base[dst] + local : 0u;
}

// ---- dispatch 1: per-file winners to the home of their read --------------------------------------------------
// thread per LOCAL record (the work scales with what this rank holds, not with the global read count): the record
// that won its read (win[read] names it) sends the row.  Records are coordinate sorted and read ids follow first
// appearance, so consecutive records carry nearby ids and the lanes that talk to one destination write neighbouring
// 16-byte rows there.
__global__ void __launch_bounds__(256)
dispatch1_kernel(XLayout lay, XPeers peers, int me, int f, int64_t n_rec, uint32_t n_reads,
                 const uint32_t* __restrict__ read_id, const long long* __restrict__ win,
                 const int32_t* __restrict__ ref_id, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                 const int32_t* __restrict__ qlen, const uint8_t* __restrict__ highq) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rec) return;
  const uint32_t q = read_id[r];
  if (q >= n_reads) return;
  const long long k = win[q];
  if (k < 0 || (int64_t)(k & 0xffffffffll) != r) return;
  const unsigned long long epoch = *reinterpret_cast<const unsigned long long*>(peers.area[me]);
  const int par = (int)(epoch & 1ull);
  const uint32_t w = (uint32_t)lay.world;
  char* dst_area = peers.area[home_rank(q, w)];
  const uint32_t hl = home_local(q, w);
  reinterpret_cast<int4*>(dst_area + lay.rows1_off(par, f, me))[hl] =
      make_int4(ref_id[r], start[r], end[r], (int)((uint32_t)qlen[r] | (highq[q] ? 0x80000000u : 0u)));
  reinterpret_cast<uint32_t*>(dst_area + lay.tags1_off(par, f, me))[hl] = (uint32_t)epoch;
}

// ---- raise this rank's flag of one phase at every peer ---------------------------------------------------------
__global__ void xchg_signal_kernel(XLayout lay, XPeers peers, int me, int phase, const uint32_t* __restrict__ send_cnt) {
  const int dst = threadIdx.x;
  if (dst >= lay.world) return;
  const unsigned long long epoch = *reinterpret_cast<const unsigned long long*>(peers.area[me]);
  const int par = (int)(epoch & 1ull);
  volatile unsigned long long* hdr = reinterpret_cast<volatile unsigned long long*>(peers.area[dst]);
  __threadfence_system();                             // the rows of the kernels before this one are out
  if (phase == 1) {
    hdr[lay.cnt2_word(par, me)] = send_cnt[dst];      // survivors this rank sent to dst
    __threadfence_system();
  }
  hdr[lay.flag_word(phase, par, me)] = epoch;
}

// ---- wait until every source's flag of this phase shows the step's epoch ---------------------------------------
__global__ void xchg_wait_kernel(XLayout lay, char* mine, int phase, unsigned long long* __restrict__ err) {
  const int src = threadIdx.x;
  if (src >= lay.world) return;
  volatile unsigned long long* hdr = reinterpret_cast<volatile unsigned long long*>(mine);
  const unsigned long long epoch = hdr[0];
  const int par = (int)(epoch & 1ull);
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (hdr[lay.flag_word(phase, par, src)] < epoch) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > XCHG_TIMEOUT_NS) {                  // a peer never arrived: report instead of hanging the GPU
      atomicOr(err, 64ull);
      break;
    }
    __nanosleep(100);
  }
  __threadfence_system();
}

// ---- home: the join of all files for the reads homed here (GCI.py:268-301), survivors to their contig's owner ---
struct HomeFile {
  int bam;                        // >= 0: inbox rows of BAM upload `bam` (index among the BAM files); < 0: a table
  const long long* win;           // table: [n_home] entry of the home read, < 0 = absent
  const int32_t *c, *s, *e, *q;   // table columns
};
struct HomeArgs {
  HomeFile f[GCI_MAX_FILES];
  int n_files;
};

__global__ void __launch_bounds__(256)
home_join_kernel(XLayout lay, XPeers peers, int me, HomeArgs a, uint32_t n_home, const uint8_t* __restrict__ hq_home,
                 double op, const int32_t* __restrict__ owner, int32_t n_contigs, uint32_t* __restrict__ send_cnt,
                 unsigned long long* __restrict__ count, unsigned long long* __restrict__ err) {
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  const char* mine = peers.area[me];
  const unsigned long long epoch = *reinterpret_cast<const unsigned long long*>(mine);
  const int par = (int)(epoch & 1ull);
  const uint32_t tag = (uint32_t)epoch;
  bool have = false;
  int32_t c = -1, s = 0, e = 0;
  if (h < n_home) {
    long long k[GCI_MAX_FILES];
    JoinEntry x[GCI_MAX_FILES];
    bool hq = hq_home[h] != 0;
    for (int f = 0; f < a.n_files; f++) {
      const HomeFile& hf = a.f[f];
      k[f] = -1;
      x[f] = JoinEntry{-1, 0, 0, 0};
      if (hf.bam < 0) {
        k[f] = hf.win[h];
        if (k[f] >= 0) {
          const uint32_t i = (uint32_t)(k[f] & 0xffffffffll);
          x[f] = JoinEntry{hf.c[i], hf.s[i], hf.e[i], hf.q[i]};
        }
      } else {
        // one candidate row per source rank; the highest contig wins (the reference's fetch order, GCI.py:260-269),
        // the high-quality marks of every rank's records count (:167-168)
        // the tags of every source first (independent loads), then the rows of the one or two that sent
        uint32_t sent = 0;
#pragma unroll
        for (int src = 0; src < GCI_MAX_RANKS; src++)
          if (src < lay.world && reinterpret_cast<const uint32_t*>(mine + lay.tags1_off(par, hf.bam, src))[h] == tag)
            sent |= 1u << src;
        while (sent) {
          const int src = __ffs(sent) - 1;
          sent &= sent - 1;
          const int4 r = reinterpret_cast<const int4*>(mine + lay.rows1_off(par, hf.bam, src))[h];
          hq = hq || ((uint32_t)r.w >> 31) != 0;
          if (r.x > x[f].c) {
            x[f] = JoinEntry{r.x, r.y, r.z, (int32_t)((uint32_t)r.w & 0x7fffffffu)};
            k[f] = src;
          }
        }
      }
    }
    have = join_one(a.n_files, k, [&x](int f) { return x[f]; }, a.n_files > 1 && hq, op, home_global(h, (uint32_t)me, (uint32_t)lay.world),
                    err, c, s, e);
  }
  {
    const bool send = have && c >= 0 && c < n_contigs;
    const int dst = send ? owner[c] : 0;
    const uint32_t slot = claim_slot(send_cnt, dst, send, lay.world);
    if (send) {
      if ((long long)slot < lay.cap2) {
        XRow2* p = reinterpret_cast<XRow2*>(peers.area[dst] + lay.rows2_off(par, me)) + slot;
        *reinterpret_cast<int4*>(p) = make_int4(c, s, e, 0);
      } else {
        atomicOr(err, 128ull);                        // cannot happen: a home sends at most one survivor per home read
      }
    }
  }
  // survivors evaluated here (the global count is the sum over the ranks)
  const unsigned m = __ballot_sync(0xffffffffu, have);
  __shared__ int s_n[8];
  if ((threadIdx.x & 31) == 0) s_n[threadIdx.x >> 5] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) n += s_n[i];
    if (n) atomicAdd(count, (unsigned long long)n);
  }
}

// ---- owner: the survivors every home sent, packed into the survivor arrays, depth events counted on the way -------
// Rows arrive in the homes' read order, i.e. scattered over the contigs: the per-contig depth sums are therefore kept
// per CTA in a small shared-memory table (contig & 63, claimed by compare-and-swap; a clash goes straight to memory)
// and flushed once at the end.  (Sending the slice length of every row of a warp that straddles contigs to
// bk.sums[contig] — 24 hot addresses — was 0.25 ms of this kernel's 0.30 ms for 4.6 M rows: a skip-the-sums ablation
// ran in 0.046 ms, profiles/r02t.)  Contiguous row range per CTA, at most CONSUME_ROWS_PER_CTA rows so the 16-bit
// halves of the slice lengths add up in 32-bit shared atomics.
constexpr int CONSUME_SLOTS = 64;
constexpr long long CONSUME_ROWS_PER_CTA = 32768;

__global__ void __launch_bounds__(256)
consume2_kernel(XLayout lay, const char* __restrict__ mine, int32_t* __restrict__ s_contig, int32_t* __restrict__ s_start,
                int32_t* __restrict__ s_end, uint32_t* __restrict__ n_surv_dev, BucketArgs bk) {
  __shared__ long long s_off[GCI_MAX_RANKS + 1];
  __shared__ int32_t s_key[CONSUME_SLOTS];
  __shared__ uint32_t s_lo[CONSUME_SLOTS], s_hi[CONSUME_SLOTS];
  const unsigned long long* hdr = reinterpret_cast<const unsigned long long*>(mine);
  const int par = (int)(hdr[0] & 1ull);
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int r = 0; r < lay.world; r++) {
      s_off[r] = t;
      t += (long long)min((unsigned long long)lay.cap2, hdr[lay.cnt2_word(par, r)]);
    }
    s_off[lay.world] = t;
    if (blockIdx.x == 0) *n_surv_dev = (uint32_t)t;
  }
  if (threadIdx.x < CONSUME_SLOTS) { s_key[threadIdx.x] = -1; s_lo[threadIdx.x] = 0; s_hi[threadIdx.x] = 0; }
  __syncthreads();
  long long begin, end;
  cta_range(s_off[lay.world], blockDim.x, begin, end);
  if (begin >= end) return;                                                   // whole CTA
  int src = 0;
  for (long long base = begin; base < end; base += blockDim.x) {
    const long long g = base + threadIdx.x;
    int32_t c = -1;
    uint32_t covered = 0;
    if (g < end) {
      while (src + 1 < lay.world && g >= s_off[src + 1]) src++;
      const int4 r = *reinterpret_cast<const int4*>(mine + lay.rows2_off(par, src) + (long long)sizeof(XRow2) * (g - s_off[src]));
      s_contig[g] = r.x;
      s_start[g] = r.y;
      s_end[g] = r.z;
      const Slice sl = survivor_slice(r.x, r.y, r.z, bk.fl, bk.len, bk.tile_off);
      if (sl.ok) {
        atomicAdd(&bk.cnt_start[sl.tile_a], 1u);
        atomicAdd(&bk.cnt_end[sl.tile_b], 1u);
        covered = (uint32_t)(sl.b - sl.a);
        c = r.x;
      }
    }
    // slice length -> the CTA's table: one lane for a warp that agrees on the contig, every lane otherwise
    const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
    if (act == 0) continue;
    const int32_t c0 = __shfl_sync(0xffffffffu, c, __ffs(act) - 1);
    uint32_t lo = covered & 0xffffu, hi = covered >> 16;
    bool add = c >= 0;
    if (__all_sync(0xffffffffu, c < 0 || c == c0)) {
      lo = __reduce_add_sync(0xffffffffu, lo);
      hi = __reduce_add_sync(0xffffffffu, hi);
      add = (threadIdx.x & 31) == 0;
      c = c0;
    }
    if (add) {
      const int slot = c & (CONSUME_SLOTS - 1);
      int32_t k = s_key[slot];
      if (k == -1) {
        const int32_t old = atomicCAS(&s_key[slot], -1, c);
        k = old == -1 ? c : old;
      }
      if (k == c) {
        atomicAdd(&s_lo[slot], lo);
        atomicAdd(&s_hi[slot], hi);
      } else {
        atomicAdd((unsigned long long*)(bk.sums + c), (unsigned long long)lo + ((unsigned long long)hi << 16));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < CONSUME_SLOTS && s_key[threadIdx.x] >= 0) {
    const unsigned long long sum = (unsigned long long)s_lo[threadIdx.x] + ((unsigned long long)s_hi[threadIdx.x] << 16);
    if (sum) atomicAdd((unsigned long long*)(bk.sums + s_key[threadIdx.x]), sum);
  }
}

// ---- host side ----------------------------------------------------------------------------------------------
static XLayout make_layout(const gci_ctx* ctx) {
  XLayout l;
  l.world = ctx->shard.world;
  l.files = ctx->shard.max_files;
  l.cap1 = ctx->shard.cap1;
  l.cap2 = ctx->shard.cap2;
  return l;
}

static XPeers make_peers(const gci_ctx* ctx) {
  XPeers p;
  for (int r = 0; r < GCI_MAX_RANKS; r++) p.area[r] = r < ctx->shard.world ? (char*)ctx->shard.peer[r] : nullptr;
  return p;
}

struct XStep {
  XLayout lay;
  XPeers peers;
  char* mine;
  uint32_t* send_cnt;
  unsigned long long* d_err;
};

static int step_args(gci_ctx* ctx, XStep* x) {
  gci_ctx::Shard& sh = ctx->shard;
  if (!sh.on || !sh.opened) return ctx->fail(GCI_E_ARG, "sharded read set: gci_shard_open / gci_shard_attach missing");
  x->lay = make_layout(ctx);
  x->peers = make_peers(ctx);
  x->mine = (char*)sh.area.p;
  if ((int64_t)ctx->n_reads > x->lay.cap1 * x->lay.world)
    return ctx->fail(GCI_E_ARG, "sharded read set: %u reads, exchange sized for %lld", ctx->n_reads, x->lay.cap1 * x->lay.world);
  int n_bam = 0;
  for (size_t i = 0; i < ctx->n_files; i++) n_bam += ctx->files[i].kind == 0 ? 1 : 0;
  if (n_bam > sh.max_files) return ctx->fail(GCI_E_ARG, "sharded read set: %d BAM files, exchange sized for %d", n_bam, sh.max_files);
  x->d_err = ctx->d_err.as<unsigned long long>();
  x->send_cnt = sh.send_cnt.as<uint32_t>();
  return GCI_OK;
}

// Dispatch 1 of one step: needs the BAM legs only (win per global read, highq), so the pipeline enqueues it right
// after them and runs the PAF legs (home-local work) while the rows travel and the slower peers catch up.
int gci_shard_dispatch_enqueue(gci_ctx* ctx) {
  gci_ctx::Shard& sh = ctx->shard;
  XStep x;
  GCI_TRY(step_args(ctx, &x));
  const XLayout& lay = x.lay;
  ctx->stage_begin(GCI_ST_XDISPATCH);
  xchg_begin_kernel<<<1, 32, 0, ctx->stream>>>(x.mine, x.send_cnt, sh.world);
  GCI_LAUNCH_CHECK(ctx);
  int f = 0;
  for (size_t i = 0; i < ctx->n_files; i++) {
    FileTable& ft = ctx->files[i];
    if (ft.kind != 0) continue;
    BamFile& b = ctx->bam[ft.src];
    if (b.n) {
      dispatch1_kernel<<<(unsigned)((b.n + 255) / 256), 256, 0, ctx->stream>>>(
          lay, x.peers, sh.rank, f, b.n, ctx->n_reads, b.read_id.as<uint32_t>(), ft.win.as<long long>(),
          b.ref_id.as<int32_t>(), b.ref_start.as<int32_t>(), b.ref_end.as<int32_t>(), b.qlen.as<int32_t>(),
          ctx->highq.as<uint8_t>());
      GCI_LAUNCH_CHECK(ctx);
    }
    f++;
  }
  xchg_signal_kernel<<<1, 32, 0, ctx->stream>>>(lay, x.peers, sh.rank, 0, x.send_cnt);
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  return GCI_OK;
}

// The rest of the step, after gci_shard_dispatch_enqueue and the PAF legs (home tables): wait for the winners, join at
// the homes, dispatch 2, consume.  Leaves the survivors in ctx->surv_* (shard.surv_slots slots), events counted.
int gci_shard_exchange_enqueue(gci_ctx* ctx, double op, int32_t track, int32_t flank_len) {
  gci_ctx::Shard& sh = ctx->shard;
  XStep x;
  GCI_TRY(step_args(ctx, &x));
  const XLayout& lay = x.lay;
  const XPeers& peers = x.peers;
  char* mine = x.mine;
  uint32_t* send_cnt = x.send_cnt;
  unsigned long long* d_err = x.d_err;
  ctx->counted_track = -1;
  HomeArgs ha;
  memset(&ha, 0, sizeof ha);
  ha.n_files = (int)ctx->n_files;
  int f = 0;
  for (size_t i = 0; i < ctx->n_files; i++) {
    FileTable& ft = ctx->files[i];
    HomeFile& hf = ha.f[i];
    if (ft.kind != 0) {
      // a table indexed by home read: the PAF election over home-local ids (or a caller's table with such ids)
      hf.bam = -1;
      hf.win = ft.win.as<long long>();
      hf.c = ft.ref_id.as<int32_t>(); hf.s = ft.start.as<int32_t>(); hf.e = ft.end.as<int32_t>(); hf.q = ft.qlen.as<int32_t>();
      continue;
    }
    hf.bam = f++;
  }
  ctx->stage_begin(GCI_ST_XWAIT);
  xchg_wait_kernel<<<1, 32, 0, ctx->stream>>>(lay, mine, 0, d_err);
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  // home: join all files of the home reads, survivors to the owners of their contigs
  const size_t slots = (size_t)std::max<int64_t>(1, sh.surv_slots);
  GCI_TRY(ctx->ensure(ctx->surv_contig, 4 * slots));
  GCI_TRY(ctx->ensure(ctx->surv_start, 4 * slots));
  GCI_TRY(ctx->ensure(ctx->surv_end, 4 * slots));
  ctx->stage_begin(GCI_ST_JOIN);
  if (sh.n_home) {
    home_join_kernel<<<(sh.n_home + 255) / 256, 256, 0, ctx->stream>>>(
        lay, peers, sh.rank, ha, sh.n_home, sh.hq_home.as<uint8_t>(), op, sh.d_owner.as<int32_t>(), ctx->n_contigs,
        send_cnt, d_err + 2, d_err);
    GCI_LAUNCH_CHECK(ctx);
  }
  xchg_signal_kernel<<<1, 32, 0, ctx->stream>>>(lay, peers, sh.rank, 1, send_cnt);
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  ctx->stage_begin(GCI_ST_XWAIT);
  xchg_wait_kernel<<<1, 32, 0, ctx->stream>>>(lay, mine, 1, d_err);
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  // owner: the survivors that arrived, their depth events counted on the way (the count arrays are zeroed right
  // here, so that they are still in L2 when the REDs arrive)
  BucketArgs bk;
  memset(&bk, 0, sizeof bk);
  if (track >= 0) GCI_TRY(gci_depth_prepare(ctx, track, flank_len, &bk));
  ctx->stage_begin(GCI_ST_XCONSUME);
  {
    // one CTA per SM x 8 unless the slot capacity asks for more to keep a CTA's range within CONSUME_ROWS_PER_CTA
    const long long by_block = (sh.surv_slots + 255) / 256;
    const long long by_cap = (sh.surv_slots + CONSUME_ROWS_PER_CTA - 1) / CONSUME_ROWS_PER_CTA;
    const long long grid = std::max<long long>(1, std::max(by_cap, std::min<long long>(by_block, (long long)ctx->sm_count * 8)));
    consume2_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(
        lay, mine, ctx->surv_contig.as<int32_t>(), ctx->surv_start.as<int32_t>(), ctx->surv_end.as<int32_t>(),
        send_cnt + GCI_MAX_RANKS, bk);
  }
  GCI_LAUNCH_CHECK(ctx);
  ctx->stage_end();
  if (track >= 0) {
    ctx->counted_track = track;
    ctx->counted_flank = flank_len;
  }
  return GCI_OK;
}

void gci_shard_destroy_internal(gci_ctx* ctx) {
  gci_ctx::Shard& sh = ctx->shard;
  for (int r = 0; r < GCI_MAX_RANKS; r++) {
    if (sh.peer[r] && sh.mapped[r]) cudaIpcCloseMemHandle(sh.peer[r]);
    sh.peer[r] = nullptr;
    sh.mapped[r] = false;
  }
  cudaGetLastError();
  for (DevBuf* d : {&sh.d_owner, &sh.area, &sh.send_cnt, &sh.hq_home}) ctx->release(*d);
  for (auto& d : sh.hwin) ctx->release(d);
  sh.on = sh.opened = false;
}

extern "C" {

int gci_shard_home(uint32_t read_id, int32_t world, int32_t* rank, uint32_t* local) {
  if (world < 1 || world > GCI_MAX_RANKS) return GCI_E_ARG;
  if (rank) *rank = (int32_t)home_rank(read_id, (uint32_t)world);
  if (local) *local = home_local(read_id, (uint32_t)world);
  return GCI_OK;
}

// contig owners and this rank's place among `world` ranks; gate_selected[n_contigs] (NULL = the contigs given to
// gci_set_contigs) = every contig selected by --chrs on ANY rank: the gates and the PAF election must see them all.
// Call after gci_set_contigs (whose `selected` then means: selected AND owned by this rank).
int gci_shard_config(gci_ctx* ctx, int32_t rank, int32_t world, const int32_t* contig_owner, const uint8_t* gate_selected) {
  if (!ctx || world < 1 || world > GCI_MAX_RANKS || rank < 0 || rank >= world || !contig_owner) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  if (ctx->n_contigs <= 0) return ctx->fail(GCI_E_ARG, "gci_shard_config before gci_set_contigs");
  ctx->epoch++;
  gci_ctx::Shard& sh = ctx->shard;
  for (int c = 0; c < ctx->n_contigs; c++) {
    if (contig_owner[c] < 0 || contig_owner[c] >= world) return ctx->fail(GCI_E_ARG, "contig %d has owner %d", c, contig_owner[c]);
    if (ctx->selected[c] && contig_owner[c] != rank)
      return ctx->fail(GCI_E_ARG, "contig %d is selected here but owned by rank %d", c, contig_owner[c]);
  }
  sh.rank = rank;
  sh.world = world;
  sh.owner.assign(contig_owner, contig_owner + ctx->n_contigs);
  GCI_TRY(gci_h2d(ctx, sh.d_owner, contig_owner, 4 * (size_t)ctx->n_contigs));
  if (gate_selected) GCI_TRY(gci_h2d(ctx, ctx->d_gate_sel, gate_selected, (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  sh.on = true;
  sh.n_home = home_count(ctx->n_reads, (uint32_t)rank, (uint32_t)world);
  return GCI_OK;
}

// this rank's exchange area for read sets of up to max_reads reads and max_bam_files BAM files per read type; `out`
// (optional) receives its CUDA IPC handle for the other processes
int gci_shard_alloc(gci_ctx* ctx, uint32_t max_reads, int32_t max_bam_files, gci_ipc_handle* out) {
  if (!ctx || max_bam_files < 1 || max_bam_files > GCI_MAX_FILES) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  gci_ctx::Shard& sh = ctx->shard;
  if (!sh.on) return ctx->fail(GCI_E_ARG, "gci_shard_alloc before gci_shard_config");
  ctx->epoch++;
  sh.opened = false;
  sh.max_files = max_bam_files;
  sh.cap1 = sh.cap2 = (int64_t)home_count(max_reads, 0, (uint32_t)sh.world);     // rank 0 is home to the most
  sh.surv_slots = sh.cap2 * sh.world;
  const XLayout lay = make_layout(ctx);
  sh.area_bytes = (size_t)lay.total();
  GCI_TRY(ctx->ensure(sh.area, sh.area_bytes));
  // epoch, flags AND rows: a row is valid when its tag equals the epoch, so recycled memory must not hold old rows
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(sh.area.p, 0, sh.area_bytes, ctx->stream));
  GCI_TRY(ctx->ensure(sh.hq_home, (size_t)sh.cap1 + 16));
  GCI_TRY(ctx->ensure(sh.send_cnt, 4 * (GCI_MAX_RANKS + 4)));      // [world] send cursors | survivors received
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (out) {
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(gci_ipc_handle), "IPC handle does not fit");
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    GCI_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, sh.area.p));
    memcpy(out, &h, sizeof h);
  }
  return GCI_OK;
}

void* gci_shard_area(gci_ctx* ctx) { return ctx ? ctx->shard.area.p : nullptr; }

// map the peers' areas: handles[r] = what rank r got from gci_shard_alloc (same max_reads / max_bam_files everywhere)
int gci_shard_open(gci_ctx* ctx, const gci_ipc_handle* handles) {
  if (!ctx || !handles) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  gci_ctx::Shard& sh = ctx->shard;
  if (!sh.on || !sh.area.p) return ctx->fail(GCI_E_ARG, "gci_shard_open before gci_shard_alloc");
  ctx->epoch++;
  for (int r = 0; r < sh.world; r++) {
    if (r == sh.rank) { sh.peer[r] = sh.area.p; sh.mapped[r] = false; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, &handles[r], sizeof h);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return ctx->fail(GCI_E_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    sh.peer[r] = p;
    sh.mapped[r] = true;
  }
  sh.opened = true;
  return GCI_OK;
}

int gci_shard_close(gci_ctx* ctx) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  gci_ctx::Shard& sh = ctx->shard;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->epoch++;
  for (int r = 0; r < GCI_MAX_RANKS; r++) {
    if (sh.peer[r] && sh.mapped[r]) cudaIpcCloseMemHandle(sh.peer[r]);
    sh.peer[r] = nullptr;
    sh.mapped[r] = false;
  }
  cudaGetLastError();
  sh.opened = false;
  return GCI_OK;
}

// contexts of ONE process (several "ranks" on the GPUs this process can address; tests): areas[r] = gci_shard_area
// of rank r's context
int gci_shard_attach(gci_ctx* ctx, void* const* areas) {
  if (!ctx || !areas) return GCI_E_ARG;
  gci_ctx::Shard& sh = ctx->shard;
  if (!sh.on || !sh.area.p) return ctx->fail(GCI_E_ARG, "gci_shard_attach before gci_shard_alloc");
  ctx->epoch++;
  cudaSetDevice(ctx->device);
  for (int r = 0; r < sh.world; r++) {
    if (!areas[r]) return ctx->fail(GCI_E_ARG, "gci_shard_attach: no area for rank %d", r);
    sh.peer[r] = r == sh.rank ? sh.area.p : areas[r];
    sh.mapped[r] = false;
    // an area on another GPU of this process: its pointer is valid here (UVA) once peer access is on
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, areas[r]) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != ctx->device) {
      const cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return ctx->fail(GCI_E_CUDA, "no peer access from GPU %d to GPU %d: %s", ctx->device, at.device, cudaGetErrorString(e));
      }
    }
    cudaGetLastError();
  }
  sh.opened = true;
  return GCI_OK;
}

}  // extern "C"
