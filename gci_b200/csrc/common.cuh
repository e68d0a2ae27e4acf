// Shared host-side state and small device helpers of libgci_cuda.so (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gci_cuda.h"

// ---- geometry --------------------------------------------------------------------------------
// The reference axis is cut into tiles of GCI_TILE positions; a tile never straddles two contigs
// and every contig owns floor(L / TILE) + 1 tiles, so at least one padded position follows the
// last base (the "-1" event of a read ending at L and the end of a run touching L - fl land there).
// One warp expands one tile.  Padded positions hold depth 0 (every +1 has met its -1 by position L);
// their flag bits are unspecified and never read (every consumer masks with the contig length).
constexpr int GCI_TILE = 1024;            // positions per depth tile = per warp (4 KB of int32)
constexpr int GCI_TILE_THREADS = 256;     // 8 warps = 8 independent tiles in flight per CTA
constexpr int GCI_CHUNK = 8192;           // positions per CTA in the streaming kernels (max / flags / sum)
constexpr int GCI_RUN_CHUNK_WORDS = 2048; // flag words (of 32 positions) per run-extraction chunk
constexpr int GCI_MAX_RANKS = 16;         // GPUs of one NVLink domain taking part in the peer-memory row exchange

#define GCI_CUDA_TRY(ctx, expr)                                                              \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      (ctx)->fail(GCI_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),       \
                  __FILE__, __LINE__);                                                       \
      return GCI_E_CUDA;                                                                     \
    }                                                                                        \
  } while (0)

#define GCI_TRY(expr)            \
  do {                           \
    int r__ = (expr);            \
    if (r__ != GCI_OK) return r__; \
  } while (0)

struct gci_ctx;

// growth-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct BamFile {
  int64_t n = 0, n_ops = 0;
  DevBuf ref_id, ref_start, mapq, flag, nm, qlen, read_id, cigar_off, cigar;
  DevBuf stats;    // uint32[n][8]: Mx(M+=+X), I, D, N, S, pad..   (K1 output)
  DevBuf ref_end;  // int32[n]                                      (K2 output)
  DevBuf tile_rec; // int2[n_op_tiles]: first / last record of every CIGAR op tile (built at upload)
  DevBuf dense_list; // int32[n_op_tiles] + counter: tiles of the staged CIGAR kernel (the others stream)
  int64_t n_dense = 0; // number of entries of dense_list (-1: the count is still on its way to the host)
  DevBuf span_list;  // int32[n]: records gated from their HBM row (no ops / cut by a tile border / streaming tile)
  int64_t n_span = 0;
};

// PAF lines of one file, resident on the device (columns 0,1,2,3,5,7,8,9,10,11 of GCI.py:218-229)
struct PafFile {
  int64_t n = 0;
  DevBuf read_id, qlen, qstart, qend, ref_id, tstart, tend, nmatch, alnlen, mapq;
  DevBuf rows;   // int4[n][2], packed at upload: (ref, qlen, qstart, qend) (tstart, tend, nmatch, alnlen) — the election
                 // gathers whole lines, and a line is then ONE 32-byte sector instead of eight
};

// one file after its per-file leg: at most one entry per read
struct FileTable {
  int kind = 0;            // 0 = BAM (entries are records of bam[src]), 1 = table (own columns)
  int src = -1;
  int paf = -1;            // >= 0: the table is (re)built from PAF file paf[...] by every gci_filter
  int64_t n = 0;
  DevBuf ref_id, start, end, qlen;   // for kind==1; for BAM these alias the BamFile columns
  DevBuf win;              // int64[n_reads]: winning entry key per read (-1 = absent)
};

struct Track {
  bool allocated = false;
  DevBuf depth;            // int32[total_padded]
  DevBuf flags;            // uint32[total_padded / 32]   bit = (lo < depth <= hi)
  bool flags_valid = false;
  int32_t flags_lo = 0, flags_hi = 0;
  DevBuf sums;             // int64[n_contigs]  sum of depth per contig
  bool sums_valid = false; // kept up to date by gci_depth / gci_merge_max / gci_mask_gaps
  // last scan
  int64_t n_intervals = 0, n_owners = 0;
  bool owners_are_windows = false;
  DevBuf iv_start, iv_end; // int32[n_intervals]
  DevBuf owner_off;        // int64[n_owners + 1]
  DevBuf win_contig, win_lo, win_hi;   // int32 / int64 / int64 [n_owners]: scan windows
  std::vector<int64_t> raw_lo, raw_hi; // windows as given by the caller (before slice normalisation)
  std::vector<int64_t> h_owner_off;    // host copy of owner_off after the last scan / load
  std::vector<int32_t> owner_contig;   // contig of every owner
  int64_t iv_cap = 0;                  // capacity of iv_start / iv_end
  int32_t scan_flank = 0;
};

struct StageTimer {
  std::vector<cudaEvent_t> pool;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t next = 0;
};

struct gci_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t h2d_done = nullptr;
  std::string err;
  int64_t launches = 0;
  int64_t dev_bytes = 0;
  int sm_count = 148;
  int depth_ctas_per_sm = 0;        // resident CTAs of depth_tile_kernel per SM (occupancy query)

  // contigs
  int32_t n_contigs = 0;
  std::vector<int64_t> len;         // host copies
  std::vector<uint8_t> selected;
  std::vector<int64_t> tile_off;    // [n+1] first tile of each contig
  std::vector<int64_t> pos_off;     // [n+1] = tile_off * TILE
  int64_t n_tiles = 0, total_padded = 0;
  DevBuf d_len, d_selected, d_tile_off, d_owner_of;
  bool owner_of_stale = true;

  // N runs (sorted by (contig,start) on upload)
  int64_t n_nruns = 0;
  DevBuf d_nr_contig, d_nr_start, d_nr_end;   // int32, int64, int64

  // read set
  uint32_t n_reads = 0;
  // bam / files are pools: gci_reads_begin resets the used counts and keeps the device buffers, so a
  // steady stream of read sets does not churn cudaMalloc / cudaFree
  std::vector<BamFile> bam;
  size_t n_bam = 0;
  std::vector<PafFile> paf;
  size_t n_paf = 0;
  DevBuf paf_keep;                  // uint8 per PAF line (all files of the read set): passed the per-line gate
  DevBuf d_name_rank;
  std::vector<FileTable> files;     // join order
  size_t n_files = 0;
  DevBuf highq;                     // uint8[n_reads]  high-quality set of the current gci_filter run
  DevBuf highq_base;                // uint8[n_reads]  marks that came with uploaded tables (kept across runs)
  DevBuf surv_contig, surv_start, surv_end;   // int32[n_reads]; contig < 0 = not a survivor
  bool filtered = false;
  // the join kernel already counted the depth events of (track, flank) into the tile table (gci_pipeline)
  int32_t counted_track = -1, counted_flank = 0;
  int64_t n_survivors = 0;
  std::vector<int32_t> name_rank;   // host copy for the PAF election tie-break

  // depth scratch
  DevBuf tile_cnt, events, scan_tmp, scan_tmp2, misc, d_err;
  DevBuf chunk_cnt, chunk_off, run_stage;
  DevBuf scan_lvl[8];               // block sums / offsets of the recursive scan, two per level
  DevBuf tmp[10];                   // small per-call scratch (score terms, fetches)
  Track track[GCI_MAX_TRACKS];

  // gci_depth_gzip: packed members of the last size query, kept until they are fetched; the encoder's own scratch
  // (gzip.cu: CRC / power tables, range descriptors, run-start bitmask, run list, member sizes and offsets)
  bool gz_valid = false;
  unsigned long long gz_key = 0;
  int64_t gz_total = 0;
  DevBuf sw_csum, sw_out;           // plot feed (gci_sliding_window): depth prefix sums of the region, packed points
  DevBuf gz_tables, gz_seg, gz_hdr, gz_bits, gz_tile_cnt, gz_tile_run, gz_run_pos, gz_run_val, gz_msize, gz_moff, gz_packed;

  // NCCL communicator of a multi-GPU run (comm.cu; resolved with dlopen)
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  // genome row over NVLink peer memory (comm.cu): every rank's receive area is mapped into all the others
  DevBuf p2p_buf;                     // [2 parities][world][p2p_row_cap] int64 + arrival flags [2][world] u64 + epoch u64
  void* p2p_peer[GCI_MAX_RANKS] = {}; // receive areas as mapped in this process (own rank: p2p_buf.p)
  int64_t p2p_row_cap = 0;            // int64 words per row the area was sized for
  bool p2p_ok = false;                // all peers mapped: gci_enqueue_genome_row pushes instead of calling NCCL

  // read-set sharding over several GPUs (shard.cu): contigs have an owner rank, reads a home rank (home_rank())
  struct Shard {
    bool on = false;
    int rank = 0, world = 1, max_files = 0;
    uint32_t n_home = 0;              // reads this rank is home to (home_count())
    int64_t cap1 = 0, cap2 = 0;       // rows per (file, source rank) / per source rank in the inboxes
    int64_t surv_slots = 0;           // survivor slots behind surv_contig / surv_start / surv_end (world * cap2)
    std::vector<int32_t> owner;       // host copy of the contig owners
    DevBuf d_owner;                   // int32[n_contigs]
    DevBuf area;                      // this rank's exchange area (header + inboxes), shared with the peers
    size_t area_bytes = 0;
    void* peer[GCI_MAX_RANKS] = {};   // every rank's area as mapped here (own rank: area.p)
    bool mapped[GCI_MAX_RANKS] = {};  // peer[r] came from cudaIpcOpenMemHandle
    bool opened = false;
    DevBuf send_cnt;                  // u32[max_files * world + world]: rows sent per (file, destination) / destination
    DevBuf hq_home;                   // u8[n_home]: high-quality marks of the home reads
    DevBuf hwin[GCI_MAX_FILES];       // per BAM upload: int64[n_home] winning inbox row per home read
  } shard;
  DevBuf d_gate_sel;                  // uint8[n_contigs]: contigs the gates accept (== selected unless sharded: the
                                      // PAF election must see every selected contig, not only the owned ones)

  // CUDA-graph replay of gci_pipeline / gci_pipeline_row (filter.cu).  With the stage timers off, the second
  // call with an unchanged signature (arguments, read-set sizes, device buffers, `epoch`) is captured into a
  // graph and every further one replays it: one cudaGraphLaunch instead of ~40 stream operations.
  bool timing = true;               // stage timers on: eager launches, one event pair per stage
  bool graph_ok = true;             // GCI_GRAPH=0 in the environment turns the replay off
  bool capturing = false, capture_abort = false;
  uint64_t alloc_gen = 0;           // bumped by every device / pinned (re)allocation or release
  uint64_t epoch = 0;               // bumped by every entry point that changes state the graph depends on
  std::vector<char> pipe_sig, pipe_sig_seen, pipe_sig_bad;
  cudaGraphExec_t pipe_exec = nullptr;
  int64_t pipe_launches = 0;        // kernels inside the captured graph
  int64_t pipe_no = 0, pipe_slots = 0;
  int64_t graph_replays = 0;        // steps that ran as a graph launch

  std::vector<int64_t> lay_cache;   // last run-chunk layout uploaded to chunk_off
  std::vector<char> ob_cache;       // last OwnerBounds uploaded to tmp[0]
  void* pipe_pin = nullptr;         // persistent pinned block of gci_pipeline
  size_t pipe_pin_cap = 0;

  StageTimer timer;
  void* pinned_scratch = nullptr;
  size_t pinned_cap = 0;

  void drop_graph();
  int fail(int code, const char* fmt, ...);
  int ensure(DevBuf& b, size_t bytes);            // grow-only
  void release(DevBuf& b);
  void* pinned(size_t bytes);
  void stage_begin(int stage);
  void stage_end();
};

// ---- host helpers implemented in api.cu ---------------------------------------------------------
int gci_h2d(gci_ctx* ctx, DevBuf& dst, const void* src, size_t bytes);
int gci_d2h(gci_ctx* ctx, void* dst, const void* src, size_t bytes);

// device-wide exclusive scans (scan.cu)
int gci_exclusive_scan_i64_from_i32(gci_ctx* ctx, const int32_t* in, int64_t* out, int64_t n, int64_t* total_dev);
int gci_exclusive_scan_i32(gci_ctx* ctx, const int32_t* in, int32_t* out, int64_t n);

// stage entry points
int gci_index_bam(gci_ctx* ctx, BamFile& b);
int gci_pack_paf_rows(gci_ctx* ctx, PafFile& p);
void gci_index_bam_finish(gci_ctx* ctx, BamFile& b);
int gci_run_bam_leg(gci_ctx* ctx, int file_idx, int bam_idx, int32_t mq, int32_t mq_cutoff, double ip, double cp);
int gci_run_paf_legs(gci_ctx* ctx, int32_t mq, int32_t mq_cutoff, double ip);
int gci_run_join(gci_ctx* ctx, double op);
int gci_alloc_track(gci_ctx* ctx, int track);
int gci_depth_enqueue(gci_ctx* ctx, int32_t track, int32_t flank_len, int32_t lo, int32_t hi);   // depth.cu
int gci_depth_prepare(gci_ctx* ctx, int32_t track, int32_t flank_len, struct BucketArgs* bk);        // depth.cu
int gci_run_join_counting(gci_ctx* ctx, double op, int32_t track, int32_t flank_len);                // filter.cu
int gci_compute_flags(gci_ctx* ctx, int track, int32_t lo, int32_t hi);
int gci_launch_score_kernels(gci_ctx* ctx, Track& t, double dist_percent, int32_t flank_len, int64_t* n_owners,
                             int64_t* n_slots, bool pending);   // scan.cu: result stays in ctx->tmp[1], no sync
int gci_enqueue_genome_row(gci_ctx* ctx, Track& t, int64_t no, int64_t sum_len, int64_t cap, int64_t* h_rows);   // comm.cu
int gci_shard_dispatch_enqueue(gci_ctx* ctx);                                                              // shard.cu
int gci_shard_exchange_enqueue(gci_ctx* ctx, double op, int32_t track, int32_t flank_len);           // shard.cu
void gci_shard_destroy_internal(gci_ctx* ctx);                                                      // shard.cu
int gci_scan_enqueue(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* pin);
int gci_scan_finish(gci_ctx* ctx, int32_t track, const int64_t* pin, bool* overflow);
extern "C" void gci_comm_destroy_internal(gci_ctx* ctx);

#define GCI_LAUNCH_CHECK(ctx)                      \
  do {                                             \
    (ctx)->launches++;                             \
    GCI_CUDA_TRY(ctx, cudaGetLastError());         \
  } while (0)

// ---- device helpers ----------------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// last index i in [0, n) with a[i] <= key  (a ascending, a[0] <= key assumed)
template <typename T>
__device__ __forceinline__ int64_t upper_bound_minus1(const T* __restrict__ a, int64_t n, T key) {
  int64_t lo = 0, hi = n;   // invariant: a[lo] <= key, a[hi] > key (virtual)
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid; else hi = mid;
  }
  return lo;
}

// Python slice index normalisation (PySlice_AdjustIndices, step 1)
__device__ __forceinline__ long long py_slice_index(long long i, long long len) {
  if (i < 0) {
    i += len;
    if (i < 0) i = 0;
  } else if (i >= len) {
    i = len;
  }
  return i;
}

// ---- survivor -> depth events (shared by the join kernel and the depth stage) -----------------------
// Every survivor contributes +1 at a = start+fl and -1 at b = end-fl+1, both normalised like a Python slice
// (GCI.py:304-306); the events are bucketed per depth tile (depth.cu).
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


struct BucketArgs {
  int32_t fl;
  const int64_t* len;
  const int64_t* tile_off;
  uint32_t* cnt_start;     // per tile: +1 events (reads starting in the tile); NULL: no counting
  uint32_t* cnt_end;       // per tile: -1 events
  long long* sums;         // per-contig depth sums
  int32_t n_contigs;
};

// The contig table (length, first tile) is looked up once per survivor; from global memory that is two more DRAM
// round trips at the end of an already long chain of dependent gathers.  Kernels that walk millions of survivors
// keep it in shared memory when it fits (assemblies have tens to hundreds of contigs).  Block-collective.
constexpr int GCI_SMEM_CONTIGS = 1024;
struct ContigCache {
  int64_t len[GCI_SMEM_CONTIGS];
  int64_t tile_off[GCI_SMEM_CONTIGS + 1];
};
__device__ __forceinline__ void contig_cache_load(ContigCache& cc, BucketArgs& bk) {
  if (!bk.cnt_start || bk.n_contigs > GCI_SMEM_CONTIGS) return;
  for (int i = threadIdx.x; i < bk.n_contigs; i += blockDim.x) cc.len[i] = bk.len[i];
  for (int i = threadIdx.x; i <= bk.n_contigs; i += blockDim.x) cc.tile_off[i] = bk.tile_off[i];
  __syncthreads();
  bk.len = cc.len;
  bk.tile_off = cc.tile_off;
}

// Home rank of a read (shard.cu): block-cyclic over the ranks in blocks of GCI_HOME_BLOCK ids.  Consecutive records
// of a coordinate-sorted file carry neighbouring read ids, so the lanes of a warp send their rows to ONE home, into
// adjacent slots: the stores leave over NVLink as a few hundred contiguous bytes instead of 24-byte pieces to every
// peer in turn (id % world measured 0.68 ms per step for the dispatch at 8 GPUs against 0.14 ms of local stores).
#define GCI_HOME_BLOCK 64u
__host__ __device__ __forceinline__ uint32_t home_rank(uint32_t q, uint32_t world) { return (q / GCI_HOME_BLOCK) % world; }
__host__ __device__ __forceinline__ uint32_t home_local(uint32_t q, uint32_t world) {
  return q / (GCI_HOME_BLOCK * world) * GCI_HOME_BLOCK + q % GCI_HOME_BLOCK;
}
__host__ __device__ __forceinline__ uint32_t home_global(uint32_t h, uint32_t rank, uint32_t world) {
  return (h / GCI_HOME_BLOCK * world + rank) * GCI_HOME_BLOCK + h % GCI_HOME_BLOCK;
}
// how many of the ids 0 .. n_reads-1 are homed at `rank` (their local ids are 0 .. count-1)
__host__ __device__ __forceinline__ uint32_t home_count(uint32_t n_reads, uint32_t rank, uint32_t world) {
  const uint32_t cyc = GCI_HOME_BLOCK * world, full = n_reads / cyc, rest = n_reads % cyc;
  const uint32_t lo = rank * GCI_HOME_BLOCK;
  const uint32_t tail = rest > lo ? (rest - lo < GCI_HOME_BLOCK ? rest - lo : GCI_HOME_BLOCK) : 0u;
  return full * GCI_HOME_BLOCK + tail;
}

// Work split of the persistent kernels: every CTA walks ONE contiguous range of the items, `step` items per round.
// (A grid-stride loop makes every warp jump by grid x block items per round: each round touches a new page of every
// array for every warp of the GPU — measured 3-4x slower on the 6 M-read walks of the join / survivor kernels,
// profiles/r02j — and neighbouring rounds of a warp stop sharing a contig, which the per-warp partial sums rely on.)
__device__ __forceinline__ void cta_range(long long total, long long step, long long& begin, long long& end) {
  long long per = (total + gridDim.x - 1) / gridDim.x;
  per = (per + step - 1) / step * step;
  begin = min(total, (long long)blockIdx.x * per);
  end = min(total, begin + per);
}

// The two events of one survivor (c < 0: none) are counted per tile with one 32-bit atomic each; its slice length
// goes to the contig's depth sum and `have` to the survivor count.  Those two land on a handful of hot addresses, so
// they are accumulated per WARP in registers across the iterations of a grid-stride loop (consecutive reads sit on
// the same contig) and flushed when the contig changes and at the end: a launch over millions of reads sends a few
// thousand atomics to them instead of one per warp and iteration.  Warp-collective: every lane calls add / flush.
struct WarpSums {
  int32_t cur_c;
  long long cur_sum;
  int n_have;
  __device__ __forceinline__ void init() { cur_c = -1; cur_sum = 0; n_have = 0; }
  __device__ __forceinline__ void add(const BucketArgs& bk, int32_t c, int32_t s, int32_t e, bool have) {
    long long covered = 0;
    if (bk.cnt_start && c >= 0) {
      const Slice sl = survivor_slice(c, s, e, bk.fl, bk.len, bk.tile_off);
      if (sl.ok) {
        atomicAdd(&bk.cnt_start[sl.tile_a], 1u);       // one 32-bit RED per event into a dense array: several times
        atomicAdd(&bk.cnt_end[sl.tile_b], 1u);         // the rate of a 64-bit RED into a 16-byte strided table
        covered = sl.b - sl.a;
      } else {
        c = -1;
      }
    } else {
      c = -1;
    }
    n_have += __popc(__ballot_sync(0xffffffffu, have));
    const unsigned act = __ballot_sync(0xffffffffu, c >= 0);
    if (act == 0) return;
    const int32_t c0 = __shfl_sync(0xffffffffu, c, __ffs(act) - 1);
    if (__all_sync(0xffffffffu, c < 0 || c == c0)) {
      const long long t = warp_sum_ll(covered);
      if (c0 != cur_c) {
        flush_sum(bk);
        cur_c = c0;
      }
      cur_sum += t;
    } else if (c >= 0) {                      // a warp straddling two contigs: rare, straight to memory
      atomicAdd((unsigned long long*)(bk.sums + c), (unsigned long long)covered);
    }
  }
  __device__ __forceinline__ void flush_sum(const BucketArgs& bk) {
    if ((threadIdx.x & 31) == 0 && cur_c >= 0 && cur_sum)
      atomicAdd((unsigned long long*)(bk.sums + cur_c), (unsigned long long)cur_sum);
    cur_sum = 0;
  }
  __device__ __forceinline__ void flush(const BucketArgs& bk, unsigned long long* count) {
    flush_sum(bk);
    if ((threadIdx.x & 31) == 0 && count && n_have) atomicAdd(count, (unsigned long long)n_have);
    n_have = 0;
  }
};

// (double)a / (double)b <= lim and >= lim exactly as fp64 div.rn decides them (Python int / int is correctly rounded,
// GCI.py:165), without paying for an fp64 division on almost every record: a single-precision quotient carries a
// relative error below 2^-21 here (two conversions + one division, each within 2^-23), so unless it lies within
// 2^-18 of the limit the comparison is already decided; only the few records on the boundary divide in fp64.
// |a|, b < 2^40; b > 0.
__device__ __forceinline__ bool ratio_le(long long a, long long b, double lim) {
  const float q = __fdividef((float)a, (float)b);
  const float l = (float)lim;
  const float tol = fabsf(l) * 3.9e-6f + 1e-30f;
  if (q < l - tol) return true;
  if (q > l + tol) return false;
  return (double)a / (double)b <= lim;
}
__device__ __forceinline__ bool ratio_ge(long long a, long long b, double lim) {
  const float q = __fdividef((float)a, (float)b);
  const float l = (float)lim;
  const float tol = fabsf(l) * 3.9e-6f + 1e-30f;
  if (q > l + tol) return true;
  if (q < l - tol) return false;
  return (double)a / (double)b >= lim;
}


// The join of ONE read from the entries of its files (GCI.py:272-301); k[f] < 0 = absent in file f.
struct JoinEntry { int32_t c, s, e, q; };
template <class Entry>
__device__ __forceinline__ bool join_one(int n_files, const long long* k, const Entry& entry, bool hq, double op,
                                         uint32_t r, unsigned long long* err, int32_t& c, int32_t& s, int32_t& e) {
  bool have = false;
  c = -1; s = 0; e = 0;
  if (n_files == 1) {                                           // :300-301
    if (k[0] >= 0) {
      const JoinEntry x = entry(0);
      have = true;
      c = x.c; s = x.s; e = x.e;
    }
    return have;
  }
  bool comm = true;
  for (int f = 0; f < n_files; f++) comm = comm && (k[f] >= 0);                            // :274-277
  if (k[0] >= 0 && (hq || comm)) {                                                         // :279-280
    const JoinEntry x = entry(0);
    have = true;
    c = x.c; s = x.s; e = x.e;
  }
  for (int f = 1; f < n_files; f++) {                                                       // :281-299
    if (k[f] < 0) continue;
    const JoinEntry x = entry(f);
    if (have) {
      if (x.c == c) {
        const long long ov = (long long)min(x.e, e) - (long long)max(x.s, s);
        if (x.q == 0) {                                                                     // ZeroDivisionError :292
          atomicOr(err, 8ull);
          atomicMin(err + 1, (unsigned long long)r);
          have = false;
        } else if (!ratio_ge(ov, x.q, op)) {                                                // ovlp / qlen < -op: delete
          have = false;
        } else {
          s = max(x.s, s);
          e = min(x.e, e);
        }
      } else {
        have = false;
      }
    } else if (hq) {
      have = true;
      c = x.c; s = x.s; e = x.e;
    }
  }
  return have;
}

