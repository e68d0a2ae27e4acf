// Multi-GPU exchange of the genome-row terms over NCCL (NVLink / NVSwitch), on the context's own stream.
//
// Contigs shard across GPUs; what the ranks have to share for the final `.gci` genome row and the mean
// depth is tiny: (sum depth, sum length, curated-contig count, curated lengths) per rank (GCI.py:572-587,
// :862-868).  The score kernels leave those terms on the device; one pack kernel builds a fixed-size row,
// one ncclAllGather moves it, one device->host copy returns every rank's row together with this rank's own
// per-contig terms: a single synchronisation per step.  NCCL is resolved at run time with dlopen so that the
// process uses the NCCL that is already loaded (torch's) and the library has no link-time dependency on it.
#include <dlfcn.h>

#include <algorithm>

#include "common.cuh"

namespace {
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, gci_nccl_id, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
  if (g_nccl.ok) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy torch already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
  g_nccl.h = h;
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, gci_nccl_id, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
    err = "libnccl lacks an expected symbol";
    return false;
  }
  g_nccl.ok = true;
  return true;
}
constexpr int NCCL_INT64 = 4;   // ncclInt64 (nccl.h)
}  // namespace

// row = [sum depth, sum length, curated contigs, n lengths, lengths ... (cap)] from the score result buffer
// res = [n50 (no+1) | n_ctg (no) | depth sum (no) | gap slots (n_slots)]; the order of lengths is irrelevant
// for an N50, so they are appended with an atomic cursor
__global__ void pack_row_kernel(const int64_t* __restrict__ res, int64_t no, int64_t n_slots_bound,
                                const int64_t* __restrict__ owner_off, long long sum_len, int64_t cap,
                                long long* __restrict__ row, int64_t iv_cap, unsigned long long* __restrict__ epoch) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0 && epoch) *epoch += 1;                        // one exchange per launch, overflow or not: ranks stay in step
  if (owner_off[no] > iv_cap) return;                      // scan overflow: the host redoes this step
  const int64_t n_slots = min(n_slots_bound, owner_off[no] + no);
  if (i == 0) {
    long long sd = 0, sc = 0;
    for (int64_t o = 0; o < no; o++) { sc += res[no + 1 + o]; sd += res[2 * no + 1 + o]; }
    row[0] = sd;
    row[1] = sum_len;
    row[2] = sc;
  }
  if (i >= n_slots) return;
  // non-positive slots were not emitted by the reference (or, for an owner without intervals, are a
  // non-positive [E - S] that can never be an N50): only positive lengths travel
  const long long v = res[3 * no + 1 + i];
  if (v > 0) {
    const unsigned long long k = atomicAdd((unsigned long long*)&row[3], 1ull);
    if ((int64_t)k < cap) row[4 + k] = v;
  }
}

// ---- all-gather of the rows over NVLink peer memory ------------------------------------------------
// The exchange is 16 KB per rank and pure latency, so instead of a library collective every rank STORES its row
// straight into the receive area of every peer (mapped with CUDA IPC at gci_comm_p2p_open; NVSwitch gives every
// pair a direct path) and then waits for the peers' rows to land in its own area.  One kernel, block b talks to
// rank b: push my row to b, fence, raise my arrival flag at b with the step's epoch; then spin on b's flag in my
// own memory and copy b's row to the fixed gather buffer the device->host copy reads.  Two receive areas
// alternate by epoch parity: a rank can be at most one step ahead of a peer (it cannot finish step e+1 before
// the peer has pushed step e+1, which the peer only does after it is done with step e), so the area of step e
// is never overwritten while it is still being read.  The epoch lives in device memory and is advanced by the
// pack kernel, so the whole thing replays inside a CUDA graph.  A wait that lasts longer than
// P2P_TIMEOUT_NS reports a negative length count for that rank instead of hanging the GPU (the host then
// returns an error and goes back to NCCL).
struct PeerAreas { long long* area[GCI_MAX_RANKS]; };
constexpr unsigned long long P2P_TIMEOUT_NS = 4000000000ull;

__device__ __forceinline__ unsigned long long* p2p_flags(long long* area, int world, int64_t row_cap) {
  return reinterpret_cast<unsigned long long*>(area + 2 * (int64_t)world * row_cap);
}

__global__ void __launch_bounds__(256)
row_exchange_kernel(const long long* __restrict__ row, int64_t row_n, PeerAreas peers, int my_rank, int world,
                    int64_t row_cap, long long* __restrict__ all /* [world][row_n] */) {
  const int b = blockIdx.x;                                   // the rank this block talks to
  long long* mine = peers.area[my_rank];
  const unsigned long long e = p2p_flags(mine, world, row_cap)[2 * world];   // epoch of this step (pack kernel)
  const int par = (int)(e & 1ull);
  // push: only the used part of the row travels (4 header words + the lengths)
  const int64_t used = min((long long)row_n, 4ll + max(0ll, row[3]));
  long long* dst = peers.area[b] + ((int64_t)par * world + my_rank) * row_cap;
  for (int64_t i = threadIdx.x; i < used; i += blockDim.x) dst[i] = row[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned long long* f = p2p_flags(peers.area[b], world, row_cap) + par * world + my_rank;
    *f = e;
  }
  // wait for rank b's row of this step in my own memory
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    volatile unsigned long long* f = p2p_flags(mine, world, row_cap) + par * world + b;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int ok = 1;
    while (*f < e) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > P2P_TIMEOUT_NS) { ok = 0; break; }
      __nanosleep(200);
    }
    __threadfence_system();
    s_ok = ok;
  }
  __syncthreads();
  if (!s_ok) {                                                // the host sees a negative length count for rank b
    if (threadIdx.x == 0) all[(int64_t)b * row_n + 3] = -1;
    return;
  }
  const volatile long long* src = mine + ((int64_t)par * world + b) * row_cap;
  long long* out = all + (int64_t)b * row_n;
  const int64_t got = min((long long)row_n, 4ll + max(0ll, (long long)src[3]));
  for (int64_t i = threadIdx.x; i < row_n; i += blockDim.x) out[i] = i < got ? src[i] : 0ll;
}

// pack + all-gather + copy of every rank's row into `h_rows` (pinned), all enqueued on the context's stream
int gci_enqueue_genome_row(gci_ctx* ctx, Track& t, int64_t no, int64_t sum_len, int64_t cap, int64_t* h_rows) {
  const int64_t row_n = 4 + cap;
  const int world = ctx->comm_world;
  const int64_t n_slots_bound = t.iv_cap + no;
  DevBuf &d_res = ctx->tmp[1], &d_row = ctx->tmp[2], &d_all = ctx->tmp[5];
  GCI_TRY(ctx->ensure(d_row, 8 * (size_t)row_n));
  GCI_TRY(ctx->ensure(d_all, 8 * (size_t)row_n * world));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(d_row.p, 0, 8 * (size_t)row_n, ctx->stream));
  unsigned long long* d_epoch = nullptr;
  if (ctx->p2p_ok && row_n <= ctx->p2p_row_cap)
    d_epoch = reinterpret_cast<unsigned long long*>(ctx->p2p_buf.as<long long>() + 2 * (int64_t)world * ctx->p2p_row_cap) +
              2 * world;
  pack_row_kernel<<<(unsigned)((std::max<int64_t>(1, n_slots_bound) + 255) / 256), 256, 0, ctx->stream>>>(
      d_res.as<int64_t>(), no, n_slots_bound, t.owner_off.as<int64_t>(), (long long)sum_len, cap,
      d_row.as<long long>(), t.iv_cap, d_epoch);
  GCI_LAUNCH_CHECK(ctx);
  if (d_epoch) {
    PeerAreas pa;
    for (int r = 0; r < GCI_MAX_RANKS; r++) pa.area[r] = r < world ? (long long*)ctx->p2p_peer[r] : nullptr;
    row_exchange_kernel<<<world, 256, 0, ctx->stream>>>(d_row.as<long long>(), row_n, pa, ctx->comm_rank, world,
                                                        ctx->p2p_row_cap, d_all.as<long long>());
    GCI_LAUNCH_CHECK(ctx);
    GCI_TRY(gci_d2h(ctx, h_rows, d_all.p, 8 * (size_t)row_n * world));
    return GCI_OK;
  }
  const int rc = g_nccl.AllGather(d_row.p, d_all.p, (size_t)row_n, NCCL_INT64, ctx->nccl_comm, ctx->stream);
  if (rc != 0)
    return ctx->fail(GCI_E_CUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->launches++;
  GCI_TRY(gci_d2h(ctx, h_rows, d_all.p, 8 * (size_t)row_n * world));
  return GCI_OK;
}

extern "C" {

int gci_comm_unique_id(gci_nccl_id* out) {
  std::string err;
  if (!out || !load_nccl(err)) return GCI_E_CUDA;
  return g_nccl.GetUniqueId(out) == 0 ? GCI_OK : GCI_E_CUDA;
}

int gci_comm_init(gci_ctx* ctx, const gci_nccl_id* id, int32_t rank, int32_t world) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  std::string err;
  if (!load_nccl(err)) return ctx->fail(GCI_E_CUDA, "%s", err.c_str());
  if (ctx->nccl_comm) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  const int rc = g_nccl.CommInitRank(&ctx->nccl_comm, world, *id, rank);
  if (rc != 0)
    return ctx->fail(GCI_E_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return GCI_OK;
}

void gci_comm_destroy_internal(gci_ctx* ctx) {
  if (!ctx) return;
  for (int r = 0; r < GCI_MAX_RANKS; r++) {
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_buf.p) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_peer[r] = nullptr;
  }
  ctx->p2p_ok = false;
  cudaGetLastError();
  if (ctx->nccl_comm && g_nccl.ok) g_nccl.CommDestroy(ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

// allocate this rank's receive area for rows of up to 4 + cap words and hand out its CUDA IPC handle
int gci_comm_p2p_alloc(gci_ctx* ctx, int64_t cap, gci_ipc_handle* out) {
  if (!ctx || !out || cap < 1) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(gci_ipc_handle), "IPC handle does not fit");
  const int world = ctx->comm_world;
  if (world < 1 || world > GCI_MAX_RANKS) return ctx->fail(GCI_E_ARG, "peer exchange supports up to %d ranks", GCI_MAX_RANKS);
  ctx->p2p_ok = false;
  ctx->p2p_row_cap = 4 + cap;
  const size_t bytes = 8 * (size_t)(2 * world * ctx->p2p_row_cap + 2 * world + 1);
  GCI_TRY(ctx->ensure(ctx->p2p_buf, bytes));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->p2p_buf.p, 0, bytes, ctx->stream));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  memset(out, 0, sizeof *out);
  cudaIpcMemHandle_t h;
  GCI_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, ctx->p2p_buf.p));
  memcpy(out, &h, sizeof h);
  return GCI_OK;
}

int gci_comm_p2p_disable(gci_ctx* ctx) {
  if (!ctx) return GCI_E_ARG;
  ctx->epoch++;
  ctx->p2p_ok = false;
  return GCI_OK;
}

// map every peer's receive area (handles[r] = what rank r got from gci_comm_p2p_alloc with the same cap)
int gci_comm_p2p_open(gci_ctx* ctx, const gci_ipc_handle* handles) {
  if (!ctx || !handles) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  if (!ctx->p2p_buf.p || ctx->p2p_row_cap <= 0) return ctx->fail(GCI_E_ARG, "gci_comm_p2p_alloc has not been called");
  const int world = ctx->comm_world;
  for (int r = 0; r < world; r++) {
    if (r == ctx->comm_rank) { ctx->p2p_peer[r] = ctx->p2p_buf.p; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, &handles[r], sizeof h);
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return ctx->fail(GCI_E_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s (the NCCL exchange stays in use)", r,
                       cudaGetErrorString(e));
    }
    ctx->p2p_peer[r] = p;
  }
  ctx->p2p_ok = true;
  return GCI_OK;
}

int gci_genome_row(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len, int64_t sum_len, int64_t cap,
                   int64_t* n50, int64_t* n_ctg, int64_t* depth_sums, int64_t* rows) {
  if (!ctx || !rows || cap < 1 || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  if (!ctx->nccl_comm) return ctx->fail(GCI_E_ARG, "gci_genome_row: gci_comm_init has not been called");
  Track& t = ctx->track[track];
  if (t.owners_are_windows || !t.sums_valid)
    return ctx->fail(GCI_E_ARG, "gci_genome_row needs a whole-contig scan of a track with valid depth sums");
  int64_t no = 0, n_slots = 0;
  ctx->stage_begin(GCI_ST_SCORE);
  GCI_TRY(gci_launch_score_kernels(ctx, t, dist_percent, flank_len, &no, &n_slots, false));
  const int64_t row_n = 4 + cap;
  const int world = ctx->comm_world;
  const int64_t n_own = 3 * no + 1;
  int64_t* h = (int64_t*)ctx->pinned(8 * (size_t)(n_own + row_n * world));
  if (!h) return ctx->fail(GCI_E_NOMEM, "pinned scratch allocation failed");
  GCI_TRY(gci_enqueue_genome_row(ctx, t, no, sum_len, cap, h + n_own));
  GCI_TRY(gci_d2h(ctx, h, ctx->tmp[1].p, 8 * (size_t)n_own));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (n50) memcpy(n50, h, 8 * (size_t)(no + 1));
  long long all_c = 0, all_d = 0;
  for (int64_t o = 0; o < no; o++) {
    if (n_ctg) n_ctg[o] = h[no + 1 + o];
    if (depth_sums) depth_sums[o] = h[2 * no + 1 + o];
    all_c += h[no + 1 + o];
    all_d += h[2 * no + 1 + o];
  }
  if (n_ctg) n_ctg[no] = all_c;
  if (depth_sums) depth_sums[no] = all_d;
  memcpy(rows, h + n_own, 8 * (size_t)row_n * world);
  for (int r = 0; r < world; r++)
    if (rows[(size_t)r * row_n + 3] < 0) {
      ctx->p2p_ok = false;
      return ctx->fail(GCI_E_CUDA, "genome row: rank %d did not deliver its row over peer memory in time", r);
    }
  for (int r = 0; r < world; r++)
    if (rows[(size_t)r * row_n + 3] > cap)
      return ctx->fail(GCI_E_ARG, "gci_genome_row: rank %d has %lld curated lengths, more than cap %lld", r,
                       (long long)rows[(size_t)r * row_n + 3], (long long)cap);
  return GCI_OK;
}

}  // extern "C"
