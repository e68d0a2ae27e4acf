// libgci_cuda.so — context, memory, uploads and fetches (the C ABI of include/gci_cuda.h).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>

#include "common.cuh"

// ---- ctx helpers --------------------------------------------------------------------------------
int gci_ctx::fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  err = buf;
  return code;
}

void gci_ctx::drop_graph() {
  if (pipe_exec) cudaGraphExecDestroy(pipe_exec);
  pipe_exec = nullptr;
  pipe_sig.clear();
}

int gci_ctx::ensure(DevBuf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.cap >= bytes) return GCI_OK;
  alloc_gen++;
  if (capturing) capture_abort = true;   // a buffer moved while its old address may already sit in the graph
  if (b.p) {
    cudaFree(b.p);
    dev_bytes -= (int64_t)b.cap;
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = (bytes + 255) & ~size_t(255);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    cudaGetLastError();
    return fail(GCI_E_NOMEM, "cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
  }
  b.cap = want;
  dev_bytes += (int64_t)want;
  return GCI_OK;
}

void gci_ctx::release(DevBuf& b) {
  if (b.p) {
    alloc_gen++;
    cudaFree(b.p);
    dev_bytes -= (int64_t)b.cap;
  }
  b.p = nullptr;
  b.cap = 0;
}

void* gci_ctx::pinned(size_t bytes) {
  if (pinned_cap >= bytes) return pinned_scratch;
  alloc_gen++;
  if (capturing) capture_abort = true;
  if (pinned_scratch) cudaFreeHost(pinned_scratch);
  pinned_scratch = nullptr;
  pinned_cap = 0;
  size_t want = std::max<size_t>(bytes, 1 << 20);
  if (cudaHostAlloc(&pinned_scratch, want, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    pinned_scratch = nullptr;
    return nullptr;
  }
  pinned_cap = want;
  return pinned_scratch;
}

void gci_ctx::stage_begin(int stage) {
  if (!timing) return;
  StageTimer& t = timer;
  if (t.next + 2 > t.pool.size()) {
    for (int i = 0; i < 64; i++) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      t.pool.push_back(e);
    }
  }
  StageTimer::Span s{stage, t.pool[t.next], t.pool[t.next + 1]};
  t.next += 2;
  cudaEventRecord(s.a, stream);
  t.spans.push_back(s);
}

void gci_ctx::stage_end() {
  if (!timing) return;
  if (!timer.spans.empty()) cudaEventRecord(timer.spans.back().b, stream);
}

int gci_h2d(gci_ctx* ctx, DevBuf& dst, const void* src, size_t bytes) {
  // a host->device copy (of a possibly short-lived host buffer) must never become a graph node
  if (ctx->capturing) ctx->capture_abort = true;
  GCI_TRY(ctx->ensure(dst, bytes));
  if (bytes) GCI_CUDA_TRY(ctx, cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return GCI_OK;
}

int gci_d2h(gci_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (bytes) GCI_CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return GCI_OK;
}

static void free_bam(gci_ctx* ctx, BamFile& b) {
  for (DevBuf* d : {&b.ref_id, &b.ref_start, &b.mapq, &b.flag, &b.nm, &b.qlen, &b.read_id, &b.cigar_off,
                    &b.cigar, &b.stats, &b.ref_end, &b.tile_rec, &b.dense_list, &b.span_list})
    ctx->release(*d);
}

static void free_table(gci_ctx* ctx, FileTable& f) {
  for (DevBuf* d : {&f.ref_id, &f.start, &f.end, &f.qlen, &f.win}) ctx->release(*d);
}

static void free_track(gci_ctx* ctx, Track& t) {
  for (DevBuf* d : {&t.depth, &t.flags, &t.sums, &t.iv_start, &t.iv_end, &t.owner_off, &t.win_contig,
                    &t.win_lo, &t.win_hi})
    ctx->release(*d);
  t.allocated = false;
  t.flags_valid = false;
  t.sums_valid = false;
  t.n_intervals = t.n_owners = 0;
  t.owners_are_windows = false;
  t.raw_lo.clear();
  t.raw_hi.clear();
  t.h_owner_off.clear();
  t.owner_contig.clear();
  t.iv_cap = 0;
}

int gci_alloc_track(gci_ctx* ctx, int track) {
  if (track < 0 || track >= GCI_MAX_TRACKS) return ctx->fail(GCI_E_ARG, "bad track %d", track);
  if (ctx->n_contigs <= 0) return ctx->fail(GCI_E_ARG, "gci_set_contigs has not been called");
  Track& t = ctx->track[track];
  if (t.allocated) return GCI_OK;
  GCI_TRY(ctx->ensure(t.depth, sizeof(int32_t) * (size_t)ctx->total_padded));
  GCI_TRY(ctx->ensure(t.flags, sizeof(uint32_t) * (size_t)(ctx->total_padded / 32)));
  GCI_TRY(ctx->ensure(t.sums, sizeof(int64_t) * (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.depth.p, 0, sizeof(int32_t) * (size_t)ctx->total_padded, ctx->stream));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(t.sums.p, 0, sizeof(int64_t) * (size_t)ctx->n_contigs, ctx->stream));
  t.allocated = true;
  t.flags_valid = false;
  t.sums_valid = false;
  return GCI_OK;
}

// ---- C ABI ------------------------------------------------------------------------------------------
extern "C" {

int gci_version(void) { return 100; }

int gci_create(int device, gci_ctx** out) {
  if (!out) return GCI_E_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return GCI_E_CUDA;   // no device: the library has no CPU fallback
  }
  gci_ctx* ctx = new gci_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return GCI_E_CUDA;
  }
  ctx->stream = ctx->own_stream;
  {
    const char* g = getenv("GCI_GRAPH");
    ctx->graph_ok = !(g && g[0] == '0');
  }
  cudaEventCreateWithFlags(&ctx->h2d_done, cudaEventDisableTiming);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = ctx;
  return GCI_OK;
}

void gci_destroy(gci_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->drop_graph();
  gci_comm_destroy_internal(ctx);
  gci_shard_destroy_internal(ctx);
  ctx->release(ctx->d_gate_sel);
  for (auto& b : ctx->bam) free_bam(ctx, b);
  for (auto& f : ctx->files) free_table(ctx, f);
  for (auto& p : ctx->paf)
    for (DevBuf* d : {&p.read_id, &p.qlen, &p.qstart, &p.qend, &p.ref_id, &p.tstart, &p.tend, &p.nmatch, &p.alnlen, &p.mapq, &p.rows})
      ctx->release(*d);
  ctx->release(ctx->paf_keep);
  ctx->release(ctx->d_name_rank);
  for (auto& t : ctx->track) free_track(ctx, t);
  for (DevBuf* d : {&ctx->d_len, &ctx->d_selected, &ctx->d_tile_off, &ctx->d_owner_of, &ctx->d_nr_contig, &ctx->d_nr_start,
                    &ctx->d_nr_end, &ctx->highq, &ctx->highq_base, &ctx->surv_contig, &ctx->surv_start, &ctx->surv_end,
                    &ctx->tile_cnt, &ctx->events,
                    &ctx->scan_tmp, &ctx->scan_tmp2, &ctx->misc, &ctx->d_err, &ctx->chunk_cnt, &ctx->chunk_off, &ctx->run_stage})
    ctx->release(*d);
  for (DevBuf* d : {&ctx->gz_tables, &ctx->gz_seg, &ctx->gz_hdr, &ctx->gz_bits, &ctx->gz_tile_cnt, &ctx->gz_tile_run,
                    &ctx->gz_run_pos, &ctx->gz_run_val, &ctx->gz_msize, &ctx->gz_moff, &ctx->gz_packed})
    ctx->release(*d);
  ctx->release(ctx->sw_csum);
  ctx->release(ctx->sw_out);
  for (auto& d : ctx->scan_lvl) ctx->release(d);
  for (auto& d : ctx->tmp) ctx->release(d);
  for (cudaEvent_t e : ctx->timer.pool) cudaEventDestroy(e);
  if (ctx->pinned_scratch) cudaFreeHost(ctx->pinned_scratch);
  if (ctx->pipe_pin) cudaFreeHost(ctx->pipe_pin);
  if (ctx->h2d_done) cudaEventDestroy(ctx->h2d_done);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* gci_last_error(gci_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gci_set_stream(gci_ctx* ctx, void* s) {
  if (!ctx) return GCI_E_ARG;
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  ctx->epoch++;
  return GCI_OK;
}

int gci_set_timing(gci_ctx* ctx, int32_t on) {
  if (!ctx) return GCI_E_ARG;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->timing = on != 0;
  ctx->timer.spans.clear();
  ctx->timer.next = 0;
  return GCI_OK;
}

int gci_sync(gci_ctx* ctx) {
  if (!ctx) return GCI_E_ARG;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

void* gci_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void gci_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int gci_stage_reset(gci_ctx* ctx) {
  if (!ctx) return GCI_E_ARG;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->timer.spans.clear();
  ctx->timer.next = 0;
  return GCI_OK;
}

int gci_stage_ms(gci_ctx* ctx, int stage, double* ms, int64_t* launches) {
  if (!ctx || !ms) return GCI_E_ARG;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  double total = 0;
  int64_t k = 0;
  for (auto& s : ctx->timer.spans) {
    if (s.stage != stage) continue;
    float f = 0;
    if (cudaEventElapsedTime(&f, s.a, s.b) == cudaSuccess) {
      total += f;
      k++;
    } else {
      cudaGetLastError();
    }
  }
  *ms = total;
  if (launches) *launches = k;
  return GCI_OK;
}

int64_t gci_kernel_launches(gci_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t gci_device_bytes(gci_ctx* ctx) { return ctx ? ctx->dev_bytes : 0; }
int64_t gci_graph_replays(gci_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }

// ---- contigs ------------------------------------------------------------------------------------
int gci_set_contigs(gci_ctx* ctx, int32_t n, const int64_t* lengths, const uint8_t* selected) {
  if (!ctx || n < 0 || (n && !lengths)) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  for (auto& t : ctx->track) free_track(ctx, t);
  ctx->n_contigs = n;
  ctx->len.assign(lengths, lengths + n);
  ctx->selected.assign(n, 1);
  if (selected) ctx->selected.assign(selected, selected + n);
  ctx->tile_off.assign(n + 1, 0);
  ctx->pos_off.assign(n + 1, 0);
  for (int i = 0; i < n; i++) {
    if (lengths[i] < 0 || lengths[i] >= (int64_t(1) << 31) - GCI_TILE)
      return ctx->fail(GCI_E_ARG, "contig %d length %lld out of range", i, (long long)lengths[i]);
    int64_t tiles = ctx->selected[i] ? lengths[i] / GCI_TILE + 1 : 0;
    ctx->tile_off[i + 1] = ctx->tile_off[i] + tiles;
    ctx->pos_off[i + 1] = ctx->tile_off[i + 1] * GCI_TILE;
  }
  ctx->n_tiles = ctx->tile_off[n];
  ctx->total_padded = ctx->n_tiles * GCI_TILE;
  GCI_TRY(gci_h2d(ctx, ctx->d_len, ctx->len.data(), sizeof(int64_t) * n));
  GCI_TRY(gci_h2d(ctx, ctx->d_selected, ctx->selected.data(), n));
  GCI_TRY(gci_h2d(ctx, ctx->d_gate_sel, ctx->selected.data(), n));      // gci_shard_config may widen it
  ctx->shard.on = false;                                                // sharding is configured per contig table
  GCI_TRY(gci_h2d(ctx, ctx->d_tile_off, ctx->tile_off.data(), sizeof(int64_t) * (n + 1)));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n_nruns = 0;
  ctx->name_rank.clear();
  ctx->owner_of_stale = true;
  ctx->lay_cache.clear();
  ctx->ob_cache.clear();
  return GCI_OK;
}

int gci_set_name_rank(gci_ctx* ctx, const int32_t* name_rank) {
  if (!ctx || !name_rank) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  ctx->name_rank.assign(name_rank, name_rank + ctx->n_contigs);
  GCI_TRY(gci_h2d(ctx, ctx->d_name_rank, name_rank, sizeof(int32_t) * (size_t)ctx->n_contigs));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_set_n_runs(gci_ctx* ctx, int64_t n, const int32_t* contig, const int64_t* start, const int64_t* end) {
  if (!ctx || n < 0 || (n && (!contig || !start || !end))) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  // keep only runs on selected contigs, normalise like a Python slice assignment, sort by (contig,start)
  std::vector<int64_t> idx;
  std::vector<int64_t> s2(n), e2(n);
  for (int64_t i = 0; i < n; i++) {
    int32_t c = contig[i];
    if (c < 0 || c >= ctx->n_contigs || !ctx->selected[c]) continue;
    int64_t L = ctx->len[c];
    auto norm = [L](int64_t v) {
      if (v < 0) { v += L; if (v < 0) v = 0; } else if (v >= L) v = L;
      return v;
    };
    s2[i] = norm(start[i]);
    e2[i] = norm(end[i]);
    if (s2[i] < e2[i]) idx.push_back(i);
  }
  std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
    return contig[a] != contig[b] ? contig[a] < contig[b] : s2[a] < s2[b];
  });
  std::vector<int32_t> c3;
  std::vector<int64_t> s3, e3;
  for (int64_t i : idx) {
    c3.push_back(contig[i]);
    s3.push_back(s2[i]);
    e3.push_back(e2[i]);
  }
  ctx->n_nruns = (int64_t)idx.size();
  GCI_TRY(gci_h2d(ctx, ctx->d_nr_contig, c3.data(), sizeof(int32_t) * c3.size()));
  GCI_TRY(gci_h2d(ctx, ctx->d_nr_start, s3.data(), sizeof(int64_t) * s3.size()));
  GCI_TRY(gci_h2d(ctx, ctx->d_nr_end, e3.data(), sizeof(int64_t) * e3.size()));
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

// ---- read set -----------------------------------------------------------------------------------
int gci_reads_begin(gci_ctx* ctx, uint32_t n_reads) {
  if (!ctx) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n_bam = 0;                 // device buffers of the pools are kept and reused
  ctx->n_files = 0;
  ctx->n_paf = 0;
  ctx->n_reads = n_reads;
  if (ctx->shard.on) {
    gci_ctx::Shard& sh = ctx->shard;
    sh.n_home = home_count(n_reads, (uint32_t)sh.rank, (uint32_t)sh.world);
    if (sh.area.p && (int64_t)sh.n_home > sh.cap1)
      return ctx->fail(GCI_E_ARG, "gci_reads_begin: %u reads exceed the exchange area (gci_shard_alloc)", n_reads);
  }
  ctx->filtered = false;
  ctx->n_survivors = 0;
  GCI_TRY(ctx->ensure(ctx->highq, n_reads));
  GCI_TRY(ctx->ensure(ctx->highq_base, n_reads));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->highq.p, 0, n_reads ? n_reads : 1, ctx->stream));
  GCI_CUDA_TRY(ctx, cudaMemsetAsync(ctx->highq_base.p, 0, n_reads ? n_reads : 1, ctx->stream));
  return GCI_OK;
}

int gci_upload_bam(gci_ctx* ctx, int64_t n, const int32_t* ref_id, const int32_t* ref_start,
                   const uint8_t* mapq, const uint16_t* flag, const int32_t* nm, const int32_t* qlen,
                   const uint32_t* read_id, const uint64_t* cigar_off, const uint32_t* cigar, int64_t n_ops) {
  if (!ctx || n < 0 || n_ops < 0 || !cigar_off) return GCI_E_ARG;
  if (n && (!ref_id || !ref_start || !mapq || !flag || !nm || !qlen || !read_id)) return GCI_E_ARG;
  if (n_ops && !cigar) return GCI_E_ARG;
  if (n >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "more than 2^31 records in one file");
  if (cigar_off[0] != 0 || cigar_off[n] != (uint64_t)n_ops)
    return ctx->fail(GCI_E_ARG, "cigar_off must start at 0 and end at n_ops");
  if ((int)ctx->n_files >= GCI_MAX_FILES) return ctx->fail(GCI_E_ARG, "too many files");
  cudaSetDevice(ctx->device);
  if (ctx->n_bam == ctx->bam.size()) ctx->bam.emplace_back();
  BamFile& b = ctx->bam[ctx->n_bam++];
  b.n = n;
  b.n_ops = n_ops;
  ctx->stage_begin(GCI_ST_H2D);
  GCI_TRY(gci_h2d(ctx, b.ref_id, ref_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, b.ref_start, ref_start, 4 * n));
  GCI_TRY(gci_h2d(ctx, b.mapq, mapq, n));
  GCI_TRY(gci_h2d(ctx, b.flag, flag, 2 * n));
  GCI_TRY(gci_h2d(ctx, b.nm, nm, 4 * n));
  GCI_TRY(gci_h2d(ctx, b.qlen, qlen, 4 * n));
  GCI_TRY(gci_h2d(ctx, b.read_id, read_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, b.cigar_off, cigar_off, 8 * (n + 1)));
  GCI_TRY(gci_h2d(ctx, b.cigar, cigar, 4 * n_ops));
  ctx->stage_end();
  GCI_TRY(gci_index_bam(ctx, b));   // op-tile -> record index: a property of the file, built once
  if (ctx->n_files == ctx->files.size()) ctx->files.emplace_back();
  FileTable& f = ctx->files[ctx->n_files++];
  f.kind = 0;
  f.src = (int)ctx->n_bam - 1;
  f.paf = -1;
  f.n = n;
  ctx->filtered = false;
  // the caller may reuse its host buffers as soon as we return
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  gci_index_bam_finish(ctx, b);
  return GCI_OK;
}

int gci_upload_paf(gci_ctx* ctx, int64_t n, const uint32_t* read_id, const int32_t* qlen, const int32_t* qstart,
                   const int32_t* qend, const int32_t* ref_id, const int32_t* tstart, const int32_t* tend,
                   const int32_t* nmatch, const int32_t* alnlen, const int32_t* mapq) {
  if (!ctx || n < 0) return GCI_E_ARG;
  if (n && (!read_id || !qlen || !qstart || !qend || !ref_id || !tstart || !tend || !nmatch || !alnlen || !mapq))
    return GCI_E_ARG;
  if ((int)ctx->n_files >= GCI_MAX_FILES) return ctx->fail(GCI_E_ARG, "too many files");
  if (n >= (int64_t(1) << 31)) return ctx->fail(GCI_E_ARG, "more than 2^31 lines in one PAF file");
  cudaSetDevice(ctx->device);
  if (ctx->n_paf == ctx->paf.size()) ctx->paf.emplace_back();
  PafFile& p = ctx->paf[ctx->n_paf++];
  p.n = n;
  ctx->stage_begin(GCI_ST_H2D);
  GCI_TRY(gci_h2d(ctx, p.read_id, read_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.qlen, qlen, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.qstart, qstart, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.qend, qend, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.ref_id, ref_id, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.tstart, tstart, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.tend, tend, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.nmatch, nmatch, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.alnlen, alnlen, 4 * n));
  GCI_TRY(gci_h2d(ctx, p.mapq, mapq, 4 * n));
  ctx->stage_end();
  GCI_TRY(gci_pack_paf_rows(ctx, p));   // a property of the file, built once (filter.cu)
  if (ctx->n_files == ctx->files.size()) ctx->files.emplace_back();
  FileTable& f = ctx->files[ctx->n_files++];
  f.kind = 1;   // a table, rebuilt from the PAF lines by every gci_filter (the gates are filter arguments)
  f.src = -1;
  f.paf = (int)ctx->n_paf - 1;
  f.n = 0;
  ctx->filtered = false;
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_upload_table(gci_ctx* ctx, int64_t n, const uint32_t* read_id, const int32_t* ref_id, const int32_t* start,
                     const int32_t* end, const int32_t* qlen, const uint8_t* highq);   // filter.cu

// ---- depth load / fetch ------------------------------------------------------------------------------
int gci_load_depth(gci_ctx* ctx, int32_t track, int32_t contig, const int32_t* depth, int64_t n) {
  if (!ctx || !depth) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  ctx->epoch++;
  GCI_TRY(gci_alloc_track(ctx, track));
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || n != ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_load_depth: contig %d / length %lld mismatch", contig, (long long)n);
  Track& t = ctx->track[track];
  ctx->stage_begin(GCI_ST_H2D);
  GCI_CUDA_TRY(ctx, cudaMemcpyAsync(t.depth.as<int32_t>() + ctx->pos_off[contig], depth, sizeof(int32_t) * n,
                                    cudaMemcpyHostToDevice, ctx->stream));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  t.flags_valid = false;
  t.sums_valid = false;   // recomputed lazily by gci_depth_sums
  return GCI_OK;
}

int gci_fetch_depth(gci_ctx* ctx, int32_t track, int32_t contig, int32_t* out, int64_t n) {
  if (!ctx || !out || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (!t.allocated) return ctx->fail(GCI_E_ARG, "track %d holds no depth", track);
  if (contig < 0 || contig >= ctx->n_contigs || !ctx->selected[contig] || n != ctx->len[contig])
    return ctx->fail(GCI_E_ARG, "gci_fetch_depth: contig %d / length %lld mismatch", contig, (long long)n);
  ctx->stage_begin(GCI_ST_D2H);
  GCI_TRY(gci_d2h(ctx, out, t.depth.as<int32_t>() + ctx->pos_off[contig], sizeof(int32_t) * n));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

int gci_fetch_intervals(gci_ctx* ctx, int32_t track, int64_t cap, int32_t* start, int32_t* end, int64_t* owner_off,
                        int64_t* n) {
  if (!ctx || track < 0 || track >= GCI_MAX_TRACKS) return GCI_E_ARG;
  cudaSetDevice(ctx->device);
  Track& t = ctx->track[track];
  if (n) *n = t.n_intervals;
  if (!start && !end && !owner_off) return GCI_OK;
  if (cap < t.n_intervals) return ctx->fail(GCI_E_ARG, "interval buffer too small (%lld < %lld)", (long long)cap,
                                            (long long)t.n_intervals);
  ctx->stage_begin(GCI_ST_D2H);
  if (start) GCI_TRY(gci_d2h(ctx, start, t.iv_start.p, sizeof(int32_t) * t.n_intervals));
  if (end) GCI_TRY(gci_d2h(ctx, end, t.iv_end.p, sizeof(int32_t) * t.n_intervals));
  if (owner_off && (int64_t)t.h_owner_off.size() == t.n_owners + 1)
    memcpy(owner_off, t.h_owner_off.data(), sizeof(int64_t) * (size_t)(t.n_owners + 1));
  ctx->stage_end();
  GCI_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return GCI_OK;
}

}  // extern "C"
