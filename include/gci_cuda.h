/*
 * libgci_cuda.so — C ABI of the B200-native GCI hot path
 * (alignment filter -> per-base depth -> gap scan -> score terms).
 *
 * The reference (yeeus/GCI @ 455e19c7) is a single Python script with no FFI; the
 * boundary this library sits behind is the set of Python functions of GCI.py that
 * make up the hot path.  Each entry point below names the reference lines it
 * replaces.  The Python host layer (gci_b200/pipeline.py) mirrors the reference's
 * function signatures and calls these through ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (GCI_E_*); the message is
 *     available from gci_last_error(ctx);
 *   - the caller owns every host buffer, the library owns every device buffer
 *     until gci_destroy; one host thread per context; all work is queued on one
 *     CUDA stream per context and host-visible results are synchronised before
 *     the call returns unless the name says `_async`;
 *   - coordinates are 0-based; intervals half-open; per-contig coordinates are
 *     int32 (contigs < 2^31), genome-wide offsets int64;
 *   - there is NO CPU fallback: without a CUDA device gci_create fails.
 */
#ifndef GCI_CUDA_H
#define GCI_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gci_ctx gci_ctx;

enum {
  GCI_OK = 0,
  GCI_E_CUDA = -1,        /* a CUDA runtime call failed */
  GCI_E_ARG = -2,         /* bad argument / call order */
  GCI_E_NOMEM = -3,
  GCI_E_REFERENCE_RAISES = -4 /* the reference would raise here (ZeroDivisionError / KeyError:
                                 GCI.py:163, :165, :292) — details in gci_last_error */
};

#define GCI_MAX_TRACKS 3   /* depth tracks: 0 = HiFi, 1 = Nano, 2 = max(HiFi, Nano) */
#define GCI_MAX_FILES 16
#define GCI_NM_MISSING INT32_MIN

/* stage ids for gci_stage_ms */
enum {
  GCI_ST_H2D = 0, GCI_ST_CIGAR = 1, GCI_ST_GATE = 2, GCI_ST_JOIN = 3, GCI_ST_BUCKET = 4,
  GCI_ST_DEPTH = 5, GCI_ST_FLAGS = 6, GCI_ST_RUNS = 7, GCI_ST_MAX = 8, GCI_ST_MASK = 9,
  GCI_ST_D2H = 10, GCI_ST_PAF = 11, GCI_ST_TEXT = 12, GCI_ST_SCORE = 13,
  /* read-set exchange of a sharded run (shard.cu): dispatch to the read homes, waiting for the peers, survivors in */
  GCI_ST_XDISPATCH = 14, GCI_ST_XWAIT = 15, GCI_ST_XCONSUME = 16, GCI_ST_COUNT = 17
};

/* ---- context ------------------------------------------------------------------------ */
int gci_version(void);
int gci_create(int device, gci_ctx** out);
void gci_destroy(gci_ctx* ctx);
const char* gci_last_error(gci_ctx* ctx);
/* run on a caller-provided cudaStream_t (e.g. torch's current stream); NULL = the context's own */
int gci_set_stream(gci_ctx* ctx, void* cuda_stream);
/* Stage timers (gci_stage_ms) on / off; on by default.  With the timers off gci_pipeline and
   gci_pipeline_row record the step into a CUDA graph the second time they see the same arguments and
   read-set shape and replay it afterwards (one launch instead of ~40 stream operations); anything that
   changes the inputs of the step (contigs, uploads of another shape, another entry point touching the
   track) drops the graph.  GCI_GRAPH=0 in the environment disables the replay. */
int gci_set_timing(gci_ctx* ctx, int32_t on);
int gci_sync(gci_ctx* ctx);
/* pinned host memory so H2D/D2H copies run at PCIe speed */
void* gci_host_alloc(uint64_t bytes);
void gci_host_free(void* p);
/* CUDA-event time of one stage accumulated since the last gci_stage_reset (synchronises) */
int gci_stage_reset(gci_ctx* ctx);
int gci_stage_ms(gci_ctx* ctx, int stage, double* ms, int64_t* launches);
int64_t gci_kernel_launches(gci_ctx* ctx);       /* kernels launched by this context so far */
int64_t gci_device_bytes(gci_ctx* ctx);          /* device memory currently owned */
int64_t gci_graph_replays(gci_ctx* ctx);         /* gci_pipeline steps that ran as one CUDA graph launch */

/* ---- contig table: GCI.py:201-207 (targets_length / depths keys, --chrs selection) -------- */
/* lengths[n]; selected[n] (NULL = all): contigs named by --chrs; unselected contigs get no
   depth storage and records on them are ignored like the reference's fetch() never sees them */
int gci_set_contigs(gci_ctx* ctx, int32_t n, const int64_t* lengths, const uint8_t* selected);
/* N-runs of the assembly: GCI.py:18-46 output, consumed by gci_mask_gaps */
int gci_set_n_runs(gci_ctx* ctx, int64_t n, const int32_t* contig, const int64_t* start, const int64_t* end);

/* ---- filter stage: one read type at a time ------------------------------------------------ */
/* start a read set of `n_reads` interned read names; drops the previous set's records */
int gci_reads_begin(gci_ctx* ctx, uint32_t n_reads);
/* one BAM file decoded to columns (GCI.py:150-166 touches exactly these).  Files are joined in
   upload order; the host layer uploads PAF tables first (GCI.py:272). */
int gci_upload_bam(gci_ctx* ctx, int64_t n_records, const int32_t* ref_id, const int32_t* ref_start,
                   const uint8_t* mapq, const uint16_t* flag, const int32_t* nm, const int32_t* qlen,
                   const uint32_t* read_id, const uint64_t* cigar_off, const uint32_t* cigar,
                   int64_t n_ops);
/* one PAF file as columns (GCI.py:218-229); name_rank[n_contigs] = rank of each contig name under
   Python str ordering (tie-break of GCI.py:252), set once per context */
int gci_set_name_rank(gci_ctx* ctx, const int32_t* name_rank);
int gci_upload_paf(gci_ctx* ctx, int64_t n_lines, const uint32_t* read_id, const int32_t* qlen,
                   const int32_t* qstart, const int32_t* qend, const int32_t* ref_id,
                   const int32_t* tstart, const int32_t* tend, const int32_t* nmatch,
                   const int32_t* alnlen, const int32_t* mapq);
/* a file already reduced to one (contig,start,end,qlen) per read on the host (e.g. a PAF leg run
   elsewhere, or survivors exchanged between GPUs); highq marks reads for the high-quality set */
int gci_upload_table(gci_ctx* ctx, int64_t n, const uint32_t* read_id, const int32_t* ref_id,
                     const int32_t* start, const int32_t* end, const int32_t* qlen,
                     const uint8_t* highq);
/* read_sam gates (GCI.py:146-169), PAF election (:211-254), last-record-wins dedup (:166, :269)
   and the cross-file join (:272-301).  n_survivors = len(file1). */
int gci_filter(gci_ctx* ctx, int32_t map_qual, int32_t mq_cutoff, double iden_percent,
               double clip_percent, double ovlp_percent, int64_t* n_survivors);
/* per-record CIGAR statistics of BAM upload number `bam` (0-based, upload order) after gci_filter: what
   pysam's get_cigar_stats()[0] and reference_end give the reference (GCI.py:157-162, :166).
   stats[r*5 .. r*5+4] = bases in (M + '=' + X, I, D, N, S) — the reference only ever uses M, '=' and X
   as one sum (:164-165); ref_end[r] = htslib bam_endpos.  Either pointer may be NULL. */
int gci_fetch_cigar_stats(gci_ctx* ctx, int32_t bam, int64_t n_records, uint32_t* stats, int32_t* ref_end);
/* survivors as (read_id, contig, start, end) in read_id order; pass NULL pointers to get the count */
int gci_fetch_survivors(gci_ctx* ctx, int64_t cap, uint32_t* read_id, int32_t* contig, int32_t* start,
                        int32_t* end, int64_t* n);
/* per-file winner tables after gci_filter (for multi-GPU exchange and tests): for file f, entries
   (read_id, contig, start, end, qlen, highq) of reads present in that file */
int gci_fetch_file_table(gci_ctx* ctx, int32_t file, int64_t cap, uint32_t* read_id, int32_t* contig,
                         int32_t* start, int32_t* end, int32_t* qlen, uint8_t* highq, int64_t* n);

/* ---- depth stage -------------------------------------------------------------------------- */
/* depths[target][start+fl : end-fl+1] += 1 for every survivor (GCI.py:302-306, Python slice
   semantics) into depth track `track`; also produces the issue flags for (lo < depth <= hi) so a
   following gci_scan with the same thresholds does not re-read the depth array.  Pass
   lo = hi = INT32_MIN to skip the fused flags. */
int gci_depth(gci_ctx* ctx, int32_t track, int32_t flank_len, int32_t lo, int32_t hi);
/* depths[target][s:e] = 0 over the N-runs (GCI.py:315-329), in place */
int gci_mask_gaps(gci_ctx* ctx, int32_t track);
/* element-wise max of two tracks (GCI.py:350) */
int gci_merge_max(gci_ctx* ctx, int32_t track_a, int32_t track_b, int32_t track_out, int32_t lo, int32_t hi);
/* resume path (utility/GCI_score.py:11-39): load one contig's depth from the host */
int gci_load_depth(gci_ctx* ctx, int32_t track, int32_t contig, const int32_t* depth, int64_t n);
int gci_fetch_depth(gci_ctx* ctx, int32_t track, int32_t contig, int32_t* out, int64_t n);
/* the same values as uint8 (width 1) or uint16 (width 2) when every depth fits: 4x / 2x fewer bytes over
   PCIe.  *overflow != 0 means some value did not fit and `out` was not written: retry wider. */
int gci_fetch_depth_narrow(gci_ctx* ctx, int32_t track, int32_t contig, void* out, int64_t n, int32_t width,
                           int32_t* overflow);
/* sum of depth per contig (mean depth of GCI.py:862-868 = sum / length) */
int gci_depth_sums(gci_ctx* ctx, int32_t track, int64_t* sums /* [n_contigs] */);
/* order-independent 64-bit checksum of every contig's depth array: sum over positions p of
   (depth[p] + 1) * splitmix64(p) mod 2^64 (0 for unselected contigs).  Lets a caller compare whole tracks with
   another implementation of GCI.py:302-306 without moving them (3.1 Gbp = 12 GB) over PCIe. */
int gci_depth_hash(gci_ctx* ctx, int32_t track, uint64_t* out /* [n_contigs] */);
/* decimal text of one contig's depth, one value per line (GCI.py:115-117), produced on the GPU;
   call with out = NULL to get the byte count first */
int gci_depth_text(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count,
                   char* out, int64_t cap, int64_t* n_bytes);

/* the same text already compressed on the GPU (run-based encoder, gzip.cu): a sequence of complete gzip members (one per 8192 positions;
   the reference also writes a multi-member file, GCI.py:134-143) of `header` (e.g. ">name\n", may be empty)
   followed by the depth lines of [first, first+count).  Call with out = NULL to get the byte count. */
int gci_depth_gzip(gci_ctx* ctx, int32_t track, int32_t contig, int64_t first, int64_t count, const char* header,
                   int32_t header_len, char* out, int64_t cap, int64_t* n_bytes);

/* the whole track in one pass (what write_depth() puts on disk, GCI.py:99-143): for every selected contig c, in
   contig order, the gzip members of header c (bytes [header_off[c], header_off[c+1]) of `headers`, e.g. ">name\n")
   followed by its depth lines.  contig_bytes[n_contigs+1] (optional) = byte offset of every contig's first member.
   out == NULL: only the sizes.  A cap smaller than *n_bytes is an error (nothing is copied). */
int gci_depth_gzip_track(gci_ctx* ctx, int32_t track, const char* headers, const int64_t* header_off /* [n_contigs+1] */,
                         char* out, int64_t cap, int64_t* n_bytes, int64_t* contig_bytes);

/* ---- gap scan ----------------------------------------------------------------------------- */
/* collapse_depth_range(depths, lo, hi, flank_len, 0) over every selected contig (GCI.py:356-390) */
int gci_scan(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int32_t flank_len, int64_t* n_intervals);
/* regions variant (GCI.py:627-628): windows [start,end) on contigs, flank 0, nothing dropped */
int gci_scan_windows(gci_ctx* ctx, int32_t track, int32_t lo, int32_t hi, int64_t n_windows,
                     const int32_t* contig, const int64_t* start, const int64_t* end, int64_t* n_intervals);
/* intervals of the last scan on `track`, in contig (or window) order then position order;
   owner_off[n_owners+1] gives each contig's / window's slice */
int gci_fetch_intervals(gci_ctx* ctx, int32_t track, int64_t cap, int32_t* start, int32_t* end,
                        int64_t* owner_off, int64_t* n);

/* intervals that did not come from a scan on this context (a BED read from disk — the `--bed` entry of
   utility/GCI_score.py:586): owner o is whole contig contig[o], its intervals are
   [owner_off[o], owner_off[o+1]) of start/end, sorted and disjoint */
int gci_load_intervals(gci_ctx* ctx, int32_t track, int64_t n_owners, const int32_t* contig,
                       const int64_t* owner_off, const int32_t* start, const int32_t* end);

/* ---- plot feed: sliding_window_average_depth (GCI.py:660-705) over depth[start:end] of one contig ---------------
   Points in position order: idx = base index inside the region the point is reported at, num / den = sum and number
   of the bases averaged (the caller divides, clips at max_depth and converts to Mbp like GCI.py:679-704),
   kind = 0 full window, 1 window cut short by a zero or the end, 2 a zero-depth base (a point of its own).
   Pass NULL buffers to get the count.  A region shorter than window_size is the CALLER's case (GCI.py:674-676 sets
   the window to 1).  Uses the track's scan state (like gci_scan_windows). */
int gci_sliding_window(gci_ctx* ctx, int32_t track, int32_t contig, int64_t start, int64_t end, int64_t window_size,
                       int64_t cap, int64_t* idx, int64_t* num, int64_t* den, uint8_t* kind, int64_t* n_points);

/* ---- score terms (GCI.py:422-519) --------------------------------------------------------- */
/* per owner (contig or window) of the last scan: curated N50 of the un-merged complement, number of
   curated contigs after the -dp merge, and the complement lengths themselves (for the genome row).
   Windows use flank = window start like GCI.py:629-634. */
int gci_score_terms(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len,
                    int64_t* n50 /* [owners] */, int64_t* n_ctg /* [owners] */,
                    int64_t cap_lengths, int64_t* lengths, int64_t* lengths_off /* [owners+1] */);

/* same, plus the per-contig depth sums (and their total) in the same device->host copy: the terms of the
   genome row / mean depth that ranks exchange in a multi-GPU run.  Whole-contig scans only. */
int gci_score_terms_sums(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len, int64_t* n50,
                         int64_t* n_ctg, int64_t cap_lengths, int64_t* lengths, int64_t* lengths_off,
                         int64_t* depth_sums /* [owners+1] */);

/* ---- the whole path in one call ---------------------------------------------------------------------- */
/* gci_filter + gci_depth + gci_scan + gci_score_terms_sums on the uploaded read set with ONE host
   synchronisation (the stages only hand device buffers to each other).  Same results as the four calls. */
int gci_pipeline(gci_ctx* ctx, int32_t track, int32_t map_qual, int32_t mq_cutoff, double iden_percent,
                 double clip_percent, double ovlp_percent, int32_t flank_len, int32_t lo, int32_t hi,
                 double dist_percent, int64_t* n_survivors, int64_t* n_intervals, int64_t* n50 /* [owners+1] */,
                 int64_t* n_ctg /* [owners+1] */, int64_t* depth_sums /* [owners+1] */);

/* ---- multi-GPU: one process per GPU, contigs sharded across ranks ---------------------------------- */
/* NCCL (resolved at run time with dlopen: the copy torch.distributed already loaded, else the system one).
   Rank 0 creates the id, the host layer broadcasts its 128 bytes, every rank calls gci_comm_init. */
typedef struct { char internal[128]; } gci_nccl_id;
int gci_comm_unique_id(gci_nccl_id* out);
int gci_comm_init(gci_ctx* ctx, const gci_nccl_id* id, int32_t rank, int32_t world);
/* Optional, after gci_comm_init: exchange the genome row over NVLink peer memory instead of ncclAllGather
   (the exchange is 16 KB per rank, i.e. pure latency).  Every rank allocates a receive area for rows of up to
   4 + cap words and gets its CUDA IPC handle; the caller all-gathers the handles (any transport) and every
   rank maps its peers.  If mapping fails the NCCL exchange stays in use. */
typedef struct { char internal[64]; } gci_ipc_handle;
int gci_comm_p2p_alloc(gci_ctx* ctx, int64_t cap, gci_ipc_handle* out);
int gci_comm_p2p_open(gci_ctx* ctx, const gci_ipc_handle* handles /* [world] */);
int gci_comm_p2p_disable(gci_ctx* ctx);          /* back to ncclAllGather (every rank must do the same) */
/* gci_score_terms_sums for this rank's contigs + ONE ncclAllGather (on the context's stream) of the terms
   every rank needs for the genome row / mean depth (GCI.py:572-587, :862-868):
   rows[r * (4 + cap) ...] = { sum depth, sum length, curated contigs, n lengths, curated lengths[cap] } of rank r.
   One device->host copy, one synchronisation. */
int gci_genome_row(gci_ctx* ctx, int32_t track, double dist_percent, int32_t flank_len, int64_t sum_len, int64_t cap,
                   int64_t* n50 /* [owners+1] */, int64_t* n_ctg /* [owners+1] */, int64_t* depth_sums /* [owners+1] */,
                   int64_t* rows /* [world * (4 + cap)] */);

/* gci_pipeline + the genome row of gci_genome_row in the same single synchronisation (rows as in gci_genome_row) */
int gci_pipeline_row(gci_ctx* ctx, int32_t track, int32_t map_qual, int32_t mq_cutoff, double iden_percent,
                     double clip_percent, double ovlp_percent, int32_t flank_len, int32_t lo, int32_t hi,
                     double dist_percent, int64_t* n_survivors, int64_t* n_intervals, int64_t* n50, int64_t* n_ctg,
                     int64_t* depth_sums, int64_t sum_len, int64_t cap, int64_t* rows);

/* ---- multi-GPU: read sets sharded over the ranks (the reference's analogue is the Pool fan-out and
   dict merge of GCI.py:257-301) -------------------------------------------------------------------------------------
   Contigs have an owner rank, reads a home rank: block-cyclic in the read id (gci_shard_home: blocks of 64 ids go
   round the ranks, so neighbouring ids share a home and their rows leave as one contiguous NVLink write).  A rank uploads
     - the BAM records lying on the contigs it owns (global read ids, global contig ids), and
     - the PAF lines of the reads it is home to, with HOME-LOCAL read ids (gci_shard_home's `local`),
   and gci_pipeline / gci_pipeline_row then run: PAF election and BAM gates locally -> every per-file winner goes to
   its read's home over NVLink peer memory -> merge + join at the home (GCI.py:268-301) -> every survivor goes to the
   owner of its contig -> depth, scan and score on the owned contigs.  n_survivors is then the number of survivors
   among this rank's home reads (the global count is the sum over the ranks).
   Call order: gci_set_contigs(selected = selected AND owned) -> gci_shard_config -> gci_shard_alloc (once, sized for
   the largest read set) -> exchange the handles, gci_shard_open -> per read type gci_reads_begin(global read count),
   uploads, gci_pipeline. */
/* home rank and home-local id of a read among `world` ranks; returns how many of the ids 0 .. read_id-1 share that
   home (== *local).  Host arithmetic only (no context, no GPU): what the caller deals PAF lines by. */
int gci_shard_home(uint32_t read_id, int32_t world, int32_t* rank, uint32_t* local);
int gci_shard_config(gci_ctx* ctx, int32_t rank, int32_t world, const int32_t* contig_owner /* [n_contigs] */,
                     const uint8_t* gate_selected /* [n_contigs] or NULL: contigs selected on ANY rank */);
int gci_shard_alloc(gci_ctx* ctx, uint32_t max_reads, int32_t max_bam_files, gci_ipc_handle* out /* or NULL */);
int gci_shard_open(gci_ctx* ctx, const gci_ipc_handle* handles /* [world] */);
/* unmap the peers' areas (before any rank reallocates its own with gci_shard_alloc; the caller puts a barrier between) */
int gci_shard_close(gci_ctx* ctx);
/* contexts of one process (tests; several GPUs driven by one process): areas[r] = gci_shard_area of rank r's context */
void* gci_shard_area(gci_ctx* ctx);
int gci_shard_attach(gci_ctx* ctx, void* const* areas /* [world] */);

#ifdef __cplusplus
}
#endif
#endif /* GCI_CUDA_H */
