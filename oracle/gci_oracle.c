/*
 * CPU ORACLE (C restatement) — TEST INFRASTRUCTURE ONLY, never linked into the product.
 *
 * Plain C + pthreads restatement of the reference's hot loops, used (a) by tests as a second checker
 * next to oracle/gci_oracle.py and (b) by bench.py as the timed CPU baseline ("port").  Each function
 * cites the reference lines it follows (yeeus/GCI @ 455e19c7).  Parity pinning: tests/test_c_oracle.py
 * checks every function against oracle/gci_oracle.py, which is itself pinned to the unmodified
 * reference's outputs (tests/golden/).
 *
 * Threading mirrors what the reference parallelises: read_sam over reference windows (GCI.py:257-270)
 * -> records are gated in parallel; everything the reference runs serially (join, depth accumulate,
 * collapse) is also given all threads here where that is trivially possible, so the baseline is not
 * handicapped.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <unistd.h>
#include <zlib.h>

/* minimal fork-join: run fn(tid, nt, arg) on nt threads (libgomp is not in the image) */
typedef void (*par_fn)(int tid, int nt, void* arg);
typedef struct { par_fn fn; int tid, nt; void* arg; } par_task;
static void* par_entry(void* p) { par_task* t = (par_task*)p; t->fn(t->tid, t->nt, t->arg); return 0; }
static void run_parallel(par_fn fn, void* arg, int nt) {
  if (nt < 1) nt = 1;
  if (nt > 256) nt = 256;
  pthread_t th[256];
  par_task tk[256];
  for (int i = 1; i < nt; i++) { tk[i].fn = fn; tk[i].tid = i; tk[i].nt = nt; tk[i].arg = arg; pthread_create(&th[i], 0, par_entry, &tk[i]); }
  fn(0, nt, arg);
  for (int i = 1; i < nt; i++) pthread_join(th[i], 0);
}

#define NM_MISSING INT32_MIN

/* GCI.py:153-166 + pysam get_cigar_stats / reference_end.  pass[i]: 1 kept, 0 dropped,
   -1 KeyError (no NM), -2 / -3 ZeroDivisionError (clip / identity denominators). */
typedef struct {
  int64_t n; const int32_t *ref_id, *ref_start; const uint8_t* mapq; const uint16_t* flag; const int32_t* nm;
  const uint64_t* cigar_off; const uint32_t* cigar; const uint8_t* selected; int32_t n_contigs, map_qual;
  double ip, cp; int8_t* pass; int32_t* ref_end; int bad[256];
} gate_args;

static void gate_part(int tid, int nt, void* p) {
  gate_args* a = (gate_args*)p;
  int bad = 0;
  const int64_t i0 = a->n * tid / nt, i1 = a->n * (tid + 1) / nt;
  for (int64_t i = i0; i < i1; i++) {
    int64_t cnt[16] = {0};
    for (uint64_t k = a->cigar_off[i]; k < a->cigar_off[i + 1]; k++) cnt[a->cigar[k] & 15] += a->cigar[k] >> 4;
    const int64_t M = cnt[0], I = cnt[1], D = cnt[2], N = cnt[3], S = cnt[4], EQ = cnt[7], X = cnt[8];
    int64_t rlen = M + D + N + EQ + X;
    if (rlen == 0) rlen = 1;
    a->ref_end[i] = (int32_t)(a->ref_start[i] + rlen);
    a->pass[i] = 0;
    const int32_t c = a->ref_id[i];
    if (c < 0 || c >= a->n_contigs || !a->selected[c]) continue;
    if (a->flag[i] & (0x4 | 0x100 | 0x800)) continue;
    if ((int32_t)a->mapq[i] < a->map_qual) continue;
    if (a->nm[i] == NM_MISSING) { a->pass[i] = -1; bad |= 1; continue; }
    const int64_t mm = (int64_t)a->nm[i] - (I + D);
    const int64_t d1 = M + EQ + X + I + S;
    if (d1 == 0) { a->pass[i] = -2; bad |= 2; continue; }
    if (!((double)S / (double)d1 <= a->cp)) continue;
    const int64_t d2 = M + EQ + X + I + D;
    if (d2 == 0) { a->pass[i] = -3; bad |= 4; continue; }
    if (!((double)(M + EQ + X - mm) / (double)d2 >= a->ip)) continue;
    a->pass[i] = 1;
  }
  a->bad[tid] = bad;
}

int orc_gate(int64_t n, const int32_t* ref_id, const int32_t* ref_start, const uint8_t* mapq, const uint16_t* flag,
             const int32_t* nm, const uint64_t* cigar_off, const uint32_t* cigar, const uint8_t* selected,
             int32_t n_contigs, int32_t map_qual, double ip, double cp, int8_t* pass, int32_t* ref_end, int threads) {
  gate_args a = {n, ref_id, ref_start, mapq, flag, nm, cigar_off, cigar, selected, n_contigs, map_qual, ip, cp, pass,
                 ref_end, {0}};
  if (threads > 256) threads = 256;
  run_parallel(gate_part, &a, threads);
  int bad = 0;
  for (int t = 0; t < 256; t++) bad |= a.bad[t];
  return bad;
}

/* GCI.py:166-168, :268-270: dict overwrite in fetch order (contig header order, then file order) */
void orc_dedup(int64_t n, const int32_t* ref_id, const uint8_t* mapq, const uint32_t* read_id, const int8_t* pass,
               uint32_t n_reads, int32_t mq_cutoff, int64_t* win, uint8_t* highq) {
  for (uint32_t q = 0; q < n_reads; q++) win[q] = -1;
  for (int64_t i = 0; i < n; i++) {
    if (pass[i] != 1) continue;
    const uint32_t q = read_id[i];
    if (q >= n_reads) continue;
    const int64_t key = ((int64_t)ref_id[i] << 32) | i;
    if (key > win[q]) win[q] = key;
    if ((int32_t)mapq[i] >= mq_cutoff) highq[q] = 1;
  }
}

/* GCI.py:272-301.  Per file f: fc[f][q] contig or -1 when the read is absent, fs/fe/fq its segment. */
typedef struct {
  int32_t n_files; uint32_t n_reads; const int32_t *const *fc, *const *fs, *const *fe, *const *fq;
  const uint8_t* highq; double op; int32_t *oc, *os, *oe; int bad[256];
} join_args;

static void join_part(int tid, int nt, void* p) {
  join_args* a = (join_args*)p;
  int bad = 0;
  const uint32_t q0 = (uint32_t)((uint64_t)a->n_reads * tid / nt), q1 = (uint32_t)((uint64_t)a->n_reads * (tid + 1) / nt);
  const int32_t *const *fc = a->fc, *const *fs = a->fs, *const *fe = a->fe, *const *fq = a->fq;
  for (uint32_t q = q0; q < q1; q++) {
    int have = 0;
    int32_t c = -1, s = 0, e = 0;
    if (a->n_files == 1) {
      if (fc[0][q] >= 0) { have = 1; c = fc[0][q]; s = fs[0][q]; e = fe[0][q]; }
    } else {
      int comm = 1;
      for (int f = 0; f < a->n_files; f++) comm = comm && fc[f][q] >= 0;
      if (fc[0][q] >= 0 && (a->highq[q] || comm)) { have = 1; c = fc[0][q]; s = fs[0][q]; e = fe[0][q]; }
      for (int f = 1; f < a->n_files; f++) {
        if (fc[f][q] < 0) continue;
        if (have) {
          if (fc[f][q] == c) {
            const int32_t s2 = fs[f][q], e2 = fe[f][q];
            const int64_t ov = (int64_t)(e2 < e ? e2 : e) - (int64_t)(s2 > s ? s2 : s);
            if (fq[f][q] == 0) { bad |= 8; have = 0; }
            else if ((double)ov / (double)fq[f][q] < a->op) have = 0;
            else { s = s2 > s ? s2 : s; e = e2 < e ? e2 : e; }
          } else have = 0;
        } else if (a->highq[q]) { have = 1; c = fc[f][q]; s = fs[f][q]; e = fe[f][q]; }
      }
    }
    a->oc[q] = have ? c : -1;
    a->os[q] = s;
    a->oe[q] = e;
  }
  a->bad[tid] = bad;
}

int orc_join(int32_t n_files, uint32_t n_reads, const int32_t* const* fc, const int32_t* const* fs,
             const int32_t* const* fe, const int32_t* const* fq, const uint8_t* highq, double op, int32_t* oc,
             int32_t* os, int32_t* oe, int threads) {
  join_args a = {n_files, n_reads, fc, fs, fe, fq, highq, op, oc, os, oe, {0}};
  if (threads > 256) threads = 256;
  run_parallel(join_part, &a, threads);
  int bad = 0;
  for (int t = 0; t < 256; t++) bad |= a.bad[t];
  return bad;
}

static int64_t py_index(int64_t i, int64_t len) {
  if (i < 0) { i += len; if (i < 0) i = 0; } else if (i >= len) i = len;
  return i;
}

/* GCI.py:302-306: depths[target][start+fl : end-fl+1] += 1 (int64 like np.zeros(dtype=int)).
   Threads split the reference axis; each adds the part of every slice that falls into its range. */
typedef struct {
  uint32_t n_reads; const int32_t *sc, *ss, *se; int32_t fl, n_contigs; const int64_t *len, *off; int64_t* depth;
} depth_args;

static void depth_part(int tid, int nt, void* p) {
  depth_args* a = (depth_args*)p;
  const int64_t total = a->off[a->n_contigs];
  const int64_t g0 = total * tid / nt, g1 = total * (tid + 1) / nt;
  for (uint32_t q = 0; q < a->n_reads; q++) {
    const int32_t c = a->sc[q];
    if (c < 0) continue;
    const int64_t L = a->len[c];
    int64_t x = py_index((int64_t)a->ss[q] + a->fl, L), y = py_index((int64_t)a->se[q] - a->fl + 1, L);
    if (x >= y) continue;
    x += a->off[c];
    y += a->off[c];
    if (x < g0) x = g0;
    if (y > g1) y = g1;
    int64_t* d = a->depth;
    for (int64_t k = x; k < y; k++) d[k] += 1;
  }
}

void orc_depth(uint32_t n_reads, const int32_t* sc, const int32_t* ss, const int32_t* se, int32_t fl,
               int32_t n_contigs, const int64_t* len, const int64_t* off, int64_t* depth, int threads) {
  depth_args a = {n_reads, sc, ss, se, fl, n_contigs, len, off, depth};
  run_parallel(depth_part, &a, threads);
}

/* GCI.py:371-389: the per-base state machine of collapse_depth_range for one contig */
int64_t orc_collapse(const int64_t* depth, int64_t chr_len, int64_t leftmost, int64_t rightmost, int64_t fl,
                     int64_t start_pos, int64_t* out_s, int64_t* out_e, int64_t cap) {
  int64_t n_out = 0;
  int start_flag = 0, end_flag = 1;
  int64_t start = 0;
  const int64_t n = chr_len - 2 * fl;
  for (int64_t i = 0; i < n; i++) {
    const int64_t d = depth[i + fl];
    if (leftmost < d && d <= rightmost) {
      if (start_flag == 0) { start = i + fl; start_flag = 1; end_flag = 0; }
      if (i == chr_len - fl * 2 - 1) {
        if (n_out < cap) { out_s[n_out] = start + start_pos; out_e[n_out] = i + fl + 1 + start_pos; }
        n_out++;
      }
    } else {
      if (end_flag == 0) {
        if (i > fl) {
          if (n_out < cap) { out_s[n_out] = start + start_pos; out_e[n_out] = i + fl + start_pos; }
          n_out++;
        }
        end_flag = 1;
        start_flag = 0;
      }
    }
  }
  return n_out;
}

/* GCI.py:328 and :350 */
void orc_mask(int64_t* depth, int64_t L, int64_t s, int64_t e) {
  s = py_index(s, L);
  e = py_index(e, L);
  for (int64_t p = s; p < e; p++) depth[p] = 0;
}

typedef struct { const int64_t *a, *b; int64_t* out; int64_t n; } max_args;
static void max_part(int tid, int nt, void* p) {
  max_args* m = (max_args*)p;
  for (int64_t i = m->n * tid / nt; i < m->n * (tid + 1) / nt; i++) m->out[i] = m->a[i] > m->b[i] ? m->a[i] : m->b[i];
}
void orc_max(const int64_t* a, const int64_t* b, int64_t* out, int64_t n, int threads) {
  max_args m = {a, b, out, n};
  run_parallel(max_part, &m, threads);
}

/* ------------------------------------------------------------------------------------------------
 * PAF leg, GCI.py:211-254 with merge_alns_properties (:64-96) and get_average_identity (:49-61).
 * Lines of all PAF files are given concatenated (file f = lines [file_off[f], file_off[f+1])); the
 * reference's `synteny` dict is created once (:214), so file f elects over the kept lines of files 0..f.
 * Output per file and read: oc (contig or -1), os / oe (largest merged target block), oq (query length).
 * name_rank[c] = rank of contig c's name under str ordering (the (score, name) sort key of :252).
 * Returns a bit mask: 16 = ZeroDivisionError at :231 (alnlen 0), 32 = at :247 (query length 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int32_t lo, hi; } orc_pair;
static int pair_cmp(const void* a, const void* b) {
  const orc_pair *x = (const orc_pair*)a, *y = (const orc_pair*)b;
  if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
  if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
  return 0;
}
/* GCI.py:64-96: sorted pairs, merge touching / overlapping blocks; total merged length, and the longest block
   (the first one on ties: sorted(..., key=len, reverse=True) is stable) */
static int64_t merge_blocks(orc_pair* p, int n, int32_t* best_lo, int32_t* best_hi) {
  qsort(p, (size_t)n, sizeof(orc_pair), pair_cmp);
  int64_t total = 0, best = -1;
  int32_t lo = p[0].lo, hi = p[0].hi;
  for (int i = 0; i < n; i++) {
    if (hi >= p[i].lo) { if (hi < p[i].hi) hi = p[i].hi; }
    else {
      const int64_t len = (int64_t)hi - lo;
      total += len;
      if (len > best) { best = len; *best_lo = lo; *best_hi = hi; }
      lo = p[i].lo; hi = p[i].hi;
    }
  }
  const int64_t len = (int64_t)hi - lo;
  total += len;
  if (len > best) { *best_lo = lo; *best_hi = hi; }
  return total;
}

typedef struct {
  int32_t n_files; const int64_t* file_off; const uint32_t* read_id; const int32_t *qlen, *q0, *q1, *ref, *t0, *t1, *nmatch,
      *alnlen, *mapq; const uint8_t* selected; int32_t n_contigs; const int32_t* name_rank; uint32_t n_reads;
  int32_t map_qual, mq_cutoff; double ip;
  int8_t* keep; double* ident;                 /* per line */
  const int64_t* grp_off; const int64_t* grp;  /* kept lines grouped by read, in line order */
  int32_t *oc, *os, *oe, *oq; uint8_t* highq; int bad[256];
} paf_args;

static void paf_gate_part(int tid, int nt, void* p) {
  paf_args* a = (paf_args*)p;
  const int64_t n = a->file_off[a->n_files];
  int bad = 0;
  for (int64_t i = n * tid / nt; i < n * (tid + 1) / nt; i++) {
    a->keep[i] = 0;
    const int32_t t = a->ref[i];
    if (t < 0 || t >= a->n_contigs || !a->selected[t]) continue;           /* :220 */
    if (a->alnlen[i] == 0) { bad |= 16; continue; }                         /* :231 raises */
    const double id = (double)a->nmatch[i] / (double)a->alnlen[i];
    a->ident[i] = id;
    if (a->mapq[i] >= a->map_qual && id >= a->ip && a->read_id[i] < a->n_reads) a->keep[i] = 1;   /* :232 */
  }
  a->bad[tid] |= bad;
}

static void paf_elect_part(int tid, int nt, void* p) {
  paf_args* a = (paf_args*)p;
  int bad = 0;
  int cap = 64;
  orc_pair* pr = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)cap);
  const uint32_t r0 = (uint32_t)((uint64_t)a->n_reads * tid / nt), r1 = (uint32_t)((uint64_t)a->n_reads * (tid + 1) / nt);
  for (int32_t f = 0; f < a->n_files; f++) {
    const int64_t limit = a->file_off[f + 1];
    int32_t *oc = a->oc + (size_t)f * a->n_reads, *os = a->os + (size_t)f * a->n_reads,
            *oe = a->oe + (size_t)f * a->n_reads, *oq = a->oq + (size_t)f * a->n_reads;
    for (uint32_t r = r0; r < r1; r++) {
      const int64_t* g = a->grp + a->grp_off[r];
      int n = 0;
      const int64_t gn = a->grp_off[r + 1] - a->grp_off[r];
      while (n < gn && g[n] < limit) n++;                                   /* lines of files 0..f */
      oc[r] = -1; os[r] = 0; oe[r] = 0; oq[r] = 0;
      if (n == 0) continue;
      if (n > cap) { cap = n * 2; pr = (orc_pair*)realloc(pr, sizeof(orc_pair) * (size_t)cap); }
      int have = 0, failed = 0;
      double best_score = 0;
      int32_t best_rank = 0;
      for (int i = 0; i < n && !failed; i++) {
        const int32_t t = a->ref[g[i]];
        int seen = 0;
        for (int j = 0; j < i && !seen; j++) seen = a->ref[g[j]] == t;
        if (seen) continue;                                                 /* dict order = first appearance */
        double sum = 0;
        int cnt = 0, m = 0;
        for (int j = i; j < n; j++)
          if (a->ref[g[j]] == t) { sum = sum + a->ident[g[j]]; cnt++; pr[m].lo = a->q0[g[j]]; pr[m].hi = a->q1[g[j]]; m++; }
        const int32_t ql = a->qlen[g[i]];                                   /* alns[0][0], :246 */
        if (ql == 0) { bad |= 32; failed = 1; break; }
        int32_t dl = 0, dh = 0;
        const int64_t mapped = merge_blocks(pr, m, &dl, &dh);
        const double rate = (double)mapped / (double)ql;                   /* :247 */
        const double score = (sum / (double)cnt) * rate;                   /* :248-249 */
        m = 0;
        for (int j = i; j < n; j++)
          if (a->ref[g[j]] == t) { pr[m].lo = a->t0[g[j]]; pr[m].hi = a->t1[g[j]]; m++; }
        int32_t tl = 0, th = 0;
        merge_blocks(pr, m, &tl, &th);
        const int32_t rank = a->name_rank[t];
        if (!have || score > best_score || (score == best_score && rank > best_rank)) {   /* :252 */
          have = 1; best_score = score; best_rank = rank;
          oc[r] = t; os[r] = tl; oe[r] = th; oq[r] = ql;
        }
      }
      if (failed) oc[r] = -1;
    }
  }
  free(pr);
  a->bad[tid] |= bad;
}

int orc_paf_leg(int32_t n_files, const int64_t* file_off, const uint32_t* read_id, const int32_t* qlen, const int32_t* q0,
                const int32_t* q1, const int32_t* ref, const int32_t* t0, const int32_t* t1, const int32_t* nmatch,
                const int32_t* alnlen, const int32_t* mapq, const uint8_t* selected, int32_t n_contigs,
                const int32_t* name_rank, uint32_t n_reads, int32_t map_qual, int32_t mq_cutoff, double ip, int32_t* oc,
                int32_t* os, int32_t* oe, int32_t* oq, uint8_t* highq, int threads) {
  const int64_t n = file_off[n_files];
  paf_args a;
  memset(&a, 0, sizeof a);
  a.n_files = n_files; a.file_off = file_off; a.read_id = read_id; a.qlen = qlen; a.q0 = q0; a.q1 = q1; a.ref = ref;
  a.t0 = t0; a.t1 = t1; a.nmatch = nmatch; a.alnlen = alnlen; a.mapq = mapq; a.selected = selected;
  a.n_contigs = n_contigs; a.name_rank = name_rank; a.n_reads = n_reads; a.map_qual = map_qual; a.mq_cutoff = mq_cutoff;
  a.ip = ip; a.oc = oc; a.os = os; a.oe = oe; a.oq = oq; a.highq = highq;
  a.keep = (int8_t*)malloc((size_t)(n > 0 ? n : 1));
  a.ident = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (threads > 256) threads = 256;
  run_parallel(paf_gate_part, &a, threads);
  /* group the kept lines by read, in line order (a counting sort: stable) */
  int64_t* goff = (int64_t*)calloc((size_t)n_reads + 2, sizeof(int64_t));
  for (int64_t i = 0; i < n; i++)
    if (a.keep[i]) { goff[read_id[i] + 1]++; if (mapq[i] >= mq_cutoff) highq[read_id[i]] = 1; }   /* :238 */
  for (uint32_t r = 0; r < n_reads; r++) goff[r + 1] += goff[r];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_reads + 1));
  memcpy(cur, goff, sizeof(int64_t) * ((size_t)n_reads + 1));
  int64_t* grp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(goff[n_reads] > 0 ? goff[n_reads] : 1));
  for (int64_t i = 0; i < n; i++)
    if (a.keep[i]) grp[cur[read_id[i]]++] = i;
  a.grp_off = goff; a.grp = grp;
  run_parallel(paf_elect_part, &a, threads);
  int bad = 0;
  for (int t = 0; t < 256; t++) bad |= a.bad[t];
  free(a.keep); free(a.ident); free(goff); free(cur); free(grp);
  return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Order-independent 64-bit checksum of one depth array (test / bench parity gate at sizes where whole
 * arrays are not compared): sum over positions of (depth + 1) * splitmix64(position), mod 2^64.
 * The CUDA side computes the same number (gci_depth_hash).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
typedef struct { const int64_t* d; int64_t n; uint64_t part[256]; } hash_args;
static void hash_part(int tid, int nt, void* p) {
  hash_args* a = (hash_args*)p;
  uint64_t h = 0;
  for (int64_t i = a->n * tid / nt; i < a->n * (tid + 1) / nt; i++) h += (uint64_t)(a->d[i] + 1) * splitmix64((uint64_t)i);
  a->part[tid] = h;
}
uint64_t orc_depth_hash(const int64_t* depth, int64_t n, int threads) {
  hash_args a;
  memset(&a, 0, sizeof a);
  a.d = depth; a.n = n;
  if (threads > 256) threads = 256;
  if (threads < 1) threads = 1;
  run_parallel(hash_part, &a, threads);
  uint64_t h = 0;
  for (int t = 0; t < threads; t++) h += a.part[t];
  return h;
}

/* write_depth, GCI.py:99-143: per contig `threads` parts of stp = 1 + (len - 1) / threads positions, each part its own
   gzip file written with gzip.open(..., 'wb') (compresslevel 9): ">name\n" in part 0, then one decimal per line; the
   parts are concatenated in order.  Here every part is formatted and deflated (zlib, level 9, gzip wrapper) by one of
   `workers` threads into memory; nothing touches the disk.  out == NULL: only the sizes are produced (the timed
   baseline); otherwise the members are copied to out in file order.  Returns the total number of bytes, -1 on error. */
typedef struct {
  const int64_t* d; int64_t n, stp; int parts; const char* name; int level;
  unsigned char** buf; int64_t* len; int err;
} wd_args;
static int wd_one(const wd_args* a, int part) {
  const int64_t lft = (int64_t)part * a->stp, rgh = lft + a->stp < a->n ? lft + a->stp : a->n;
  z_stream z;
  memset(&z, 0, sizeof z);
  if (deflateInit2(&z, a->level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
  size_t cap = (size_t)(rgh - lft) / 8 + 4096, used = 0;
  unsigned char* out = (unsigned char*)malloc(cap);
  char text[1 << 16];
  size_t fill = 0;
  int rc = 0;
  if (part == 0) fill = (size_t)snprintf(text, sizeof text, ">%s\n", a->name);
  for (int64_t i = lft; i <= rgh && rc == 0; i++) {
    if (i < rgh) {                                  /* f'{depth}\n' */
      long long v = (long long)a->d[i];
      char tmp[24];
      int k = 0, neg = v < 0;
      unsigned long long u = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
      do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
      if (neg) text[fill++] = '-';
      while (k) text[fill++] = tmp[--k];
      text[fill++] = '\n';
    }
    if (fill + 32 > sizeof text || i == rgh) {
      z.next_in = (unsigned char*)text;
      z.avail_in = (uInt)fill;
      const int flush = i == rgh ? Z_FINISH : Z_NO_FLUSH;
      for (;;) {
        if (cap - used < 4096) {
          cap *= 2;
          unsigned char* grown = (unsigned char*)realloc(out, cap);
          if (!grown) { rc = -1; break; }
          out = grown;
        }
        z.next_out = out + used;
        z.avail_out = (uInt)(cap - used > 0x40000000u ? 0x40000000u : cap - used);
        const uInt before = z.avail_out;
        const int r = deflate(&z, flush);
        used += before - z.avail_out;
        if (r == Z_STREAM_ERROR) { rc = -1; break; }
        if (flush == Z_FINISH ? r == Z_STREAM_END : z.avail_in == 0) break;
      }
      fill = 0;
    }
  }
  deflateEnd(&z);
  a->buf[part] = out;
  a->len[part] = (int64_t)used;
  return rc;
}
static void wd_part(int tid, int nt, void* p) {
  wd_args* a = (wd_args*)p;
  for (int part = tid; part < a->parts; part += nt)
    if (wd_one(a, part) != 0) a->err = 1;
}
int64_t orc_write_depth(const int64_t* depth, int64_t n, const char* name, int threads, int level, int workers,
                        unsigned char* out, int64_t out_cap) {
  if (n <= 0 || threads < 1) return 0;
  wd_args a;
  memset(&a, 0, sizeof a);
  a.d = depth; a.n = n; a.name = name; a.level = level;
  a.stp = 1 + (n - 1) / threads;
  a.parts = (int)((n + a.stp - 1) / a.stp);
  a.buf = (unsigned char**)calloc((size_t)a.parts, sizeof *a.buf);
  a.len = (int64_t*)calloc((size_t)a.parts, sizeof *a.len);
  run_parallel(wd_part, &a, workers < a.parts ? workers : a.parts);
  int64_t total = 0;
  for (int k = 0; k < a.parts; k++) {
    if (out && !a.err && total + a.len[k] <= out_cap) memcpy(out + total, a.buf[k], (size_t)a.len[k]);
    total += a.len[k];
    free(a.buf[k]);
  }
  free(a.buf);
  free(a.len);
  return a.err ? -1 : total;
}

int orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
