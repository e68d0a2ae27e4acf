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
#include <unistd.h>

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

int orc_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
