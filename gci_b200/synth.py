"""Seeded synthetic long-read alignments in the columnar schema.

The generator emits `AlnTable`/`PafTable` columns directly (vectorised numpy), so
the same function serves the small parity cases and the BASELINE.json-sized
bench configs.  It exercises every gate of the filter stage: soft clips around
the `-cp` limit, a MAPQ mix around `-mq`/`--mq-cutoff`, secondary /
supplementary / unmapped flags, hard clips, `=`/`X` and `M` style CIGARs, NM
values around the `-ip` limit, and — for the second aligner — shifted,
relocated and missing reads for the `-op` join.  Coverage holes and N-runs make
the BED / .gci outputs non-trivial.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .records import (AlnTable, PafTable, ContigTable, OP_M, OP_I, OP_D, OP_S, OP_H, OP_EQ, OP_X)


@dataclass
class SynthSpec:
    contig_lengths: list
    coverage: float = 30.0
    read_mean: float = 15000.0
    read_sigma: float = 0.35          # lognormal sigma
    read_min: int = 5000
    read_max: int = 30000
    events_per_read: float = 15.0     # I/D/X events -> ops ~ 2*events+1  (HiFi ~30 ops)
    events_per_base: float = 0.0      # ONT: 0.04 events/base -> 0.08 ops/base (overrides events_per_read)
    hole_fraction: float = 0.003      # fraction of the genome in zero-coverage holes
    hole_mean: float = 2000.0
    n_runs_per_contig: int = 2        # N-runs in the assembly FASTA
    n_run_mean: float = 500.0
    seed: int = 20240633
    contig_prefix: str = "chr"
    contig_names: list | None = None


@dataclass
class SynthData:
    contigs: ContigTable
    bam: AlnTable
    n_reads: int
    holes: list                       # per contig [(s,e)]
    n_runs: list                      # per contig [(s,e)]
    aligned_bases: int
    spec: SynthSpec = None


def _intervals(rng, length, total_frac, mean_len, count=None):
    """Sorted, disjoint random intervals inside [1000, length-1000)."""
    if length < 5000:
        return []
    if count is None:
        count = int(round(length * total_frac / mean_len))
    if count <= 0:
        return []
    lens = np.maximum(1, rng.lognormal(np.log(mean_len) - 0.125, 0.5, count)).astype(np.int64)
    starts = np.sort(rng.integers(1000, max(1001, length - 1000 - int(lens.max())), count))
    out = []
    last = 0
    for s, l in zip(starts.tolist(), lens.tolist()):
        if s <= last + 50:
            continue
        e = min(s + l, length - 1000)
        if e <= s:
            continue
        out.append((s, e))
        last = e
    return out


def make_reads(spec: SynthSpec) -> SynthData:
    rng = np.random.default_rng(spec.seed)
    lengths = np.asarray(spec.contig_lengths, dtype=np.int64)
    nct = len(lengths)
    names = spec.contig_names or [f"{spec.contig_prefix}{i + 1}" for i in range(nct)]
    contigs = ContigTable(names, lengths)
    holes = [_intervals(rng, int(l), spec.hole_fraction, spec.hole_mean) for l in lengths]
    n_runs = [_intervals(rng, int(l), 0.0, spec.n_run_mean, count=spec.n_runs_per_contig) for l in lengths]

    # ---- reads per contig -----------------------------------------------------------
    n_per = np.maximum(1, (lengths * spec.coverage / spec.read_mean).astype(np.int64))
    n = int(n_per.sum())
    ref_id = np.repeat(np.arange(nct, dtype=np.int32), n_per)
    mu = np.log(spec.read_mean) - spec.read_sigma ** 2 / 2
    want = np.clip(rng.lognormal(mu, spec.read_sigma, n), spec.read_min, spec.read_max).astype(np.int64)
    want = np.minimum(want, np.maximum(50, lengths[ref_id] - 1))

    # ---- CIGAR bodies: M-run, (event, M-run)*k ---------------------------------------
    if spec.events_per_base > 0:
        k = rng.poisson(want * spec.events_per_base).astype(np.int64)
    else:
        k = rng.poisson(spec.events_per_read, n).astype(np.int64)
    k = np.minimum(k, np.maximum(0, want // 4))
    eqx = rng.random(n) < 0.2                        # reads written with '=' / 'X' instead of 'M'
    n_body = 2 * k + 1
    body_off = np.concatenate([[0], np.cumsum(n_body)])
    tot = int(body_off[-1])
    rec_of = np.repeat(np.arange(n, dtype=np.int64), n_body)
    pos_in = np.arange(tot, dtype=np.int64) - body_off[rec_of]
    is_event = (pos_in & 1) == 1
    # match-run lengths: exponential around want/(k+1), >= 1
    mean_run = (want / (k + 1))[rec_of]
    run_len = np.maximum(1, rng.exponential(1.0, tot) * mean_run).astype(np.int64)
    ev_kind = rng.random(tot)
    ev_len = np.minimum(rng.geometric(0.6, tot), 30).astype(np.int64)
    op = np.where(is_event,
                  np.where(ev_kind < 0.35, OP_I, np.where(ev_kind < 0.7, OP_D, OP_X)),
                  np.where(eqx[rec_of], OP_EQ, OP_M)).astype(np.int64)
    # mismatch events inside an 'M'-style read are invisible in the CIGAR: make them M
    op = np.where(is_event & (op == OP_X) & ~eqx[rec_of], OP_M, op)
    ln = np.where(is_event, ev_len, run_len)

    def seg_sum(mask_vals):
        cs = np.concatenate([[0], np.cumsum(mask_vals)])
        return cs[body_off[1:]] - cs[body_off[:-1]]

    ins = seg_sum(np.where(op == OP_I, ln, 0))
    dele = seg_sum(np.where(op == OP_D, ln, 0))
    xs = seg_sum(np.where(op == OP_X, ln, 0))
    hidden_mm = seg_sum(np.where(is_event & (op == OP_M), ln, 0))
    ref_len = seg_sum(np.where((op == OP_M) | (op == OP_D) | (op == OP_EQ) | (op == OP_X), ln, 0))
    qry_body = seg_sum(np.where((op == OP_M) | (op == OP_I) | (op == OP_EQ) | (op == OP_X), ln, 0))

    # ---- clips ----------------------------------------------------------------------
    clip_mode = rng.random(n)
    has_soft = clip_mode < 0.05
    soft_frac = rng.random(n) * 0.15
    soft_total = np.where(has_soft, (qry_body * soft_frac / (1 - soft_frac)).astype(np.int64), 0)
    soft_left = (soft_total * rng.random(n)).astype(np.int64)
    soft_right = soft_total - soft_left
    has_hard = (clip_mode >= 0.05) & (clip_mode < 0.06)
    hard_left = np.where(has_hard, rng.integers(1, 5000, n), 0).astype(np.int64)

    pre_ops = (hard_left > 0).astype(np.int64) + (soft_left > 0).astype(np.int64)
    post_ops = (soft_right > 0).astype(np.int64)
    n_ops = n_body + pre_ops + post_ops
    off = np.concatenate([[0], np.cumsum(n_ops)])
    cigar = np.zeros(int(off[-1]), dtype=np.uint32)
    # body
    dst = np.repeat(off[:-1] + pre_ops, n_body) + pos_in
    cigar[dst] = ((ln << 4) | op).astype(np.uint32)
    # leading hard, then soft
    idx = np.flatnonzero(hard_left > 0)
    cigar[off[idx]] = ((hard_left[idx] << 4) | OP_H).astype(np.uint32)
    idx = np.flatnonzero(soft_left > 0)
    cigar[off[idx] + (hard_left[idx] > 0)] = ((soft_left[idx] << 4) | OP_S).astype(np.uint32)
    idx = np.flatnonzero(soft_right > 0)
    cigar[off[idx + 1] - 1] = ((soft_right[idx] << 4) | OP_S).astype(np.uint32)

    # ---- NM (identity gate) -----------------------------------------------------------
    extra_mm = np.where(eqx, 0, rng.poisson(2.0, n))           # mismatches hidden inside M runs
    nm = ins + dele + xs + hidden_mm + extra_mm
    bad_id = rng.random(n) < 0.02                               # low-identity reads
    nm = np.where(bad_id, nm + (ref_len * rng.uniform(0.08, 0.2, n)).astype(np.int64), nm)
    edge = rng.random(n) < 0.005                                # sit exactly on the 0.9 boundary when possible
    den = (qry_body - ins) + ins + dele                         # M+=+X + I + D
    nm_edge = ins + dele + (den // 10)
    nm = np.where(edge & ~eqx, nm_edge, nm)

    # ---- flags / mapq -----------------------------------------------------------------
    u = rng.random(n)
    flag = np.where(rng.random(n) < 0.5, 16, 0).astype(np.int64)
    flag = np.where(u < 0.03, flag | 0x100, flag)
    flag = np.where((u >= 0.03) & (u < 0.06), flag | 0x800, flag)
    flag = np.where((u >= 0.06) & (u < 0.065), flag | 0x4, flag)
    m = rng.random(n)
    mapq = np.where(m < 0.85, 60,
                    np.where(m < 0.90, rng.integers(30, 50, n),
                             np.where(m < 0.95, rng.integers(1, 30, n), 0))).astype(np.int64)

    # ---- placement (avoid holes) ------------------------------------------------------
    room = np.maximum(1, lengths[ref_id] - ref_len)
    start = (rng.random(n) * room).astype(np.int64)
    # keep every ref_end >= 14 (avoid the negative-slice wrap quirk unless tested on purpose)
    keep = np.ones(n, dtype=bool)
    keep &= (start + ref_len) <= lengths[ref_id]
    for c in range(nct):
        if not holes[c]:
            continue
        hs = np.array([h[0] for h in holes[c]])
        he = np.array([h[1] for h in holes[c]])
        sel = np.flatnonzero(ref_id == c)
        s = start[sel]
        e = s + ref_len[sel]
        j = np.searchsorted(he, s, side="right")              # first hole ending after read start
        jj = np.minimum(j, len(hs) - 1)
        hit = (j < len(hs)) & (hs[jj] < e)
        keep[sel[hit]] = False
    keep &= (start + ref_len) >= 64

    # primary reads get their own read id; secondary/supplementary reuse a primary's id
    is_extra = (flag & (0x100 | 0x800)) != 0
    read_id = np.zeros(n, dtype=np.int64)
    prim = np.flatnonzero(~is_extra)
    read_id[prim] = np.arange(len(prim))
    ext = np.flatnonzero(is_extra)
    if len(prim):
        read_id[ext] = rng.integers(0, len(prim), len(ext))
    n_reads = max(1, len(prim))

    order = np.lexsort((start, ref_id))
    order = order[keep[order]]
    tab = AlnTable(ref_id, start, mapq, flag, nm, qry_body + soft_total, read_id,
                   off.astype(np.uint64), cigar).take(order)
    aligned = int(ref_len[order].sum())
    return SynthData(contigs, tab, n_reads, holes, n_runs, aligned, spec)


def second_aligner(data: SynthData, seed: int | None = None, as_paf: bool = False):
    """The same reads through a second aligner: 90 % identical
    coordinates, 5 % shifted by U(0, 0.2·len), 3 % on another contig, 2 % missing
    (and the caller may drop 2 % from the first file with `drop_reads`)."""
    rng = np.random.default_rng(data.spec.seed + 7 if seed is None else seed)
    t = data.bam
    n = t.n_records
    lengths = data.contigs.lengths
    ref_len = t.ref_len()
    u = rng.random(n)
    keep = u >= 0.02
    shift = (u >= 0.02) & (u < 0.07)
    move = (u >= 0.07) & (u < 0.10)
    start = t.ref_start.astype(np.int64).copy()
    ref_id = t.ref_id.astype(np.int64).copy()
    dlt = (rng.random(n) * 0.2 * ref_len).astype(np.int64) * np.where(rng.random(n) < 0.5, -1, 1)
    start = np.where(shift, start + dlt, start)
    if len(lengths) > 1:
        ref_id = np.where(move, (ref_id + 1 + rng.integers(0, len(lengths) - 1, n)) % len(lengths), ref_id)
    else:
        start = np.where(move, start + ref_len + 1000, start)
    start = np.clip(start, 0, np.maximum(0, lengths[ref_id] - ref_len))
    keep &= (start + ref_len) >= 64
    keep &= (start + ref_len) <= lengths[ref_id]
    # second aligner reports its own MAPQ for a tenth of the reads
    mapq = t.mapq.astype(np.int64).copy()
    re_mq = rng.random(n) < 0.1
    mapq = np.where(re_mq, rng.choice(np.array([0, 20, 35, 45, 60]), n), mapq)
    order = np.lexsort((start, ref_id))
    order = order[keep[order]]
    tmp = AlnTable(ref_id, start, mapq, t.flag, t.nm, t.qlen, t.read_id, t.cigar_off, t.cigar)
    out = tmp.take(order)
    if not as_paf:
        return out
    return aln_to_paf(out)


def aln_to_paf(t: AlnTable, drop_unmapped: bool = True) -> PafTable:
    """A PAF view of BAM records (what `minimap2` would print): query span excludes
    clips, query length includes them (S and H), nmatch = M+=+X - mismatches,
    alnlen = M+=+X+I+D."""
    st = t.op_sums()
    M, I, D, S, H, EQ, X = (st[:, k] for k in (OP_M, OP_I, OP_D, OP_S, OP_H, OP_EQ, OP_X))
    off = t.cigar_off.astype(np.int64)
    has = off[1:] > off[:-1]
    first = np.where(has, t.cigar[np.minimum(off[:-1], max(0, len(t.cigar) - 1))] if len(t.cigar) else 0, 0).astype(np.int64)
    second = np.where(off[1:] - off[:-1] > 1,
                      t.cigar[np.minimum(off[:-1] + 1, max(0, len(t.cigar) - 1))] if len(t.cigar) else 0, 0).astype(np.int64)
    left = np.where((first & 15) == OP_S, first >> 4, 0)
    left = np.where((first & 15) == OP_H, (first >> 4) + np.where((second & 15) == OP_S, second >> 4, 0), left)
    body = M + EQ + X
    qlen = body + I + S + H
    qstart = left
    qend = qstart + body + I
    mm = np.maximum(0, t.nm.astype(np.int64) - I - D)
    nmatch = np.maximum(0, body - mm)
    alnlen = body + I + D
    ref_len = t.ref_len()
    keep = np.ones(t.n_records, dtype=bool)
    if drop_unmapped:
        keep &= (t.flag & 0x4) == 0
    keep &= alnlen > 0
    k = np.flatnonzero(keep)
    return PafTable(t.read_id[k], qlen[k], qstart[k], qend[k], t.ref_id[k], t.ref_start[k],
                    t.ref_start[k].astype(np.int64) + ref_len[k], nmatch[k], alnlen[k], t.mapq[k].astype(np.int32))


def drop_reads(tab: AlnTable, frac: float, seed: int) -> AlnTable:
    rng = np.random.default_rng(seed)
    keep = np.flatnonzero(rng.random(tab.n_records) >= frac)
    return tab.take(keep)


# ---------------------------------------------------------------------------------------------------
# whole-genome workloads (BASELINE.json configs 2-4): one BAM (+ one PAF from a second aligner)
# ---------------------------------------------------------------------------------------------------
# CHM13v2.0-like chromosome lengths (chr1..chr22, chrX, chrY): 24 contigs, 3.117 Gbp
CHM13_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
CHM13_LENGTHS = [248_387_328, 242_696_752, 201_105_948, 193_574_945, 182_045_439, 172_126_628, 160_567_428,
                 146_259_331, 150_617_247, 134_758_134, 135_127_769, 133_324_548, 113_566_686, 101_161_492,
                 99_753_195, 96_330_374, 84_276_897, 80_542_538, 61_707_364, 66_210_255, 45_090_682, 51_324_926,
                 154_259_566, 62_460_029]


def paf_second_aligner(paf: PafTable, lengths, seed: int, split_frac=0.03, alt_frac=0.02, tie_frac=0.003) -> PafTable:
    """The reads of `paf` (one line per BAM record, `aln_to_paf`) as a second aligner would report them
    (SURVEY.md §8d): 2 % missing, 5 % shifted by U(0, 0.2 len), 3 % placed on another contig, own MAPQ for a
    tenth; plus what only a PAF has (GCI.py:211-254 elects among them): `split_frac` of the lines come as two
    blocks of the same read on the same contig, `alt_frac` of the reads carry a weaker extra line on another
    contig, `tie_frac` an identical line on another contig (equal score: the contig NAME decides, :252)."""
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, np.int64)
    n = paf.n_records
    nct = len(lengths)
    cols = {k: getattr(paf, k).astype(np.int64) for k in
            ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")}
    span = cols["tend"] - cols["tstart"]
    u = rng.random(n)
    keep = u >= 0.02
    shift = (u >= 0.02) & (u < 0.07)
    move = (u >= 0.07) & (u < 0.10)
    dlt = (rng.random(n) * 0.2 * span).astype(np.int64) * np.where(rng.random(n) < 0.5, -1, 1)
    ts = np.where(shift, cols["tstart"] + dlt, cols["tstart"])
    ref = cols["ref_id"].copy()
    if nct > 1:
        ref = np.where(move, (ref + 1 + rng.integers(0, nct - 1, n)) % nct, ref)
    else:
        ts = np.where(move, ts + span + 1000, ts)
    ts = np.clip(ts, 0, np.maximum(0, lengths[ref] - span))
    keep &= (ts + span) <= lengths[ref]
    cols["tstart"], cols["tend"], cols["ref_id"] = ts, ts + span, ref
    re_mq = rng.random(n) < 0.1
    cols["mapq"] = np.where(re_mq, rng.choice(np.array([0, 20, 35, 45, 60]), n), cols["mapq"])

    def rows(mask):
        idx = np.flatnonzero(mask & keep)
        return {k: v[idx] for k, v in cols.items()}, idx

    base, base_idx = rows(np.ones(n, bool))
    parts = [base]
    order_key = [base_idx.astype(np.float64)]
    # split lines: [q0, qm) + [qm - ov, q1) on the query, the same cut on the target
    v = rng.random(n)
    sp, sp_idx = rows(v < split_frac)
    if len(sp_idx):
        frac = rng.uniform(0.2, 0.8, len(sp_idx))
        qcut = (sp["qstart"] + (sp["qend"] - sp["qstart"]) * frac).astype(np.int64)
        tcut = (sp["tstart"] + (sp["tend"] - sp["tstart"]) * frac).astype(np.int64)
        ov = rng.integers(-50, 50, len(sp_idx))
        a_len = (sp["alnlen"] * frac).astype(np.int64).clip(1)
        a_match = (sp["nmatch"] * frac).astype(np.int64).clip(0)
        first = dict(sp, qend=qcut, tend=tcut, alnlen=a_len, nmatch=np.minimum(a_match, a_len))
        b_len = (sp["alnlen"] - a_len).clip(1)
        second = dict(sp, qstart=np.maximum(sp["qstart"], qcut - ov), tstart=np.maximum(sp["tstart"], tcut - ov),
                      alnlen=b_len, nmatch=np.minimum((sp["nmatch"] - a_match).clip(0), b_len))
        # the whole line of these reads is replaced by the two blocks
        gone = np.isin(base_idx, sp_idx)
        parts[0] = {k: x[~gone] for k, x in base.items()}
        order_key[0] = order_key[0][~gone]
        parts += [first, second]
        order_key += [sp_idx + 0.1, sp_idx + 0.2]
    if nct > 1:
        for lo, hi, tie in ((split_frac, split_frac + alt_frac, False),
                            (split_frac + alt_frac, split_frac + alt_frac + tie_frac, True)):
            al, al_idx = rows((v >= lo) & (v < hi))
            if not len(al_idx):
                continue
            other = (al["ref_id"] + 1 + rng.integers(0, nct - 1, len(al_idx))) % nct
            sp_len = al["tend"] - al["tstart"]
            t0 = (rng.random(len(al_idx)) * np.maximum(1, lengths[other] - sp_len)).astype(np.int64)
            ok = t0 + sp_len <= lengths[other]
            extra = dict(al, ref_id=other, tstart=t0, tend=t0 + sp_len)
            if not tie:
                cover = rng.uniform(0.3, 1.0, len(al_idx))
                extra["qend"] = (al["qstart"] + (al["qend"] - al["qstart"]) * cover).astype(np.int64)
                extra["nmatch"] = (al["nmatch"] * rng.uniform(0.93, 1.0, len(al_idx))).astype(np.int64)
            extra = {k: x[ok] for k, x in extra.items()}
            parts.append(extra)
            # half of the extra lines arrive before the read's main line, half after
            order_key.append(al_idx[ok] + np.where(rng.random(int(ok.sum())) < 0.5, -0.5, 0.5))
    key = np.concatenate(order_key)
    order = np.argsort(key, kind="stable")
    out = {k: np.concatenate([p[k] for p in parts])[order] for k in cols}
    return PafTable(out["read_id"], out["qlen"], out["qstart"], out["qend"], out["ref_id"], out["tstart"], out["tend"],
                    out["nmatch"], out["alnlen"], out["mapq"])


@dataclass
class GenomeWorkload:
    contigs: ContigTable
    bam: AlnTable                     # first aligner, coordinate sorted
    paf: PafTable                     # second aligner (None without `with_paf`)
    n_reads: int
    holes: list
    n_runs: list
    aligned_bases: int                # sum over the records of BOTH files of ref_end - ref_start


def concat_aln(tabs) -> AlnTable:
    ops = np.cumsum([0] + [t.n_ops for t in tabs])
    off = np.concatenate([np.zeros(1, np.uint64)] + [t.cigar_off[1:] + np.uint64(o) for t, o in zip(tabs, ops)])
    return AlnTable(*[np.concatenate([getattr(t, k) for t in tabs]) for k in
                      ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id")], off,
                    np.concatenate([t.cigar for t in tabs]))


def concat_paf(tabs) -> PafTable:
    return PafTable(*[np.concatenate([getattr(t, k) for t in tabs]) for k in
                      ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")])


def make_genome(lengths, names=None, coverage=30.0, seed=20240635, with_paf=True, drop_first=0.02, threads=0,
                contig_ids=None, **spec_kw) -> GenomeWorkload:
    """One synthetic read set over a whole genome, generated contig by contig (seed + contig index) on a thread
    pool: a minimap2-like BAM and, with `with_paf`, the same reads as a winnowmap-like PAF.  `contig_ids`
    restricts the BAM to those contigs (a rank's shard of a multi-GPU run; read ids stay global)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    lengths = [int(x) for x in lengths]
    nct = len(lengths)
    names = list(names) if names is not None else [f"chr{i + 1}" for i in range(nct)]
    todo = list(range(nct)) if contig_ids is None else list(contig_ids)

    def one(c):
        d = make_reads(SynthSpec([lengths[c]], coverage=coverage, seed=seed + c, **spec_kw))
        # read ids in order of first appearance in the coordinate-sorted BAM: what the decoder's name interner
        # hands out when it reads the file (gci_b200/io.py), so records of one neighbourhood carry neighbouring ids
        n = d.bam.n_records
        first = np.full(d.n_reads, n, np.int64)
        np.minimum.at(first, d.bam.read_id, np.arange(n, dtype=np.int64))
        new_id = np.empty(d.n_reads, np.uint32)
        new_id[np.argsort(first, kind="stable")] = np.arange(d.n_reads, dtype=np.uint32)
        d.bam.read_id = new_id[d.bam.read_id]
        paf = aln_to_paf(d.bam) if with_paf else None
        bam = drop_reads(d.bam, drop_first, seed + 1000 + c) if drop_first > 0 else d.bam
        return d, bam, paf

    threads = threads or min(len(todo), os.cpu_count() or 1)
    with ThreadPoolExecutor(max(1, threads)) as ex:
        parts = list(ex.map(one, todo))
    # global read ids: contig c's reads start at the sum of the earlier contigs' record budgets (an upper bound of
    # their read counts that needs no generation), so a shard sees the same ids as the whole genome
    budget = np.maximum(1, (np.asarray(lengths, np.float64) * coverage /
                            spec_kw.get("read_mean", SynthSpec.read_mean)).astype(np.int64))
    base = np.concatenate([[0], np.cumsum(budget)])
    n_reads = int(base[-1])
    bams, pafs, holes, n_runs, aligned = [], [], [None] * nct, [None] * nct, 0
    for c, (d, bam, paf) in zip(todo, parts):
        assert d.n_reads <= budget[c]
        bam.ref_id[:] = c
        bam.read_id += np.uint32(base[c])
        if paf is not None:
            paf.ref_id[:] = c
            paf.read_id += np.uint32(base[c])
            pafs.append(paf)
        bams.append(bam)
        holes[c], n_runs[c] = d.holes[0], d.n_runs[0]
        aligned += int(bam.ref_len().sum())
    bam = concat_aln(bams)
    paf = None
    if with_paf:
        paf = paf_second_aligner(concat_paf(pafs), lengths, seed + 7)
        aligned += int((paf.tend.astype(np.int64) - paf.tstart).sum())
    return GenomeWorkload(ContigTable(names, lengths), bam, paf, n_reads, holes, n_runs, aligned)
