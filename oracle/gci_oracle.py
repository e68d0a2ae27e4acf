"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A from-scratch CPU restatement (numpy + small Python loops) of the reference's
filter -> depth -> gap-scan -> score path, operating on the same columnar record
tables the CUDA path consumes.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module; the
product (`gci_b200/`) never does.

Pinning (see DESIGN.md "Oracle"):
  * depth -> BED -> .gci (L4+L5): pinned by the reference's own example
    (`example/MH63.{depth.gz,0.depth.bed,gci}`), committed as
    tests/golden/mh63_*.  [pinned]
  * BAM/PAF -> depth (L2+L3): the reference holds no fixture; pinned instead by
    outputs of the UNMODIFIED reference `filter()` run in the build container
    through a pysam-surface shim (tests/golden/make_golden.py ->
    tests/golden/filter_cases.npz) and by the known-answer vectors of
    SURVEY.md §4.4.  The arithmetic that lives in pysam/htslib (not vendored,
    no version pin — README.md:32-33) is restated from the SAM/BAM spec:
    `get_cigar_stats` = per-op base counts, `reference_end` = htslib
    bam_endpos (pos + max(1, ref-consuming length)), `query_length` = l_seq.

Every function cites the reference lines it follows (paths into the reference
repo, e.g. GCI.py:146-169).
"""
from __future__ import annotations

from math import log2

import numpy as np

# op codes MIDNSHP=XB
_M, _I, _D, _N, _S, _H, _P, _EQ, _X, _B = range(10)
NM_MISSING = -(2**31)


class ReferenceWouldRaise(Exception):
    """The reference raises here (ZeroDivisionError / KeyError) — GCI.py:163, :165, :231, :292."""


# --------------------------------------------------------------------------------------
# L2: BAM leg  (GCI.py:146-169 read_sam; pysam get_cigar_stats / reference_end)
# --------------------------------------------------------------------------------------

def cigar_stats(ops: np.ndarray) -> np.ndarray:
    """pysam `get_cigar_stats()[0][:10]`: bases per op code (uint32 sums)."""
    out = np.zeros(10, dtype=np.int64)
    if len(ops):
        np.add.at(out, (ops & 15).astype(np.int64), (ops >> 4).astype(np.int64))
    return out


def cigar_stats_all(cigar: np.ndarray, cigar_off: np.ndarray) -> np.ndarray:
    """[A,10] int64 per-record op base counts (vectorised `cigar_stats`)."""
    a = len(cigar_off) - 1
    out = np.zeros((a, 10), dtype=np.int64)
    if len(cigar) == 0:
        return out
    op = (cigar & 15).astype(np.int64)
    ln = (cigar >> 4).astype(np.int64)
    off = cigar_off.astype(np.int64)
    rec = np.repeat(np.arange(a, dtype=np.int64), off[1:] - off[:-1])
    np.add.at(out, (rec, op), ln)
    return out


def bam_leg(tab, selected: np.ndarray, map_qual: int, clip_percent: float, iden_percent: float,
            mq_cutoff: int):
    """GCI.py:146-169 + the fan-out merge :257-270 at `-t 1`.

    `selected[ref_id]` is True for contigs in `targets_length` (GCI.py:202-207);
    records elsewhere are never fetched (:151).  Returns
    (dict read_id -> (ref_id, start, end, qlen), set of high-quality read_ids).
    Records are visited contig by contig in header order, file order inside a
    contig; a later record of the same read overwrites an earlier one (:166).
    """
    st = cigar_stats_all(tab.cigar, tab.cigar_off)
    out: dict = {}
    highq: set = set()
    a = tab.n_records
    ref_id = tab.ref_id
    order = np.arange(a)
    if a:
        # fetch order: contigs in header order, file order inside (stable)
        order = np.argsort(ref_id, kind="stable")
    for i in order:
        rid = int(ref_id[i])
        if rid < 0 or rid >= len(selected) or not selected[rid]:
            continue
        fl = int(tab.flag[i])
        if (fl & 0x4) or (fl & 0x100) or (fl & 0x800):           # :153-156
            continue
        mq = int(tab.mapq[i])
        if mq < map_qual:                                        # :156
            continue
        M, I, D, N, S, H, P, EQ, X, B = (int(v) for v in st[i])  # :157-162
        nm = int(tab.nm[i])
        if nm == NM_MISSING:
            raise ReferenceWouldRaise("record without NM tag (KeyError at GCI.py:163)")
        mm = nm - (I + D)                                        # :164
        d1 = M + EQ + X + I + S
        if d1 == 0:
            raise ReferenceWouldRaise("ZeroDivisionError at GCI.py:165 (clip ratio)")
        if not (S / d1 <= clip_percent):
            continue
        d2 = M + EQ + X + I + D
        if d2 == 0:
            raise ReferenceWouldRaise("ZeroDivisionError at GCI.py:165 (identity)")
        if not ((M + EQ + X - mm) / d2 >= iden_percent):
            continue
        rlen = M + D + N + EQ + X                                # htslib bam_cigar2rlen
        if rlen == 0:
            rlen = 1                                             # htslib bam_endpos
        start = int(tab.ref_start[i])
        q = int(tab.read_id[i])
        out[q] = (rid, start, start + rlen, int(tab.qlen[i]))    # :166
        if mq >= mq_cutoff:                                      # :167-168
            highq.add(q)
    return out, highq


# --------------------------------------------------------------------------------------
# L2: PAF leg  (GCI.py:49-96, :211-254)
# --------------------------------------------------------------------------------------

def _merge_blocks(pairs):
    """GCI.py:64-96: sort [lo,hi] pairs, merge touching/overlapping, return
    (sum of merged lengths, lo, hi of the longest merged block — first one on ties)."""
    pairs = sorted(pairs)
    blocks = []
    lo, hi = pairs[0]
    for l, h in pairs:
        if hi >= l:
            if hi < h:
                hi = h
        else:
            blocks.append((hi - lo, lo, hi))
            lo, hi = l, h
    blocks.append((hi - lo, lo, hi))
    total = sum(b[0] for b in blocks)
    best = blocks[0]
    for b in blocks[1:]:
        if b[0] > best[0]:
            best = b
    return total, best[1], best[2]


def paf_leg(paf_tabs, selected: np.ndarray, contig_names, map_qual: int, iden_percent: float,
            mq_cutoff: int):
    """GCI.py:211-254.  Returns (list of dict read_id -> (ref_id,start,end,qlen), highq set).

    Quirk kept: `synteny` is created once, outside the per-file loop (:214 vs
    :215), so reads collected from earlier PAFs are re-emitted for later ones.
    """
    highq: set = set()
    outs = []
    synteny: dict = {}
    for tab in paf_tabs:
        for i in range(tab.n_records):
            t = int(tab.ref_id[i])
            if t < 0 or t >= len(selected) or not selected[t]:    # :220
                continue
            alnlen = int(tab.alnlen[i])
            if alnlen == 0:
                raise ReferenceWouldRaise("ZeroDivisionError at GCI.py:231")
            identity = int(tab.nmatch[i]) / alnlen                # :231
            mq = int(tab.mapq[i])
            if mq >= map_qual and identity >= iden_percent:       # :232
                q = int(tab.read_id[i])
                synteny.setdefault(q, {}).setdefault(t, []).append(
                    (int(tab.qlen[i]), int(tab.qstart[i]), int(tab.qend[i]),
                     int(tab.tstart[i]), int(tab.tend[i]), identity))
                if mq >= mq_cutoff:                               # :238
                    highq.add(q)
        out = {}
        for q, per_t in synteny.items():                          # :241-254
            best = None
            for t, alns in per_t.items():
                mapped, _, _ = _merge_blocks([(a[1], a[2]) for a in alns])
                qlen = alns[0][0]
                if qlen == 0:
                    raise ReferenceWouldRaise("ZeroDivisionError at GCI.py:247")
                rate = mapped / qlen
                s = 0
                for a in alns:                                    # sum() starts from int 0, file order
                    s = s + a[5]
                score = (s / len(alns)) * rate
                _, ts, te = _merge_blocks([(a[3], a[4]) for a in alns])
                key = (score, contig_names[t])
                if best is None or key > best[0]:
                    best = (key, (t, ts, te, qlen))
            out[q] = best[1]
        outs.append(out)
    return outs, highq


# --------------------------------------------------------------------------------------
# L2: cross-file join  (GCI.py:272-301)
# --------------------------------------------------------------------------------------

def join(files, highq, ovlp_percent: float):
    """files: list of dict read -> (ref_id, start, end, qlen), PAFs first (:272).
    Returns dict read -> (ref_id, start, end[, qlen])."""
    if len(files) <= 1:                                           # :300-301
        return dict(files[0]) if files else {}
    comm = set(files[0])
    for f in files[1:]:
        comm &= set(f)                                            # :277
    final = set(highq) | comm                                     # :279
    cur = {q: seg for q, seg in files[0].items() if q in final}   # :280
    for f in files[1:]:
        for q, seg in f.items():
            if q in cur:
                seg1 = cur[q]
                if seg[0] == seg1[0]:
                    ovlp = min(seg[2], seg1[2]) - max(seg[1], seg1[1])
                    if seg[3] == 0:
                        raise ReferenceWouldRaise("ZeroDivisionError at GCI.py:292")
                    if ovlp / seg[3] < ovlp_percent:              # :292 (later file's qlen)
                        del cur[q]
                    else:
                        cur[q] = (seg1[0], max(seg[1], seg1[1]), min(seg[2], seg1[2]))
                else:
                    del cur[q]
            elif q in highq:                                      # :298-299
                cur[q] = (seg[0], seg[1], seg[2])
    return cur


# --------------------------------------------------------------------------------------
# L3: depth  (GCI.py:302-306, :315-329, :332-353)
# --------------------------------------------------------------------------------------

def accumulate_depth(survivors, lengths, flank_len: int):
    """GCI.py:201-208 + :302-306.  Python slice semantics are kept on purpose
    (negative stop wraps)."""
    depths = [np.zeros(int(l), dtype=np.int64) for l in lengths]
    for seg in survivors.values():
        start = seg[1] + flank_len
        end = seg[2] - flank_len
        depths[seg[0]][start:end + 1] += 1
    return depths


def mask_gaps(depths, n_runs):
    """GCI.py:315-329.  n_runs: list (per contig) of [(start,end), ...] or None."""
    if n_runs is not None:
        for c, segs in enumerate(n_runs):
            for s, e in segs or ():
                depths[c][s:e] = 0
    return depths


def merge_two_types(hifi, nano):
    """GCI.py:350 — element-wise max."""
    return [np.maximum(h, n) for h, n in zip(hifi, nano)]


def filter_depth(paf_tabs, bam_tabs, contig_names, lengths, selected=None, map_qual=30, mq_cutoff=50,
                 iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9, flank_len=15):
    """Whole `filter()` (GCI.py:172-312) minus file writing: returns (depths list, survivors dict)."""
    n = len(lengths)
    if selected is None:
        selected = np.ones(n, dtype=bool)
    paf_files, highq = paf_leg(paf_tabs, selected, contig_names, map_qual, iden_percent, mq_cutoff)
    bam_files = []
    for tab in bam_tabs:
        d, hq = bam_leg(tab, selected, map_qual, clip_percent, iden_percent, mq_cutoff)
        bam_files.append(d)
        highq |= hq
    surv = join(paf_files + bam_files, highq, ovlp_percent)
    lens_sel = [int(l) if selected[i] else 0 for i, l in enumerate(lengths)]
    depths = accumulate_depth(surv, lens_sel, flank_len)
    return depths, surv


# --------------------------------------------------------------------------------------
# L4: gap scan  (GCI.py:356-390)
# --------------------------------------------------------------------------------------

def collapse_depth_range_loop(depth, leftmost=-1, rightmost=0, flank_len=15, start_pos=0):
    """Per-base state machine — literal semantics of GCI.py:371-389 for ONE contig (small inputs)."""
    out = []
    n_all = len(depth)
    view = depth[flank_len:n_all - flank_len]
    in_run = False
    run_start = 0
    last = n_all - 2 * flank_len - 1
    for i, d in enumerate(view):
        if leftmost < d <= rightmost:
            if not in_run:
                run_start = i + flank_len
                in_run = True
            if i == last:
                out.append((run_start + start_pos, i + flank_len + 1 + start_pos))
        elif in_run:
            if i > flank_len:
                out.append((run_start + start_pos, i + flank_len + start_pos))
            in_run = False
    return out


def collapse_depth_range(depth, leftmost=-1, rightmost=0, flank_len=15, start_pos=0):
    """Vectorised equivalent of `collapse_depth_range_loop` for ONE contig.

    Maximal runs of `leftmost < depth <= rightmost` inside the view
    depth[fl : L-fl] (Python slice); a run ending before the view end is kept
    only if its (view-relative) end index is > fl (:385); a run touching the
    view end is always kept (:380-382)."""
    n_all = len(depth)
    view = np.asarray(depth[flank_len:n_all - flank_len])
    n = len(view)
    if n == 0:
        return []
    f = (view > leftmost) & (view <= rightmost)
    if n_all - 2 * flank_len != n:
        # the `i == chr_len - 2*fl - 1` test (:380) can only fire when the slice was not clamped
        last_fires = False
    else:
        last_fires = True
    pad = np.concatenate([[False], f, [False]])
    d = np.diff(pad.astype(np.int8))
    starts = np.flatnonzero(d == 1)           # view index of first True
    ends = np.flatnonzero(d == -1)            # view index one past last True
    out = []
    for s, e in zip(starts.tolist(), ends.tolist()):
        if e == n:
            if last_fires:
                out.append((s + flank_len + start_pos, e + flank_len + start_pos))
        elif e > flank_len:
            out.append((s + flank_len + start_pos, e + flank_len + start_pos))
    return out


# --------------------------------------------------------------------------------------
# L5: score  (GCI.py:422-519, :522-607)
# --------------------------------------------------------------------------------------

def complement_lengths(bed, length, flank_len=15, start=None, end=None):
    """GCI.py:422-462 for ONE contig: lengths of the stretches between issue intervals."""
    if start is None or end is None:
        start, end = flank_len, length - flank_len
    out = []
    n = len(bed)
    if n == 0:
        return [end - start]
    last = start
    for i, (s, e) in enumerate(bed):
        if s > last:
            out.append(s - last)
        if i != n - 1:
            last = e
        elif end > e:
            out.append(end - e)
    return out


def n50(lengths):
    """GCI.py:465-480."""
    if len(lengths) == 0:
        return 0
    srt = sorted(lengths, reverse=True)
    cum = np.cumsum(srt)
    half = cum[-1] / 2
    for i, c in enumerate(cum):
        if c >= half:
            return srt[i]
    return 0


def merge_close(bed, length, dist_percent=0.005, flank_len=15, start=None, end=None):
    """GCI.py:483-519 for ONE contig."""
    if start is None or end is None:
        start, end = flank_len, length - flank_len
    dist = length * dist_percent
    out = []
    cur = (start, start)
    for seg in bed:
        if (seg[0] - cur[1]) <= dist:
            cur = (cur[0], seg[1])
        else:
            out.append(cur)
            cur = tuple(seg)
    if (end - cur[1]) <= dist:
        cur = (cur[0], end)
    out.append(cur)
    return out


def gci_value(obs_n50, exp_n50, obs_ctg, exp_ctg):
    """GCI.py:601-604."""
    if obs_ctg == 0:
        return 0
    return round(100 * log2(obs_n50 / exp_n50 + 1) / log2(obs_ctg / exp_ctg + 1), 4)


def score_rows(names, lengths, beds, flank_len=15, dist_percent=0.005, chrs_given=False):
    """Rows of one `.gci` block (GCI.py:554-605): list of (name, exp_n50, obs_n50, exp_ctg, obs_ctg, gci)."""
    rows = []
    all_obs = []
    all_new = 0
    for name, length, bed in zip(names, lengths, beds):
        length = int(length)
        obs = complement_lengths(bed, length, flank_len)
        merged = merge_close(bed, length, dist_percent, flank_len)
        new_obs = complement_lengths(merged, length, flank_len)
        all_obs += obs
        all_new += len(new_obs)
        o50 = n50(obs)
        rows.append((name, length, o50, 1, len(new_obs), gci_value(o50, length, len(new_obs), 1)))
    g_exp = n50([int(l) for l in lengths])
    g_obs = n50(all_obs)
    label = "All_chromosomes" if chrs_given else "Genome"
    rows.append((label, g_exp, g_obs, len(lengths), all_new, gci_value(g_obs, g_exp, all_new, len(lengths))))
    return rows


GCI_HEADER = ("Chromosome\tTheoretical maximum N50\tCurated N50\tTheoretical minimum contigs number\t"
              "Curated contigs number\tGCI score\n")
GCI_RULE = "-" * 136 + "\n\n\n"


def gci_text(type_label, rows):
    """One block of the `.gci` file (GCI.py:593-606)."""
    s = f"{type_label}:\n" + GCI_HEADER
    for r in rows:
        s += "\t".join(str(int(v)) if not isinstance(v, (float, str)) else str(v) for v in r) + "\n"
    return s + GCI_RULE


def bed_text(names, beds):
    """`.{ts}.depth.bed` (GCI.py:414-417)."""
    return "".join(f"{n}\t{s}\t{e}\n" for n, bed in zip(names, beds) for s, e in bed)


def depth_text(names, depths):
    """Decompressed `.depth.gz` stream (GCI.py:110-117)."""
    parts = []
    for n, d in zip(names, depths):
        parts.append(f">{n}\n")
        parts.append("".join(f"{int(v)}\n" for v in d))
    return "".join(parts)


def region_scores(depths_by_type, contig_index, regions, threshold=0, dist_percent=0.005):
    """`-R` regions variant (GCI.py:610-657).  regions: list of (contig_name, start, end) in file order,
    grouped by contig in first-appearance order like the reference's dict.  Returns
    (per-region rows [(name,start,end,[gci per type])], all_regions row [gci per type])."""
    grouped = {}
    for name, s, e in regions:
        grouped.setdefault(name, []).append((s, e))
    n_types = len(depths_by_type)
    all_len = []
    all_obs = [[] for _ in range(n_types)]
    all_ctg = [0] * n_types
    rows = []
    for name, segs in grouped.items():
        c = contig_index[name]
        for s, e in segs:
            exp = e - s
            if exp > 0:
                all_len.append(exp)
            vals = []
            for t in range(n_types):
                sub = depths_by_type[t][c][s:e]
                bed = collapse_depth_range(sub, -1, threshold, 0, s)
                obs = complement_lengths(bed, exp, s, s, e)
                o50 = n50(obs)
                if exp > 0:
                    all_obs[t] += obs
                merged = merge_close(bed, exp, dist_percent, s, s, e)
                nctg = len(complement_lengths(merged, exp, s, s, e))
                if exp > 0:
                    all_ctg[t] += nctg
                vals.append(gci_value(o50, exp, nctg, 1))
            rows.append((name, s, e, vals))
    e50 = n50(all_len)
    tot = [gci_value(n50(all_obs[t]), e50, all_ctg[t], len(all_len)) for t in range(n_types)]
    return rows, tot


# --------------------------------------------------------------------------------------
# L1: the driver's three branches (GCI.py:991-1026), outputs as in-memory "files"
# --------------------------------------------------------------------------------------

def run_gci(names, lengths, hifi=None, nano=None, n_runs=None, chrs=None, regions=None, prefix="GCI",
            map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9, flank_len=15,
            threshold=0, dist_percent=0.005):
    """hifi / nano: None or list of AlnTable / PafTable in CLI order.  n_runs: per-contig N-run lists.
    Returns {filename: text | {contig: depth array}} for every file the reference would write."""
    chrs_list = chrs.strip().split(",") if chrs else []
    selected = np.array([(n in chrs_list) if chrs_list else True for n in names], dtype=bool)
    sel_idx = [i for i in range(len(names)) if selected[i]]
    sel_names = [names[i] for i in sel_idx]
    sel_lengths = [int(lengths[i]) for i in sel_idx]
    out = {}
    if n_runs is not None and any(len(r or ()) for r in n_runs):
        out[f"{prefix}.gaps.bed"] = bed_text(names, [r or [] for r in n_runs])          # GCI.py:37-44
    else:
        n_runs = None
    kw = dict(map_qual=map_qual, mq_cutoff=mq_cutoff, iden_percent=iden_percent, clip_percent=clip_percent,
              ovlp_percent=ovlp_percent, flank_len=flank_len)

    def one_type(tabs, pfx):
        pafs = [t for t in tabs if getattr(t, "kind", "bam") == "paf"]
        bams = [t for t in tabs if getattr(t, "kind", "bam") == "bam"]
        depths, _ = filter_depth(pafs, bams, names, lengths, selected, **kw)
        out[f"{pfx}.depth.gz"] = {names[i]: depths[i].copy() for i in sel_idx}         # written before masking
        return mask_gaps(depths, n_runs)

    def bed_of(depths, pfx):
        beds = [collapse_depth_range(depths[i], -1, threshold, flank_len, 0) for i in sel_idx]
        out[f"{pfx}.{threshold}.depth.bed"] = bed_text(sel_names, beds)
        return beds

    tracks, labels, beds = [], [], []
    if nano is None:
        d = one_type(hifi, prefix)
        tracks, labels, beds = [d], ["HiFi"], [bed_of(d, prefix)]
    elif hifi is None:
        d = one_type(nano, prefix)
        tracks, labels, beds = [d], ["Nano"], [bed_of(d, prefix)]
    else:
        dh = one_type(hifi, prefix + "_hifi")
        dn = one_type(nano, prefix + "_nano")
        dm = merge_two_types(dh, dn)
        out[f"{prefix}_two_type.depth.gz"] = {names[i]: dm[i].copy() for i in sel_idx}
        dm = mask_gaps(dm, n_runs)
        tracks, labels = [dh, dn, dm], ["HiFi", "Nano", "HiFi + Nano"]
        beds = [bed_of(dh, prefix + "_hifi"), bed_of(dn, prefix + "_nano"), bed_of(dm, prefix + "_two_type")]
    txt = ""
    for label, b in zip(labels, beds):
        txt += gci_text(label, score_rows(sel_names, sel_lengths, b, flank_len, dist_percent, bool(chrs_list)))
    out[f"{prefix}.gci"] = txt
    if regions:
        rows, tot = region_scores(tracks, {n: i for i, n in enumerate(names)}, regions, threshold, dist_percent)
        t = "Chromosome\tStart\tEnd\t" + "\t".join(labels) + "\n"
        for name, s, e, vals in rows:
            t += f"{name}\t{s}\t{e}\t" + "\t".join(map(str, vals)) + "\n"
        t += GCI_RULE + "All_regions\t*\t*\t" + "\t".join(map(str, tot)) + "\n"
        out[f"{prefix}.regions.gci"] = t
    return out
