"""Contig-sharded filter -> depth for multi-GPU runs.

Each rank owns a set of contigs and receives only the BAM records of those contigs.  Gates and the
last-record-wins dedup are local.  The cross-file join is keyed by READ (a read aligned to different
contigs by two aligners must be dropped, GCI.py:296-297), so the per-file winner tables are exchanged
before the join: every rank then evaluates the join for all reads and accumulates depth only on the
contigs it owns.  PAF files are small (GCI.py:211-254 needs all lines of a read to elect its primary
target), so every rank runs the PAF leg on the whole file.

Two phases so that the exchange can be NCCL (`dist.exchange_file_tables`) or, in tests, a plain merge:

    tables = local_tables(ctx, ...)            # phase 1, per rank
    merged = exchange(tables)                  # all ranks' rows per file
    n_surv = join_and_depth(ctx, merged, ...)  # phase 2, per rank
"""
from __future__ import annotations

import numpy as np

from ._lib import Context, NO_FLAGS


def owned_mask(n_contigs, owner, rank, chrs_selected=None):
    own = np.array([owner[c] == rank for c in range(n_contigs)], dtype=bool)
    if chrs_selected is not None:
        own &= np.asarray(chrs_selected, dtype=bool)
    return own


def local_tables(ctx: Context, lengths, name_rank, selected_all, owned, pafs, bams_local, n_reads, map_qual=30,
                 mq_cutoff=50, iden_percent=0.9, clip_percent=0.1):
    """Phase 1.  `bams_local`: this rank's records (contigs it owns) of every BAM file, in CLI order.
    Returns one table per file in join order (PAFs first): (read_id, contig, start, end, qlen, highq)."""
    # PAF election must see every selected contig; BAM gates only this rank's contigs
    out = []
    if pafs:
        ctx.set_contigs(lengths, selected_all)
        ctx.set_name_rank(name_rank)
        ctx.reads_begin(n_reads)
        for t in pafs:
            ctx.upload_paf(t)
        ctx.filter(map_qual, mq_cutoff, iden_percent, clip_percent, 0.0)
        out += [ctx.fetch_file_table(i) for i in range(len(pafs))]
    ctx.set_contigs(lengths, owned)
    ctx.set_name_rank(name_rank)
    if bams_local:
        ctx.reads_begin(n_reads)
        for t in bams_local:
            ctx.upload_bam(t)
        ctx.filter(map_qual, mq_cutoff, iden_percent, clip_percent, 0.0)
        out += [ctx.fetch_file_table(i) for i in range(len(bams_local))]
    return out


def merge_tables(per_rank_tables):
    """Single-process stand-in for dist.exchange_file_tables: per file, concatenate the ranks' rows; a read
    present on several ranks keeps the row of the highest contig (the reference's fetch order)."""
    n_files = len(per_rank_tables[0])
    out = []
    for f in range(n_files):
        cols = [np.concatenate([rt[f][k] for rt in per_rank_tables]) for k in range(6)]
        r, c = cols[0], cols[1]
        if len(r):
            order = np.lexsort((c, r))
            rs = r[order]
            last = np.ones(len(r), bool)
            last[:-1] = rs[1:] != rs[:-1]
            # high-quality marks are per read: OR over the rows that are dropped too
            hq = np.zeros(int(r.max()) + 1, np.uint8)
            np.maximum.at(hq, r, cols[5])
            keep = order[last]
            cols = [x[keep] for x in cols]
            cols[5] = hq[cols[0]]
        out.append(tuple(cols))
    return out


def join_and_depth(ctx: Context, merged, n_reads, track=0, ovlp_percent=0.9, flank_len=15, lo=NO_FLAGS, hi=NO_FLAGS):
    """Phase 2: join over the exchanged tables, depth on the contigs this context owns."""
    ctx.reads_begin(n_reads)
    for r, c, s, e, q, h in merged:
        ctx.upload_table(r, c, s, e, q, h)
    n_surv = ctx.filter(0, 0, 0.0, 1.0, ovlp_percent)     # tables carry no gates of their own
    ctx.depth(track, flank_len, lo, hi)
    return n_surv
