#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200 import synth  # noqa: E402
from gci_b200._lib import Context  # noqa: E402

lengths = [40_000, 9_000, 1_500]
d = synth.make_reads(synth.SynthSpec(lengths, coverage=12, seed=5, read_mean=3000, read_min=400, read_max=9000,
                                     hole_fraction=0.05, hole_mean=300))
ont = synth.make_reads(synth.SynthSpec(lengths, coverage=6, seed=6, read_mean=12000, read_sigma=0.6, read_min=2000,
                                       read_max=30000, events_per_base=0.05))
b2 = synth.second_aligner(d, seed=2)
paf = synth.aln_to_paf(synth.second_aligner(d, seed=3))
order = sorted(range(3), key=lambda i: d.contigs.names[i])
rank = np.empty(3, np.int32)
rank[order] = np.arange(3)
with Context(0) as ctx:
    ctx.set_contigs(lengths)
    ctx.set_name_rank(rank)
    ctx.set_n_runs([0, 1], [100, 50], [400, 90])
    ctx.reads_begin(d.n_reads)
    ctx.upload_paf(paf)
    ctx.upload_bam(d.bam)
    ctx.upload_bam(b2)
    n = ctx.filter()
    ctx.depth(0, 15, -1, 0)
    ctx.mask_gaps(0)
    ctx.reads_begin(ont.n_reads)
    ctx.upload_bam(ont.bam)
    ctx.filter()
    ctx.depth(1, 15)
    ctx.mask_gaps(1)
    ctx.merge_max(0, 1, 2, -1, 0)
    for t in (0, 1, 2):
        k = ctx.scan(t, -1, 0, 15)
        ctx.fetch_intervals(t, 3)
        ctx.score_terms(t, 3, k, with_sums=(t != 1) or True)
    ctx.scan_windows(2, [0, 0, 2], [0, 500, 10], [40_000, 9_000, 1_400], -1, 0)
    ctx.depth_sums(1)
    ctx.fetch_depth_narrow(2, 0)
    ctx.depth_text(2, 1)
    ctx.depth_gzip(2, 0, header=b">c\n")
    ctx.fetch_survivors()
    ctx.fetch_file_table(0)
    ctx.reads_begin(d.n_reads)
    ctx.upload_bam(d.bam)
    ctx.pipeline(0, 3)
    ctx.fetch_cigar_stats(0, d.bam.n_records)
    ctx.set_timing(False)               # third call replays the step as a CUDA graph
    for _ in range(3):
        ctx.pipeline(0, 3)
    ctx.reads_begin(ont.n_reads)
    ctx.upload_bam(ont.bam)
    ctx.pipeline(0, 3)
    ctx.fetch_cigar_stats(0, ont.bam.n_records)
    print("sanitize run ok", n, "graph replays", ctx.graph_replays)
