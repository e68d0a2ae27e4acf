#!/usr/bin/env python
"""Large-genome run of the hot path on one GPU (BASELINE.json configs 3-5 territory): timings per stage and
size-independent correctness properties (no oracle at this size).

    python tools/scale_check.py --gbp 3.1 --contigs 24 --coverage 30 [--second-aligner]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200 import synth  # noqa: E402
from gci_b200._lib import Context  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbp", type=float, default=1.0)
    ap.add_argument("--contigs", type=int, default=24)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--second-aligner", action="store_true")
    ap.add_argument("--ont", action="store_true")
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    total = int(a.gbp * 1e9)
    w = np.linspace(2.0, 0.6, a.contigs)
    lengths = [int(x) for x in (w / w.sum() * total)]
    t0 = time.time()
    kw = dict(read_mean=30000, read_sigma=0.6, read_min=2000, read_max=150000, events_per_base=0.04) if a.ont else {}
    parts = []
    # generate contig by contig to bound host memory
    tabs, n_reads, aligned, holes = [], 0, 0, []
    for c, L in enumerate(lengths):
        d = synth.make_reads(synth.SynthSpec([L], coverage=a.coverage, seed=77 + c, **kw))
        t = d.bam
        t.ref_id[:] = c
        t.read_id += n_reads
        n_reads += d.n_reads
        aligned += d.aligned_bases
        holes.append(d.holes[0])
        tabs.append(t)
    from gci_b200.records import AlnTable
    off = np.concatenate([[0]] + [t.cigar_off[1:].astype(np.int64) + sum(x.n_ops for x in tabs[:i])
                                  for i, t in enumerate(tabs)]).astype(np.uint64)
    bam = AlnTable(*[np.concatenate([getattr(t, k) for t in tabs]) for k in
                     ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id")], off,
                   np.concatenate([t.cigar for t in tabs]))
    gen_s = time.time() - t0
    files = [bam]
    if a.second_aligner:
        class _D:  # minimal SynthData view for second_aligner
            pass
        dd = _D()
        dd.bam, dd.spec = bam, synth.SynthSpec(lengths, seed=5)
        dd.contigs = type("C", (), {"lengths": np.asarray(lengths, np.int64)})()
        files.append(synth.second_aligner(dd, seed=11))
        aligned += int(files[1].ref_len().sum())
    out = {"genome_bases": int(sum(lengths)), "contigs": a.contigs, "records": [f.n_records for f in files],
           "cigar_ops": [f.n_ops for f in files], "aligned_bases": aligned, "n_reads": n_reads, "generate_s": gen_s,
           "host_record_bytes": sum(f.nbytes() for f in files)}
    with Context(0) as ctx:
        ctx.set_contigs(lengths)
        t0 = time.time()
        ctx.reads_begin(n_reads)
        for f in files:
            ctx.upload_bam(f)
        out["upload_s"] = time.time() - t0
        for step in range(a.steps + 1):
            if step == 1:
                ctx.stage_reset()
                t0 = time.time()
            n_surv = ctx.filter()
            ctx.depth(0, 15, -1, 0)
            n_iv = ctx.scan(0, -1, 0, 15)
            n50, nctg, lens, loff = ctx.score_terms(0, len(lengths), n_iv)
        ctx.sync()
        dt = (time.time() - t0) / a.steps
        out.update(step_ms=dt * 1e3, gbases_per_s=aligned / dt / 1e9, survivors=n_surv, intervals=n_iv,
                   device_bytes=ctx.device_bytes,
                   stage_ms={k: v[0] / a.steps for k, v in ctx.stage_report().items() if v[1]})
        tiles = sum(l // 1024 + 1 for l in lengths)
        dms = out["stage_ms"]["depth"]
        out["depth_kernel_GBps"] = (4.125 * tiles * 1024 + 4 * n_surv + 16 * tiles) / (dms * 1e-3) / 1e9
        # properties
        r, c, s, e = ctx.fetch_survivors()
        L = np.asarray(lengths, np.int64)[c]
        x = np.clip(s.astype(np.int64) + 15, 0, L)
        y = np.clip(e.astype(np.int64) - 15 + 1, 0, L)
        want_sums = np.bincount(c, weights=np.maximum(0, y - x), minlength=len(lengths)).astype(np.int64)
        sums = ctx.depth_sums(0)
        out["sum_depth_matches_survivor_slices"] = bool(np.array_equal(sums, want_sums))
        gs, ge, goff = ctx.fetch_intervals(0, len(lengths))
        ok = True
        for ci in range(len(lengths)):
            a0, b0 = goff[ci], goff[ci + 1]
            for hs, he in holes[ci]:
                k = np.searchsorted(gs[a0:b0], hs, side="right") - 1
                ok &= bool(k >= 0 and gs[a0 + k] <= hs and ge[a0 + k] >= he)
        out["every_hole_inside_an_issue_interval"] = ok
        # one contig checked base by base against a numpy delta/cumsum of its survivors
        ci = len(lengths) - 1
        m = c == ci
        delta = np.zeros(lengths[ci] + 1, np.int64)
        xs, ys = x[m], y[m]
        np.add.at(delta, xs[ys > xs], 1)
        np.add.at(delta, ys[ys > xs], -1)
        out["last_contig_depth_exact"] = bool(np.array_equal(ctx.fetch_depth_narrow(0, ci).astype(np.int64),
                                                             np.cumsum(delta[:-1])))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
