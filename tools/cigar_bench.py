#!/usr/bin/env python
"""CIGAR statistics stage alone (cigar_stats_kernel / cigar_stream_kernel) on a synthetic op stream.

    python tools/cigar_bench.py --records 100000 --mean-ops 2400      # ONT-like: ~1 GB of packed ops
    python tools/cigar_bench.py --records 6000000 --mean-ops 31       # HiFi-like

Prints one JSON line: stage milliseconds per filter run (CUDA events around the memset + kernels of the stage),
GB/s over the 4 B/op stream, and an exact check of every record's sums against numpy."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200._lib import Context  # noqa: E402
from gci_b200.records import AlnTable  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000)
    ap.add_argument("--mean-ops", type=float, default=2400.0)
    ap.add_argument("--sigma", type=float, default=0.6)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--clip-frac", type=float, default=0.3, help="share of records that open with a soft clip")
    ap.add_argument("--hard-frac", type=float, default=0.02, help="share of records that open with a hard clip")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    mu = np.log(a.mean_ops) - a.sigma ** 2 / 2
    n_ops = np.maximum(1, rng.lognormal(mu, a.sigma, a.records)).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(n_ops)]).astype(np.uint64)
    c = int(off[-1])
    op = rng.choice(np.array([0, 0, 0, 0, 0, 1, 2, 7, 8], np.uint32), c)
    ln = rng.integers(1, 40, c, dtype=np.uint32)
    starts = off[:-1].astype(np.int64)
    u = rng.random(a.records)
    op[starts[u < a.clip_frac]] = 4                        # soft clip at the start of the read
    op[starts[u > 1.0 - a.hard_frac]] = 5                  # hard clip (supplementary-style record)
    cigar = (ln << np.uint32(4)) | op
    n = a.records
    tab = AlnTable(np.zeros(n, np.int32), np.zeros(n, np.int32), np.full(n, 60, np.uint8), np.full(n, 4, np.uint16),
                   np.zeros(n, np.int32), np.full(n, 100, np.int32), np.arange(n, dtype=np.uint32), off, cigar)
    # exact sums per record and class with cumulative sums (int64)
    want = np.zeros((n, 5), np.int64)
    o = off.astype(np.int64)
    for k, codes in enumerate(((0, 7, 8), (1,), (2,), (3,), (4,))):
        w = np.where(np.isin(op, codes), ln, 0).astype(np.int64)
        cs = np.concatenate([[0], np.cumsum(w)])
        want[:, k] = cs[o[1:]] - cs[o[:-1]]
    with Context(0) as ctx:
        ctx.set_contigs([2_000_000_000])
        ctx.reads_begin(n)
        ctx.upload_bam(tab)
        for step in range(a.steps + 2):
            if step == 2:
                ctx.stage_reset()
            assert ctx.filter() == 0
        ms, k = ctx.stage_ms("cigar")
        got, _ = ctx.fetch_cigar_stats(0, n)
        ok = bool(np.array_equal(got.astype(np.int64), want))
    ms /= max(1, k)
    print(json.dumps({"records": n, "cigar_ops": c, "mean_ops_per_record": c / n, "stream_bytes": 4 * c,
                      "stage_ms": ms, "stream_GBps": 4 * c / (ms * 1e-3) / 1e9,
                      "with_stats_GBps": (4 * c + 72 * n) / (ms * 1e-3) / 1e9,
                      "note": "stage = memset of the 32 B/record statistics + the CIGAR kernel(s); with_stats adds the "
                              "statistics rows (32 B zeroed + 32 B written) and 8 B of offsets per record",
                      "exact_vs_numpy": ok}))
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
