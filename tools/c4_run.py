#!/usr/bin/env python
"""BASELINE.json configs[3] shape on N GPUs (one process per GPU, torchrun): CHM13-like 3.1 Gbp genome, HiFi 30x
(BAM + PAF) AND ONT (two BAMs = dual aligner; coverage reduced with --ont-coverage because 60x ONT is 2 x 60 GB of
CIGARs that a numpy generator cannot produce inside a GPU call), contigs dealt to the ranks, read sets exchanged over
NVLink peer memory inside every pipeline call, two-type max, N-masking, scan, score, .depth.gz of all three tracks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/c4_run.py

Prints one JSON line (rank 0): per-stage milliseconds (max over ranks), aligned Gbases/s of the filter -> depth -> scan
part, and size-independent checks on every rank: max track == max(HiFi, Nano) base by base on its smallest contig,
depth sums == sums of the fetched arrays, issue intervals == the oracle's collapse of the fetched array, `.depth.gz`
members inflate to the fetched array.  (The join itself is checked against the C port by tests/test_gpu_multi.py and
tests/test_gpu_shard.py at sizes the port can hold in one process.)
"""
import argparse
import gzip
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200 import dist as D, sharded, synth  # noqa: E402

GATES = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--ont-coverage", type=float, default=8.0)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    rank, world, local = D.init("nccl")
    torch.cuda.set_device(local)
    from gci_b200._lib import Context
    from oracle import c_oracle as CO
    lengths = [max(20000, int(x * a.scale)) for x in synth.CHM13_LENGTHS]
    names = list(synth.CHM13_NAMES)
    plan = sharded.make_plan(rank, world, lengths)
    owned = np.flatnonzero(plan.owned).tolist()
    t0 = time.time()
    hifi = synth.make_genome(lengths, names, coverage=30, seed=401, contig_ids=owned)
    ont = synth.make_genome(lengths, names, coverage=a.ont_coverage, seed=402, contig_ids=owned, with_paf=False,
                            read_mean=30000, read_sigma=0.6, read_min=2000, read_max=150000, events_per_base=0.04)
    view = type("V", (), {"bam": ont.bam, "spec": synth.SynthSpec(lengths, seed=403 + rank), "contigs": ont.contigs})()
    ont2 = synth.second_aligner(view, seed=404 + rank)
    gen_s = time.time() - t0
    # the second ONT aligner may place a read on a contig of another rank: deal those records to the contig owners
    ont2 = sharded.deal_bam_over_process_group(ont2, plan)
    ctx = Context(local)
    n_reads = max(hifi.n_reads, ont.n_reads)
    handle = sharded.configure(ctx, plan, lengths, CO.name_rank(names), n_reads, max_bam_files=2)
    sharded.open_over_process_group(ctx, plan, handle)
    D.init_native_comm(ctx, cap=16384)
    n_runs = hifi.n_runs
    ctx.set_n_runs([c for c in owned for _ in n_runs[c]], [iv[0] for c in owned for iv in n_runs[c]],
                   [iv[1] for c in owned for iv in n_runs[c]])
    paf = sharded.deal_paf_over_process_group(hifi.paf, plan)
    ctx.set_timing(False)
    kw = dict(flank_len=15, lo=-1, hi=0, dist_percent=0.005, **GATES)
    own_len = sum(lengths[c] for c in owned)
    ms = {}

    def timed(name, fn, reps=1):
        torch.cuda.synchronize()
        D.barrier()
        t = time.perf_counter()
        for _ in range(reps):
            out = fn()
        ctx.sync()
        ms[name] = (time.perf_counter() - t) * 1e3 / reps
        return out

    # HiFi: PAF + BAM
    ctx.reads_begin(hifi.n_reads)
    ctx.upload_paf(paf)
    ctx.upload_bam(hifi.bam)
    for _ in range(2):
        ctx.pipeline_row(0, len(owned), own_len, cap=16384, **kw)
    h = timed("hifi_pipeline", lambda: ctx.pipeline_row(0, len(owned), own_len, cap=16384, **kw), a.steps)
    # ONT: two BAMs
    ctx.reads_begin(ont.n_reads)
    ctx.upload_bam(ont.bam)
    ctx.upload_bam(ont2)
    for _ in range(2):
        ctx.pipeline_row(1, len(owned), own_len, cap=16384, **kw)
    o = timed("ont_pipeline", lambda: ctx.pipeline_row(1, len(owned), own_len, cap=16384, **kw), a.steps)
    timed("mask_hifi_nano", lambda: (ctx.mask_gaps(0), ctx.mask_gaps(1)))
    timed("merge_max", lambda: ctx.merge_max(0, 1, 2, -1, 0))
    timed("mask_two_type", lambda: ctx.mask_gaps(2))
    n_iv = timed("scan_two_type", lambda: ctx.scan(2, -1, 0, 15))
    terms = timed("score_two_type", lambda: ctx.score_terms(2, len(owned), n_iv, 0.005, 15, with_sums=True))
    heads = [f">{n}\n".encode() for n in names]
    gz = {t: timed(f"depth_gz_track{t}", lambda t=t: ctx.depth_gzip_track(t, heads)) for t in (0, 1, 2)}
    # ---- checks on this rank's smallest contig ----
    c = min(owned, key=lambda i: lengths[i])
    d0, d1, d2 = (ctx.fetch_depth(t, c).astype(np.int64) for t in (0, 1, 2))
    ok = {"max_track": bool(np.array_equal(d2, np.maximum(d0, d1)))}
    k = owned.index(c)
    ok["depth_sum"] = int(terms[4][k]) == int(d2.sum())
    gs, ge, off = ctx.fetch_intervals(2, len(owned))
    ok["intervals"] = list(zip(gs[off[k]:off[k + 1]].tolist(), ge[off[k]:off[k + 1]].tolist())) == CO.collapse(d2, -1, 0, 15, 0)
    blob, boff = gz[2]
    text = gzip.decompress(blob[boff[c]:boff[c + 1]].tobytes())
    ok["depth_gz"] = text == heads[c] + b"".join(b"%d\n" % v for v in d2.tolist())
    for s, e in n_runs[c]:
        ok["n_runs_masked"] = ok.get("n_runs_masked", True) and not d2[s:e].any()
    aligned = hifi.aligned_bases + int(ont.bam.ref_len().sum()) + int(ont2.ref_len().sum())
    tot_aligned = int(D.allreduce(np.array([aligned], np.int64))[0])
    all_ms = {k_: float(D.allreduce(np.array([v]), "max")[0]) for k_, v in ms.items()}
    all_ok = {k_: bool(int(D.allreduce(np.array([int(v)], np.int64), "sum")[0]) == world) for k_, v in ok.items()}
    surv = D.allreduce(np.array([h[0], o[0]], np.int64))
    if rank == 0:
        path_ms = all_ms["hifi_pipeline"] + all_ms["ont_pipeline"] + all_ms["merge_max"] + all_ms["mask_hifi_nano"] + \
            all_ms["mask_two_type"] + all_ms["scan_two_type"] + all_ms["score_two_type"]
        print(json.dumps({"workload": f"chm13like_{sum(lengths) / 1e9:.2f}Gbp_hifi30x_bam+paf__ont{a.ont_coverage:g}x_2bam",
                          "n_gpus": world, "aligned_bases": tot_aligned, "generate_s": gen_s, "stage_ms_max_over_ranks": all_ms,
                          "path_ms": path_ms, "gbases_per_s": tot_aligned / path_ms / 1e6, "survivors": surv.tolist(),
                          "checks_all_ranks": all_ok, "mean_depth_hifi": h[5], "mean_depth_nano": o[5],
                          "gz_bytes_rank0": {str(t): int(gz[t][0].size) for t in gz},
                          "device_bytes_rank0": ctx.device_bytes}))
    ctx.close()
    D.barrier()


if __name__ == "__main__":
    main()
