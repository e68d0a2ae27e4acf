"""One process per GPU (torchrun, NCCL, CUDA IPC): the sharded read-set exchange between real ranks, and the
multi-GPU command line — `torchrun ... GCI.py ...` must write the same .depth.gz / .bed / .gci files as one GPU
(SURVEY.md §4.5).  Needs at least two visible GPUs; skipped otherwise (tests/test_gpu_shard.py covers the exchange
protocol with several contexts on one GPU)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, gzip, json
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch
    from gci_b200 import dist as D, sharded, synth, io as gio
    from gci_b200._lib import Context
    from oracle import c_oracle as CO
    rank, world, local = D.init("nccl")
    tmp = %(tmp)r
    GATES = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)

    # ---- 1. library level: BAM + PAF sharded over the ranks, three steps (eager, captured, replayed) ----
    lengths = [x // 60 for x in synth.CHM13_LENGTHS]
    names = list(synth.CHM13_NAMES)
    w = synth.make_genome(lengths, names, coverage=25, seed=11)
    want_d, want_b, want_n = CO.hot_path([w.bam], lengths, w.n_reads, threads=4, pafs=[w.paf], names=names, **GATES)
    plan = sharded.make_plan(rank, world, lengths)
    ctx = Context(local)
    handle = sharded.configure(ctx, plan, lengths, CO.name_rank(names), w.n_reads, max_bam_files=1)
    sharded.open_over_process_group(ctx, plan, handle)
    D.init_native_comm(ctx)
    ctx.set_timing(False)
    ctx.reads_begin(w.n_reads)
    ctx.upload_paf(sharded.deal_paf_over_process_group(sharded.shard_paf_source(w.paf, plan, w.bam), plan))
    ctx.upload_bam(sharded.shard_bam(w.bam, plan))
    owned = np.flatnonzero(plan.owned)
    for _ in range(3):
        out = ctx.pipeline_row(0, len(owned), sum(lengths[c] for c in owned.tolist()), flank_len=15, lo=-1, hi=0, dist_percent=0.005, **GATES)
    n_surv = int(D.allreduce(np.array([out[0]], np.int64))[0])
    assert n_surv == want_n, (n_surv, want_n)
    gs, ge, off = ctx.fetch_intervals(0, len(owned))
    for o, c in enumerate(owned.tolist()):
        assert np.array_equal(ctx.fetch_depth(0, c).astype(np.int64), want_d[c]), ("depth", rank, c)
        assert list(zip(gs[off[o]:off[o + 1]].tolist(), ge[off[o]:off[o + 1]].tolist())) == want_b[c], ("bed", rank, c)
    mean = out[5]
    assert abs(mean - sum(int(d.sum()) for d in want_d) / sum(lengths)) < 1e-9
    assert ctx.graph_replays >= 2
    ctx.close()
    print("rank", rank, "library ok")

    # ---- 2. command line: the same files from N GPUs as from one ----
    from gci_b200 import pipeline as P
    small = [x // 200 for x in synth.CHM13_LENGTHS[:8]]
    snames = names[:8]
    h = synth.make_genome(small, snames, coverage=20, seed=21)
    o = synth.make_genome(small, snames, coverage=25, seed=22, with_paf=False, read_mean=20000, read_sigma=0.5,
                          read_min=2000, read_max=80000, events_per_base=0.03)
    files = {k: os.path.join(tmp, k) for k in ("hifi.bam", "hifi.paf", "ont.bam", "ref.fa", "regions.bed")}
    if rank == 0:
        gio.write_bam(files["hifi.bam"], snames, small, h.bam)
        gio.write_paf(files["hifi.paf"], h.paf, snames, small)
        gio.write_bam(files["ont.bam"], snames, small, o.bam)
        gio.write_fasta(files["ref.fa"], snames, small, h.n_runs)
        with open(files["regions.bed"], "w") as f:
            f.write(f"{snames[0]}\\t1000\\t90000\\n{snames[5]}\\t500\\t40000\\n{snames[2]}\\t0\\t{small[2]}\\n")
    D.barrier()
    args = dict(hifi=[files["hifi.bam"], files["hifi.paf"]], nano=[files["ont.bam"]], reference=files["ref.fa"],
                regions=files["regions.bed"], prefix="T", force=True, threads=4)
    P.set_default_session(P.Session(local))
    P.GCI(directory=os.path.join(tmp, "multi"), **args)
    D.barrier()
    if rank == 0:
        os.environ["WORLD_SIZE"] = "1"                 # the same driver on one GPU
        P.set_default_session(P.Session(local))
        P.GCI(directory=os.path.join(tmp, "single"), **args)
        os.environ["WORLD_SIZE"] = str(world)
        a, b = os.path.join(tmp, "multi"), os.path.join(tmp, "single")
        assert sorted(os.listdir(a)) == sorted(os.listdir(b)), (os.listdir(a), os.listdir(b))
        for fn in sorted(os.listdir(b)):
            x, y = open(os.path.join(a, fn), "rb").read(), open(os.path.join(b, fn), "rb").read()
            if fn.endswith(".gz"):
                x, y = gzip.decompress(x), gzip.decompress(y)
            assert x == y, fn
        print("files identical:", sorted(os.listdir(b)))
    D.barrier()
    print("rank", rank, "cli ok")
""")


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_ranks_over_processes(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, tmp=str(tmp_path)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert r.stdout.count("library ok") == world and r.stdout.count("cli ok") == world
    assert "files identical" in r.stdout
