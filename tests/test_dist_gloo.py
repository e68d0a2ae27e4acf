"""world_size-2 gloo test of the multi-GPU host logic (contig assignment, dealing a read set by contig owner / read
home, genome row)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %r)
    from gci_b200 import dist as D
    rank, world, local = D.init("gloo")
    assert world == 2
    # genome row: each rank owns some contigs
    lens = [np.array([10, 7, 3]), np.array([8, 8])][rank]
    mean, nctg, all_len = D.genome_row([300, 500][rank], [100, 100][rank], [3, 2][rank], lens)
    assert abs(mean - 4.0) < 1e-12 and nctg == 5 and sorted(all_len.tolist()) == [3, 7, 8, 8, 10]
    # dealing a read set: records by contig owner, PAF lines by read home with home-local ids
    from gci_b200 import sharded, synth
    lengths = [50_000, 30_000, 20_000]
    w = synth.make_genome(lengths, ["a", "b", "c"], coverage=10, seed=3, read_mean=3000, read_min=500, read_max=8000)
    plan = sharded.make_plan(rank, world, lengths)
    assert sorted(set(plan.owner)) == [0, 1]
    bam, paf = sharded.shard_bam(w.bam, plan), sharded.shard_paf(w.paf, plan)
    own = np.asarray(plan.owner) == rank
    assert own[bam.ref_id].all() and bam.n_records == int(own[w.bam.ref_id].sum())
    mine = sharded.home_rank(w.paf.read_id, 2) == rank
    assert 0 < int(mine.sum()) < w.paf.n_records
    assert paf.n_records == int(mine.sum()) and np.array_equal(paf.read_id, sharded.home_local(w.paf.read_id[mine], 2))
    counts = D.allreduce(np.array([bam.n_records, paf.n_records], np.int64))
    assert counts.tolist() == [w.bam.n_records, w.paf.n_records]
    assert D.allreduce(np.array([rank + 1.5]), "max").tolist() == [2.5]
    # every rank 'decoded' a part of the files: PAF lines move to the read homes, BAM records to the contig owners
    cols = ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")
    rows = lambda t: sorted(zip(*[np.asarray(getattr(t, c)).astype(np.int64).tolist() for c in cols]))
    src = sharded.shard_paf_source(w.paf, plan, w.bam)
    assert 0 < src.n_records < w.paf.n_records
    assert rows(sharded.deal_paf_over_process_group(src, plan)) == rows(paf)
    part = w.bam.take(np.arange(rank, w.bam.n_records, 2))             # every other record, any contig
    dealt = sharded.deal_bam_over_process_group(part, plan)
    def recs(t):
        off = t.cigar_off.astype(np.int64)
        return sorted((int(t.ref_id[i]), int(t.ref_start[i]), int(t.read_id[i]), int(t.mapq[i]), int(t.flag[i]),
                       int(t.nm[i]), int(t.qlen[i]), t.cigar[off[i]:off[i + 1]].tobytes()) for i in range(t.n_records))
    assert recs(dealt) == recs(bam)
    key = dealt.ref_id.astype(np.int64) * 2**32 + dealt.ref_start.astype(np.int64)
    assert np.all(np.diff(key) >= 0)                                   # coordinate sorted, as the CIGAR kernels expect
    print("rank", rank, "ok")
""")


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sk:               # a port that is free right now (a fixed one may sit in TIME_WAIT)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_assign_contigs_balances():
    from gci_b200.dist import assign_contigs
    lengths = [248, 242, 201, 193, 182, 172, 160, 146, 150, 134, 135, 133, 113, 101, 99, 96, 84, 80, 61, 66, 45, 51, 154, 62]
    owner = assign_contigs(lengths, None, 8)
    load = np.bincount(owner, weights=lengths, minlength=8)
    assert load.max() / load.mean() < 1.08
