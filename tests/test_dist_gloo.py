"""world_size-2 gloo test of the multi-GPU host logic (contig assignment, table exchange, genome row)."""
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
    # table exchange: read 5 won on both ranks -> the higher contig wins; read ids stay unique
    if rank == 0:
        t = (np.array([1, 5], np.uint32), np.array([0, 0], np.int32), np.array([10, 20], np.int32),
             np.array([110, 120], np.int32), np.array([100, 100], np.int32), np.array([1, 0], np.uint8))
    else:
        t = (np.array([5, 9], np.uint32), np.array([1, 1], np.int32), np.array([30, 40], np.int32),
             np.array([130, 140], np.int32), np.array([100, 100], np.int32), np.array([0, 1], np.uint8))
    (r, c, s, e, q, h), = D.exchange_file_tables([t])
    assert r.tolist() == [1, 5, 9] and c.tolist() == [0, 1, 1] and s.tolist() == [10, 30, 40], (r, c, s)
    assert D.allreduce(np.array([rank + 1.5]), "max").tolist() == [2.5]
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
