"""The C restatement (oracle/gci_oracle.c) against the Python oracle (itself pinned to the reference)."""
import numpy as np
import pytest

from oracle import gci_oracle as O, c_oracle as CO
from gci_b200 import synth


@pytest.mark.parametrize("seed,threads", [(0, 1), (1, 4), (2, 3)])
def test_c_hot_path_equals_python_oracle(seed, threads):
    lengths = [150_000, 70_000, 30_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=20, seed=500 + seed, read_mean=7000, read_min=1000,
                                         read_max=20000, hole_fraction=0.02))
    bams = [synth.drop_reads(d.bam, 0.03, seed)] + ([synth.second_aligner(d, seed=seed)] if seed else [])
    want_d, want_s = O.filter_depth([], bams, d.contigs.names, lengths)
    got_d, got_b, n_surv = CO.hot_path(bams, lengths, d.n_reads, threads=threads)
    assert n_surv == len(want_s)
    for i in range(len(lengths)):
        assert np.array_equal(got_d[i], want_d[i])
        assert got_b[i] == O.collapse_depth_range(want_d[i], -1, 0, 15, 0)


def test_c_collapse_equals_loop():
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = int(rng.integers(0, 150))
        fl = int(rng.integers(0, 20))
        d = (rng.random(n) < rng.random()).astype(np.int64) * rng.integers(1, 4, n)
        ts = int(rng.integers(0, 3))
        assert CO.collapse(d, -1, ts, fl, 3) == O.collapse_depth_range_loop(d, -1, ts, fl, 3)
