"""Plot feed on the GPU (gci_sliding_window, GCI.py:660-705): the device points, finished on the host exactly like
gci_b200.plotfeed does, against the 24 known answers harvested from the unmodified reference and against the host
implementation on random tracks."""
import contextlib
import io
import json
import os

import numpy as np
import pytest

from gci_b200 import plotfeed

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    from gci_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _same(a, b):
    (pa, va), (pb, vb) = a, b
    assert pa == pb
    assert va.dtype == vb.dtype and va.shape == vb.shape
    assert np.array_equal(va, vb)


def test_known_answers_of_the_reference(ctx):
    with open(os.path.join(GOLDEN, "sliding_window_kat.json")) as f:
        kats = json.load(f)["cases"]
    assert len(kats) >= 20
    for k in kats:
        d = np.asarray(k["depths"], np.int64)
        ctx.set_contigs([max(1, len(d))])
        if len(d):
            ctx.load_depth(0, 0, d.astype(np.int32))
        else:
            continue                                      # an empty region never reaches the device
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            pos, val = plotfeed.sliding_window_average_depth_gpu(ctx, 0, 0, 0, len(d), k["window_size"], k["max_depth"],
                                                                 k["start"], k["target"])
        assert pos == k["positions"], k
        assert val.tolist() == k["values"] and str(val.dtype) == k["dtype"], k
        assert err.getvalue() == k["stderr"]


@pytest.mark.parametrize("seed", range(4))
def test_random_tracks_equal_the_host_function(ctx, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50_000, 400_000))
    d = np.repeat(rng.poisson(8, n // 50 + 1), 50)[:n].astype(np.int64)
    for _ in range(int(rng.integers(0, 30))):              # zero stretches of every length, also at both ends
        a = int(rng.integers(0, n))
        d[a:a + int(rng.integers(1, 3000))] = 0
    if seed == 1:
        d[:10] = 0; d[-1] = 0
    if seed == 2:
        d[:] = np.maximum(d, 1)                            # no zero at all
    ctx.set_contigs([n, 5])
    ctx.load_depth(0, 0, d.astype(np.int32))
    for ws, lo, hi in ((50_000, 0, n), (997, 0, n), (1, 1000, 1500), (20_000, n // 3, 2 * n // 3), (n + 5, 0, n)):
        with contextlib.redirect_stderr(io.StringIO()):
            want = plotfeed.sliding_window_average_depth(d[lo:hi], ws, 4.0 * 8, lo, "t")
            got = plotfeed.sliding_window_average_depth_gpu(ctx, 0, 0, lo, hi, ws, 4.0 * 8, lo, "t")
        _same(got, want)
