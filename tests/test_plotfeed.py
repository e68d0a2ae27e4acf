"""Plot feed: gci_b200.plotfeed.sliding_window_average_depth against known answers of the
reference's GCI.py:660-705 (tests/golden/sliding_window_kat.json, made by make_golden.py) and, where the reference is
mounted, against the reference itself on fresh random inputs."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import pytest

from gci_b200.plotfeed import sliding_window_average_depth
from helpers import GOLDEN


def _run(fn, *args):
    err = io.StringIO()
    with contextlib.redirect_stderr(err):
        pos, val = fn(*args)
    return pos, val, err.getvalue()


def test_known_answers_of_the_reference():
    kat = json.load(open(os.path.join(GOLDEN, "sliding_window_kat.json")))
    assert len(kat["cases"]) >= 20
    for c in kat["cases"]:
        for depths in (c["depths"], np.asarray(c["depths"], np.int32)):
            pos, val, err = _run(sliding_window_average_depth, depths, c["window_size"], c["max_depth"], c["start"],
                                 c["target"])
            assert pos == c["positions"] and val.tolist() == c["values"] and str(val.dtype) == c["dtype"]
            assert err == c["stderr"]


@pytest.mark.skipif(not os.path.exists("/root/reference/GCI.py"), reason="the unmodified reference is not mounted here")
def test_against_the_reference_live():
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    ref = MG.import_reference()
    rng = np.random.default_rng(5)
    for trial in range(120):
        n = int(rng.integers(0, 3000))
        w = int(rng.choice([1, 3, 50, 500, 5000]))
        d = rng.integers(1, 60, n)
        d[rng.random(n) < rng.choice([0.0, 0.01, 0.2, 0.95])] = 0
        args = (d, w, float(rng.choice([5, 30, 1000])), int(rng.integers(0, 10**8)), "chr1")
        a, b = _run(ref.sliding_window_average_depth, *args), _run(sliding_window_average_depth, *args)
        assert a[0] == b[0] and a[1].dtype == b[1].dtype and np.array_equal(a[1], b[1]) and a[2] == b[2], trial
