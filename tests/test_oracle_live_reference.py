"""The oracle against the UNMODIFIED reference, live: fresh seeded random cases are pushed through the reference's
own GCI() driver (tests/golden/make_golden.py: pysam / Bio / matplotlib shims) and through the
oracle in the same process, and every output file must agree byte for byte.  Runs only where the reference is
mounted (the build container); the committed fixtures of tests/golden/ carry the same check everywhere else."""
import os
import sys

import pytest

from oracle import gci_oracle as O
from helpers import GOLDEN, load_case, assert_outputs_equal

REF = "/root/reference/GCI.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="the unmodified reference is not mounted here")


@pytest.fixture(scope="module")
def live_cases():
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    ref = MG.import_reference()
    store, meta = MG.make_random_cases(ref, n_cases=12, seed=77001, prefix="live", save=False)
    assert len(meta["cases"]) >= 8          # the reference itself may give up on a few random argument mixes
    return store, meta


def test_oracle_matches_reference_on_fresh_random_cases(live_cases):
    store, meta = live_cases
    for case in meta["cases"]:
        _, kw, expected = load_case(case["name"], store, meta)
        assert_outputs_equal(O.run_gci(**kw), expected)
