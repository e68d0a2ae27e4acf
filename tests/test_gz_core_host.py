"""The `.depth.gz` encoder core (gci_b200/csrc/gz_core.cuh: run bit counts, bit emission, CRC-32 of a run by table
powers, CRC combination) driven sequentially on the host and inflated with zlib — the CUDA kernels of gzip.cu call
the same functions per run."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("gz") / "gz_core_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "native", "gz_core_host.cpp")], check=True)
    lib = C.CDLL(so)
    lib.gz_host_encode.restype = C.c_int64
    return lib


def _encode(lib, depth, header=b""):
    depth = np.ascontiguousarray(depth, np.int32)
    cap = 4096 + len(depth) * 16 + 2 * len(header)
    cap += -cap % 4
    out = np.zeros(cap, np.uint8)
    n = lib.gz_host_encode(depth.ctypes.data_as(C.c_void_p), C.c_int64(len(depth)), header, len(header),
                           out.ctypes.data_as(C.c_void_p), C.c_int64(cap))
    assert n > 0, n
    return out[:n].tobytes()


def _text(depth, header=b""):
    return header + b"".join(b"%d\n" % int(v) for v in depth)


@pytest.mark.parametrize("case", ["coverage", "constant", "alternating", "wide", "short", "empty", "negative"])
def test_members_inflate_to_the_reference_text(harness, case):
    rng = np.random.default_rng(3)
    if case == "coverage":       # piecewise constant, runs of ~250, values around 30, a zero hole
        d = np.repeat(rng.poisson(30, 400), rng.integers(1, 600, 400))
        d[5000:9000] = 0
    elif case == "constant":     # whole members of one value: runs of 8192 lines
        d = np.concatenate([np.full(30000, 7), np.full(20000, 123456)])
    elif case == "alternating":  # every run has length 1 or 2 (the 1- and 2-byte tails)
        d = np.repeat(rng.integers(0, 12, 9000), rng.integers(1, 3, 9000))
    elif case == "wide":         # values with and without a table row, 1 to 10 digits
        d = np.repeat(rng.choice([0, 9, 10, 99, 100, 999, 1023, 1024, 99999, 2**31 - 1], 3000), rng.integers(1, 40, 3000))
    elif case == "short":
        d = np.array([5])
    elif case == "empty":
        d = np.zeros(0, np.int64)
    else:
        d = np.repeat(rng.integers(-20, 20, 500), rng.integers(1, 300, 500))
    for header in (b"", b">chr1_some_name\n"):
        blob = _encode(harness, d, header)
        assert gzip.decompress(blob) == _text(d, header)          # the gzip module checks every member's CRC-32 / ISIZE
        assert blob.count(b"\x1f\x8b\x08\x00") >= max(1, -(-len(d) // 8192))


def test_run_lengths_around_the_match_limits(harness):
    # (k - 1) * Lb around multiples of 258 (full matches, 1- and 2-byte tails) for 2-, 3- and 4-byte lines
    for v in (3, 42, 512):
        Lb = len(str(v)) + 1
        for rem in list(range(0, 9)) + [255, 256, 257, 258, 259, 260, 261, 515, 516, 517, 518, 519, 774, 775, 776]:
            if rem % Lb:
                continue
            d = np.concatenate([[1], np.full(rem // Lb + 1, v), [2]])
            assert gzip.decompress(_encode(harness, d)) == _text(d)
