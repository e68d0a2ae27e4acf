"""Shared test helpers: golden-case loading and output comparison."""
from __future__ import annotations

import json
import os

import numpy as np

from gci_b200.records import AlnTable, PafTable

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_ALN_COLS = ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id", "cigar_off", "cigar")
_PAF_COLS = ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")

_cache = {}


class _Stores:
    """the hand-picked cases (filter_cases) and the seeded random ones (filter_cases_rand) behind one mapping"""

    def __init__(self, stores):
        self.stores = stores

    def __getitem__(self, key):
        for s in self.stores:
            if key in s.files:
                return s[key]
        raise KeyError(key)


def golden_store():
    if "store" not in _cache:
        stores, cases = [], []
        for stem in ("filter_cases", "filter_cases_rand"):
            stores.append(np.load(os.path.join(GOLDEN, stem + ".npz")))
            with open(os.path.join(GOLDEN, stem + ".json")) as f:
                meta = json.load(f)
            cases += meta["cases"]
        _cache["store"] = _Stores(stores)
        _cache["meta"] = {"cases": cases, "reference_commit": meta["reference_commit"]}
    return _cache["store"], _cache["meta"]


def case_names():
    _, meta = golden_store()
    return [c["name"] for c in meta["cases"]]


def load_case(name, store=None, meta=None):
    """-> (case meta, kwargs for run_gci-style drivers, expected outputs); from the committed fixtures unless a
    freshly generated (store, meta) pair is given."""
    if store is None:
        store, meta = golden_store()
    case = next(c for c in meta["cases"] if c["name"] == name)
    kw = dict(names=case["names"], lengths=case["lengths"], n_runs=[[tuple(iv) for iv in r] for r in case["n_runs"]],
              chrs=case["chrs"], regions=[tuple(r) for r in case["regions"]] if case["regions"] else None,
              prefix="T")
    kw.update(case["args"])
    for rtype in ("hifi", "nano"):
        tabs = None
        if rtype in case["files"]:
            tabs = []
            for f in case["files"][rtype]:
                if f["kind"] == "bam":
                    tabs.append(AlnTable(*[store[f"{f['key']}.{c}"] for c in _ALN_COLS]))
                else:
                    tabs.append(PafTable(*[store[f"{f['key']}.{c}"] for c in _PAF_COLS]))
        kw[rtype] = tabs
    expected = {}
    for fn, o in case["outputs"].items():
        if "text" in o:
            expected[fn] = o["text"]
        else:
            expected[fn] = {cn: store[f"{name}.out.{fn}.{cn}"] for cn in o["contigs"]}
    return case, kw, expected


def assert_outputs_equal(got: dict, expected: dict):
    assert sorted(got) == sorted(expected), (sorted(got), sorted(expected))
    for fn, exp in expected.items():
        g = got[fn]
        if isinstance(exp, dict):
            assert list(g.keys()) == list(exp.keys()), fn
            for cn in exp:
                a, b = np.asarray(g[cn]).astype(np.int64), np.asarray(exp[cn]).astype(np.int64)
                assert a.shape == b.shape, (fn, cn, a.shape, b.shape)
                if not np.array_equal(a, b):
                    bad = np.flatnonzero(a != b)
                    raise AssertionError(f"{fn}:{cn} depth differs at {len(bad)} positions, first {bad[:5]} "
                                         f"got {a[bad[:5]]} want {b[bad[:5]]}")
        else:
            assert g == exp, f"{fn} differs:\n--- got\n{g[:2000]}\n--- want\n{exp[:2000]}"


def mh63_depths():
    z = np.load(os.path.join(GOLDEN, "mh63_depth_rle.npz"))
    names = [str(n) for n in z["names"]]
    counts = z["run_counts"]
    off = np.concatenate([[0], np.cumsum(counts)])
    depths = []
    for i in range(len(names)):
        v = z["values"][off[i]:off[i + 1]]
        r = z["run_lengths"][off[i]:off[i + 1]]
        depths.append(np.repeat(v, r).astype(np.int32))
    return names, [int(x) for x in z["lengths"]], depths


# ---- running the product (GPU) on a case ---------------------------------------------------------
def attach_contigs(tabs, names, lengths):
    for t in tabs or ():
        if isinstance(t, AlnTable):
            t.contig_names, t.contig_lengths = list(names), [int(x) for x in lengths]
    return tabs


def run_product(kw, workdir, session=None, threads=2):
    """Drive gci_b200.pipeline.GCI() with the in-memory tables of a case; returns the output files in the
    same shape as oracle.run_gci / the golden fixtures."""
    import contextlib
    import io as _io
    from gci_b200 import pipeline as P
    from gci_b200 import io as gio

    kw = dict(kw)
    names, lengths = kw.pop("names"), kw.pop("lengths")
    n_runs = kw.pop("n_runs", None)
    hifi, nano = kw.pop("hifi", None), kw.pop("nano", None)
    attach_contigs(hifi, names, lengths)
    attach_contigs(nano, names, lengths)
    regions = kw.pop("regions", None)
    regions_bed = None
    if regions:
        regions_bed = {}
        for c, s, e in regions:
            regions_bed.setdefault(c, []).append((s, e))
    gaps = {n: list(r) for n, r in zip(names, n_runs or []) if r}
    prefix = kw.pop("prefix", "T")
    out_dir = os.path.join(workdir, "out")
    buf = _io.StringIO()
    with contextlib.redirect_stdout(buf):
        P.GCI(hifi=hifi, nano=nano, directory=out_dir, prefix=prefix, reference={"ids": names, "gaps": gaps},
              regions=regions_bed, chrs=kw.pop("chrs", None), threads=threads, force=True, session=session, **kw)
    outs = {}
    for fn in sorted(os.listdir(out_dir)):
        p = os.path.join(out_dir, fn)
        outs[fn] = gio.read_depth_gz_py(p) if fn.endswith(".depth.gz") else open(p).read()
    return outs, buf.getvalue()
