"""The oracle against the reference: golden files made by the unmodified reference
(tests/golden/make_golden.py), the reference's own MH63 example, and the KATs of SURVEY.md §4.4."""
import os

import numpy as np
import pytest

from oracle import gci_oracle as O
from gci_b200.records import AlnTable, PafTable
from helpers import GOLDEN, case_names, load_case, assert_outputs_equal, mh63_depths


@pytest.mark.parametrize("name", case_names())
def test_oracle_matches_reference_outputs(name):
    case, kw, expected = load_case(name)
    got = O.run_gci(**kw)
    assert_outputs_equal(got, expected)


def test_oracle_mh63_example():
    """depth -> BED -> .gci on the reference's own example."""
    names, lengths, depths = mh63_depths()
    assert sum(lengths) == 395765488
    beds = [O.collapse_depth_range(d, -1, 0, 15, 0) for d in depths]
    assert O.bed_text(names, beds) == open(os.path.join(GOLDEN, "mh63.0.depth.bed")).read()
    rows = O.score_rows(names, lengths, beds, 15, 0.005)
    assert O.gci_text("HiFi", rows) == open(os.path.join(GOLDEN, "mh63.gci")).read()


# ---- SURVEY.md §4.4 known-answer vectors -------------------------------------------------

def _bam(rows):
    base = dict(ref_id=0, mapq=60, flag=0, nm=0)
    out = []
    for i, r in enumerate(rows):
        d = dict(base)
        d.update(r)
        d.setdefault("read_id", i)
        from gci_b200.records import pack_cigar
        ops = pack_cigar(d["cigar"])
        d.setdefault("qlen", int(sum((o >> 4) for o in ops if (o & 15) in (0, 1, 4, 7, 8))))
        out.append(d)
    return AlnTable.from_rows(out)


def _covered(depth):
    nz = np.flatnonzero(depth)
    return (int(nz[0]), int(nz[-1]), len(nz)) if len(nz) else None


def test_kat_depth_interval():
    for start, cigar, want in [(100, "200M", (115, 285, 171)), (500, "20M", None), (0, "14M", None),
                               (0, "10M", (15, 995, 981))]:
        t = _bam([dict(ref_start=start, cigar=cigar)])
        d, _ = O.filter_depth([], [t], ["c"], [1000])
        assert _covered(d[0]) == want, (start, cigar)


def test_kat_record_gates():
    keep = ["10S100M", "500H100M", ("100M", 10), ("90=10X", 10), ("50M5I50M5D", 10)]
    drop = ["12S100M", ("100M", 11)]
    for spec, expect in [(k, True) for k in keep] + [(k, False) for k in drop]:
        cigar, nm = spec if isinstance(spec, tuple) else (spec, 0)
        t = _bam([dict(ref_start=100, cigar=cigar, nm=nm)])
        surv, hq = O.bam_leg(t, np.array([True]), 30, 0.1, 0.9, 50)
        assert (0 in surv) == expect, spec
    for kwargs, expect, isq in [(dict(flag=0x100), False, False), (dict(flag=0x800), False, False),
                                (dict(flag=0x4), False, False), (dict(mapq=29), False, False),
                                (dict(mapq=30), True, False), (dict(mapq=50), True, True)]:
        t = _bam([dict(ref_start=100, cigar="100M", **kwargs)])
        surv, hq = O.bam_leg(t, np.array([True]), 30, 0.1, 0.9, 50)
        assert (0 in surv) == expect and (0 in hq) == isq, kwargs


def test_kat_last_record_wins():
    t = _bam([dict(ref_start=100, cigar="200M", read_id=0), dict(ref_start=400, cigar="200M", read_id=0)])
    d, _ = O.filter_depth([], [t], ["c"], [1000])
    assert _covered(d[0]) == (415, 585, 171)


def _one(ref_id, start, mapq, qlen=1000):
    return _bam([dict(ref_id=ref_id, ref_start=start, cigar="1000M", mapq=mapq, qlen=qlen)])


def _empty():
    return _bam([])


@pytest.mark.parametrize("a,b,want", [
    ((0, 1000, 40), (0, 1000, 40), (0, 1015, 1985)),
    ((0, 1000, 40), (0, 1050, 40), (0, 1065, 1985)),
    ((0, 1000, 40), (0, 1150, 40), None),
    ((0, 1000, 60), (0, 1150, 40), None),
    ((0, 1000, 40), (1, 1000, 40), None),
    ((0, 1000, 40), None, None),
    ((0, 1000, 60), None, (0, 1015, 1985)),
    (None, (0, 1000, 60), (0, 1015, 1985)),
    (None, (0, 1000, 40), None),
])
def test_kat_two_file_join(a, b, want):
    ta = _one(*a) if a else _empty()
    tb = _one(*b) if b else _empty()
    d, _ = O.filter_depth([], [ta, tb], ["c1", "c2"], [5000, 5000])
    got = None
    for c in range(2):
        cv = _covered(d[c])
        if cv:
            got = (c, cv[0], cv[1])
    assert got == want


def test_kat_join_later_file_denominator_and_readd():
    d, _ = O.filter_depth([], [_one(0, 1000, 40, 1000), _one(0, 1000, 40, 2000)], ["c1", "c2"], [5000, 5000])
    assert _covered(d[0]) is None
    d, _ = O.filter_depth([], [_one(0, 1000, 60), _one(1, 1000, 40), _one(1, 3000, 40)], ["c1", "c2"], [5000, 5000])
    assert _covered(d[0]) is None and _covered(d[1])[:2] == (3015, 3985)


def test_kat_paf_first_and_synteny_leak():
    # PAFs come first regardless of CLI order; a read seen in PAF #1 only still appears in paf_lines[1]
    p1 = PafTable.from_rows([dict(read_id=0, qlen=1000, qstart=0, qend=1000, ref_id=0, tstart=1000, tend=2000,
                                  nmatch=1000, alnlen=1000, mapq=40)])
    p2 = PafTable.from_rows([])
    bam = _one(0, 1000, 40)
    d, surv = O.filter_depth([p1, p2], [bam], ["c1"], [5000])
    assert _covered(d[0])[:2] == (1015, 1985)


@pytest.mark.parametrize("zeros,want", [
    ([], []), ([(0, 100)], [(15, 85)]), ([(0, 20)], []), ([(0, 30)], []), ([(20, 25)], []),
    ([(0, 31)], [(15, 31)]), ([(40, 50)], [(40, 50)]), ([(80, 100)], [(80, 85)]), ([(84, 85)], [(84, 85)]),
    ([(85, 86)], []),
])
def test_kat_gap_scan(zeros, want):
    d = np.ones(100, dtype=np.int64)
    for s, e in zeros:
        d[s:e] = 0
    assert O.collapse_depth_range(d, -1, 0, 15, 0) == want
    assert O.collapse_depth_range_loop(d, -1, 0, 15, 0) == want


def test_kat_gap_scan_short_and_regions_mode():
    assert O.collapse_depth_range(np.zeros(10, np.int64), -1, 0, 15, 0) == []
    d = np.ones(100, dtype=np.int64)
    d[0:3] = 0
    d[40:50] = 0
    assert O.collapse_depth_range(d, -1, 0, 0, 1000) == [(1000, 1003), (1040, 1050)]


def test_gap_scan_vectorised_equals_loop():
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(0, 120))
        fl = int(rng.integers(0, 20))
        d = (rng.random(n) < rng.random()).astype(np.int64) * rng.integers(1, 4, n)
        ts = int(rng.integers(0, 3))
        assert O.collapse_depth_range(d, -1, ts, fl, 7) == O.collapse_depth_range_loop(d, -1, ts, fl, 7)


def test_kat_score_helpers():
    L = 100000
    bed = [(40000, 40100), (40300, 40400), (90000, 90010)]
    assert O.merge_close(bed, L) == [(15, 15), (40000, 40400), (90000, 90010)]
    assert O.complement_lengths(bed, L) == [39985, 200, 49600, 9975]
    assert O.complement_lengths(O.merge_close(bed, L), L) == [39985, 49600, 9975]
    bed = [(100, 200), (99700, 99800)]
    assert O.merge_close(bed, L) == [(15, 200), (99700, 99985)]
    assert O.complement_lengths(O.merge_close(bed, L), L) == [99500]
    assert O.merge_close([], L) == [(15, 15)] and O.complement_lengths([], L) == [99970]
    assert O.complement_lengths(O.merge_close([(15, 99985)], L), L) == []
    assert O.n50([]) == 0 and O.n50([5, 4, 3, 2, 1]) == 4 and O.n50([2, 2]) == 2


def test_depth_text_format():
    txt = O.depth_text(["c1", "c2"], [np.zeros(10, np.int64), np.zeros(7, np.int64)])
    assert txt == ">c1\n" + "0\n" * 10 + ">c2\n" + "0\n" * 7
