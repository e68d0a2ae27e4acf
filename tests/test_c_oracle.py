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


@pytest.mark.parametrize("seed,threads,n_paf", [(0, 1, 1), (1, 4, 2), (2, 3, 3)])
def test_c_paf_leg_equals_python_oracle(seed, threads, n_paf):
    """orc_paf_leg (GCI.py:211-254 incl. the synteny dict leaking across PAF files) against the Python oracle,
    on reads with split lines, weaker lines on other contigs and exact score ties (contig name decides)."""
    lengths = [x // 300 for x in synth.CHM13_LENGTHS[:7]]
    names = ["chr10", "chr9", "chrB", "chrA", "chr1", "chr2", "chr11"]         # str order differs from index order
    w = synth.make_genome(lengths, names, coverage=12, seed=40 + seed, read_mean=6000, read_min=800, read_max=15000)
    pafs = [w.paf] + [synth.paf_second_aligner(w.paf, lengths, 90 + seed + k, split_frac=0.1, alt_frac=0.1,
                                               tie_frac=0.05) for k in range(1, n_paf)]
    sel = np.ones(len(lengths), bool)
    sel[3] = seed != 2
    want, want_hq = O.paf_leg(pafs, sel, names, 30, 0.9, 50)
    got, got_hq = CO.paf_leg(pafs, sel, CO.name_rank(names), w.n_reads, threads=threads)
    assert sorted(want_hq) == np.flatnonzero(got_hq).tolist()
    for f in range(n_paf):
        c, s, e, q = got[f]
        have = np.flatnonzero(c >= 0)
        assert sorted(want[f]) == have.tolist()
        for r in have.tolist():
            assert want[f][r] == (int(c[r]), int(s[r]), int(e[r]), int(q[r])), (f, r)
    # and the whole path with the PAFs joined first (GCI.py:272)
    want_d, want_s = O.filter_depth(pafs, [w.bam], names, lengths, selected=sel)
    got_d, got_b, n_surv = CO.hot_path([w.bam], lengths, w.n_reads, selected=sel, threads=threads, pafs=pafs, names=names)
    assert n_surv == len(want_s)
    for i in range(len(lengths)):
        assert np.array_equal(got_d[i], want_d[i])


def test_c_depth_hash_is_position_sensitive():
    rng = np.random.default_rng(5)
    d = rng.integers(0, 60, 100_000).astype(np.int64)
    h = CO.depth_hash(d, threads=3)
    assert h == CO.depth_hash(d, threads=1)
    x = d.copy(); x[[10, 11]] = x[[11, 10]]
    assert (CO.depth_hash(x) == h) == (d[10] == d[11])
    y = d.copy(); y[77] += 1; y[78] -= 1
    assert CO.depth_hash(y) != h
    # the definition the CUDA side restates: sum (depth + 1) * splitmix64(position) mod 2^64
    def sm(v):
        v = (v + 0x9E3779B97F4A7C15) & (2**64 - 1)
        v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return v ^ (v >> 31)
    assert CO.depth_hash(d[:50]) == sum((int(v) + 1) * sm(i) for i, v in enumerate(d[:50])) % 2**64


def test_write_depth_port_matches_reference_text_and_member_layout():
    """orc_write_depth (GCI.py:99-143): `threads` gzip members per contig, ">name" only in the first, one decimal per
    line; the concatenation inflates to the text the reference writes"""
    import gzip
    import zlib
    rng = np.random.default_rng(7)
    d = np.repeat(rng.integers(0, 60, 300), rng.integers(1, 900, 300)).astype(np.int64)
    d[3], d[77], d[-1] = -4, 2**40, 0
    want = ">chr1 some text\n" + "".join(f"{int(v)}\n" for v in d.tolist())
    for threads in (1, 4, 7):
        blob = CO.write_depth(d, "chr1 some text", threads, fetch=True)
        assert gzip.decompress(blob).decode() == want
        assert CO.write_depth(d, "chr1 some text", threads) == len(blob)
        members, rest = 0, blob
        while rest:
            o = zlib.decompressobj(31)
            o.decompress(rest)
            rest = o.unused_data
            members += 1
        stp = 1 + (len(d) - 1) // threads
        assert members == (len(d) + stp - 1) // stp
    assert CO.write_depth(np.zeros(0, np.int64), "empty", 4) == 0
