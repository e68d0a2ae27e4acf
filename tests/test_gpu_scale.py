"""Parity at scale: the CUDA path against the oracle's C port (oracle/gci_oracle.c, itself pinned to the Python
oracle and through it to the unmodified reference's outputs) — FULL depth arrays per contig, every issue interval,
the score rows and the whole-track checksum, on genomes of 50-310 Mbp.

What these cases cannot pin (stated wherever parity is claimed): the BAM byte stream is produced by
this repo's writer and decoded by this repo's reader — no htslib-produced file exists in this image."""
import os

import numpy as np
import pytest

from oracle import gci_oracle as O, c_oracle as CO
from gci_b200 import io as gio, synth
from gci_b200.records import AlnTable

pytestmark = pytest.mark.gpu

GATES = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)
THREADS = min(16, os.cpu_count() or 1)


@pytest.fixture(scope="module")
def ctx():
    from gci_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _upload(ctx, names, lengths, n_reads, pafs, bams):
    ctx.set_contigs(lengths)
    ctx.set_name_rank(CO.name_rank(names))
    ctx.reads_begin(n_reads)
    for t in pafs:
        ctx.upload_paf(t)
    for t in bams:
        ctx.upload_bam(t)


def _compare_track(ctx, track, lengths, want_depths, fl=15, ts=0, dp=0.005, names=None):
    """full depth, checksum, sums, intervals, score terms of one track against the oracle's arrays"""
    n = len(lengths)
    hashes = ctx.depth_hash(track)
    sums = ctx.depth_sums(track)
    for c in range(n):
        got = ctx.fetch_depth(track, c)
        assert np.array_equal(got, want_depths[c].astype(np.int32)), f"depth differs on contig {c}"
        assert int(hashes[c]) == CO.depth_hash(want_depths[c], THREADS)
        assert int(sums[c]) == int(want_depths[c].sum())
    n_iv = ctx.scan(track, -1, ts, fl)
    gs, ge, off = ctx.fetch_intervals(track, n)
    beds = [CO.collapse(want_depths[c], -1, ts, fl, 0) for c in range(n)]
    for c in range(n):
        assert list(zip(gs[off[c]:off[c + 1]].tolist(), ge[off[c]:off[c + 1]].tolist())) == beds[c], f"bed differs on contig {c}"
    assert n_iv == sum(len(b) for b in beds)
    n50, nctg, _, _ = ctx.score_terms(track, n, n_iv, dp, fl)
    rows = O.score_rows(names or [f"c{i}" for i in range(n)], [int(x) for x in lengths], beds, fl, dp)
    for c in range(n + 1):
        assert (int(n50[c]), int(nctg[c])) == (rows[c][2], rows[c][4]), (c, rows[c])
    return beds


def test_bam_plus_paf_310mbp_full_depth(ctx):
    """BASELINE configs[2] at 1/10 size (24 contigs, 311 Mbp, 30x HiFi, BAM + PAF, -op 0.9): the fused pipeline
    call and the staged calls both equal the C port base by base."""
    lengths = [x // 10 for x in synth.CHM13_LENGTHS]
    w = synth.make_genome(lengths, synth.CHM13_NAMES, coverage=30, seed=77)
    want_d, want_b, want_n = CO.hot_path([w.bam], lengths, w.n_reads, threads=THREADS, pafs=[w.paf],
                                         names=w.contigs.names, **GATES)
    _upload(ctx, w.contigs.names, lengths, w.n_reads, [w.paf], [w.bam])
    n_surv, n_iv, n50, nctg, sums = ctx.pipeline(0, len(lengths), flank_len=15, lo=-1, hi=0, dist_percent=0.005, **GATES)
    assert n_surv == want_n and n_iv == sum(len(b) for b in want_b)
    assert sums[:-1].tolist() == [int(d.sum()) for d in want_d]
    _compare_track(ctx, 0, lengths, want_d, names=w.contigs.names)
    # staged entry points on the same upload
    assert ctx.filter(**GATES) == want_n
    r, c, s, e = ctx.fetch_survivors()
    sc, ss, se = CO.survivors([w.bam], lengths, w.n_reads, threads=THREADS, pafs=[w.paf], names=w.contigs.names, **GATES)
    keep = np.flatnonzero(sc >= 0)
    assert np.array_equal(r, keep.astype(np.uint32)) and np.array_equal(c, sc[keep])
    assert np.array_equal(s, ss[keep]) and np.array_equal(e, se[keep])
    ctx.depth(0, 15, -1, 0)
    assert [int(h) for h in ctx.depth_hash(0)] == [CO.depth_hash(d, THREADS) for d in want_d]


def test_ont_60x_from_bam_file_with_cg_tags(ctx, tmp_path):
    """ONT 60x (thousands of ops per record, a few records beyond 65 535 ops stored in the CG tag) written to a
    real BAM file, decoded by libgci_io.so, filtered on the GPU: full depth equals the C port's on the same
    records."""
    lengths = [4_000_000, 2_000_000]
    names = ["ctgA", "ctgB"]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=60, seed=91, read_mean=30000, read_sigma=0.6, read_min=2000,
                                         read_max=150000, events_per_base=0.04, contig_names=names))
    ul = synth.make_reads(synth.SynthSpec([lengths[0]], coverage=1.5, seed=92, read_mean=1_200_000, read_sigma=0.1,
                                          read_min=1_000_000, read_max=1_500_000, events_per_base=0.04, hole_fraction=0.0))
    ul.bam.read_id += np.uint32(d.n_reads)
    both = synth.concat_aln([d.bam, ul.bam])
    both = both.take(np.lexsort((both.ref_start, both.ref_id)))
    n_ops = np.diff(both.cigar_off.astype(np.int64))
    assert (n_ops > 65535).sum() >= 2 and n_ops.mean() > 1500
    path = str(tmp_path / "ont.bam")
    gio.write_bam(path, names, lengths, both)
    rn, rl, tab = gio.read_bam(path, threads=THREADS)
    assert list(rn) == names and [int(x) for x in rl] == lengths and tab.n_records == both.n_records
    assert np.array_equal(np.diff(tab.cigar_off.astype(np.int64)), n_ops)          # CG records carry their real CIGAR
    n_reads = int(tab.read_id.max()) + 1
    gates = dict(GATES, iden_percent=0.85)
    want_d, want_b, want_n = CO.hot_path([tab], lengths, n_reads, threads=THREADS, flank_len=0, threshold=5, **gates)
    _upload(ctx, names, lengths, n_reads, [], [tab])
    assert ctx.filter(**gates) == want_n
    ctx.depth(0, 0, -1, 5)
    _compare_track(ctx, 0, lengths, want_d, fl=0, ts=5, names=names)
    st, end = ctx.fetch_cigar_stats(0, tab.n_records)
    ops = tab.op_sums()
    assert np.array_equal(st[:, 0].astype(np.int64), ops[:, 0] + ops[:, 7] + ops[:, 8])
    assert np.array_equal(end.astype(np.int64), tab.ref_start + np.maximum(1, tab.ref_len()))


def test_two_type_tracks_60mbp_full_depth(ctx):
    """HiFi (BAM + PAF) and ONT (two BAMs) on one 62 Mbp genome with N-runs: HiFi, Nano and max(HiFi, Nano) tracks,
    masked like the driver does (GCI.py:993-1023), equal the C port base by base."""
    lengths = [x // 50 for x in synth.CHM13_LENGTHS]
    names = synth.CHM13_NAMES
    h = synth.make_genome(lengths, names, coverage=30, seed=5)
    o = synth.make_genome(lengths, names, coverage=40, seed=6, with_paf=False, read_mean=30000, read_sigma=0.6,
                          read_min=2000, read_max=150000, events_per_base=0.04)
    o2 = synth.second_aligner(type("D", (), {"bam": o.bam, "spec": synth.SynthSpec(lengths, seed=8),
                                             "contigs": o.contigs})(), seed=9)
    n_runs = [r for r in h.n_runs]
    hd, _, hn = CO.hot_path([h.bam], lengths, h.n_reads, threads=THREADS, pafs=[h.paf], names=names, **GATES)
    od, _, on = CO.hot_path([o.bam, o2], lengths, o.n_reads, threads=THREADS, **GATES)
    ctx.set_contigs(lengths)
    ctx.set_name_rank(CO.name_rank(names))
    ctx.set_n_runs([c for c, r in enumerate(n_runs) for _ in r], [iv[0] for r in n_runs for iv in r],
                   [iv[1] for r in n_runs for iv in r])
    ctx.reads_begin(h.n_reads)
    ctx.upload_paf(h.paf)
    ctx.upload_bam(h.bam)
    assert ctx.filter(**GATES) == hn
    ctx.depth(0, 15, -1, 0)
    ctx.mask_gaps(0)
    ctx.reads_begin(o.n_reads)
    ctx.upload_bam(o.bam)
    ctx.upload_bam(o2)
    assert ctx.filter(**GATES) == on
    ctx.depth(1, 15, -1, 0)
    ctx.mask_gaps(1)
    ctx.merge_max(0, 1, 2, -1, 0)
    ctx.mask_gaps(2)
    hd, od = O.mask_gaps(hd, n_runs), O.mask_gaps(od, n_runs)
    md = O.mask_gaps(O.merge_two_types(hd, od), n_runs)
    for track, want in ((0, hd), (1, od), (2, md)):
        _compare_track(ctx, track, lengths, want, names=names)


def test_chr19_full_size_vs_c_port(ctx):
    """BASELINE configs[1] (58 Mbp, 30x HiFi, one BAM) at full size: survivors, full depth, intervals from the C port
    (replaces the round-1 check that derived its expectation from the GPU's own survivors)."""
    L = 58_000_000
    d = synth.make_reads(synth.SynthSpec([L], coverage=30, seed=20240634, contig_names=["chr19"]))
    want_d, want_b, want_n = CO.hot_path([d.bam], [L], d.n_reads, threads=THREADS, **GATES)
    _upload(ctx, ["chr19"], [L], d.n_reads, [], [d.bam])
    assert ctx.filter(**GATES) == want_n
    ctx.depth(0, 15, -1, 0)
    beds = _compare_track(ctx, 0, [L], want_d, names=["chr19"])
    for hs, he in d.holes[0]:                      # single file: every synthetic hole lies inside an issue interval
        assert any(s <= hs and e >= he for s, e in beds[0])
