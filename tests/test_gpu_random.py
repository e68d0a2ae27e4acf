"""Seeded random inputs: CUDA path vs the CPU oracle, stage by stage through the C ABI."""
import numpy as np
import pytest

from oracle import gci_oracle as O
from gci_b200 import synth
from gci_b200.records import AlnTable, PafTable, pack_cigar, NM_MISSING

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gci_b200._lib import Context
    c = Context(0)
    yield c
    c.close()


def _name_rank(names):
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.empty(len(names), np.int32)
    rank[order] = np.arange(len(names))
    return rank


def _run_gpu(ctx, names, lengths, pafs, bams, n_reads, selected=None, fl=15, ts=0, params=None, n_runs=None):
    p = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)
    p.update(params or {})
    ctx.set_contigs(lengths, selected)
    ctx.set_name_rank(_name_rank(names))
    if n_runs:
        c = [i for i, r in enumerate(n_runs) for _ in r]
        s = [iv[0] for r in n_runs for iv in r]
        e = [iv[1] for r in n_runs for iv in r]
        ctx.set_n_runs(c, s, e)
    ctx.reads_begin(n_reads)
    for t in pafs:
        ctx.upload_paf(t)
    for t in bams:
        ctx.upload_bam(t)
    n_surv = ctx.filter(**p)
    ctx.depth(0, fl, -1, ts)
    if n_runs:
        ctx.mask_gaps(0)
    return n_surv


def _check(ctx, names, lengths, pafs, bams, n_reads, selected=None, fl=15, ts=0, params=None, n_runs=None, dp=0.005):
    p = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)
    p.update(params or {})
    sel = np.ones(len(lengths), bool) if selected is None else np.asarray(selected, bool)
    n_surv = _run_gpu(ctx, names, lengths, pafs, bams, n_reads, selected, fl, ts, params, n_runs)
    want_d, want_s = O.filter_depth(pafs, bams, names, lengths, sel, flank_len=fl, **p)
    want_d = O.mask_gaps(want_d, n_runs)
    assert n_surv == len(want_s)
    r, c, s, e = ctx.fetch_survivors()
    got_s = {int(a): (int(b), int(x), int(y)) for a, b, x, y in zip(r, c, s, e)}
    assert got_s == {k: tuple(v[:3]) for k, v in want_s.items()}
    sums = ctx.depth_sums(0)
    owners = [i for i in range(len(lengths)) if sel[i]]
    for i in owners:
        got = ctx.fetch_depth(0, i)
        assert np.array_equal(got.astype(np.int64), want_d[i]), f"depth differs on contig {i}"
        assert int(sums[i]) == int(want_d[i].sum())
    n_iv = ctx.scan(0, -1, ts, fl)
    gs, ge, off = ctx.fetch_intervals(0, len(owners))
    beds = []
    for o, i in enumerate(owners):
        want = O.collapse_depth_range(want_d[i], -1, ts, fl, 0)
        got = list(zip(gs[off[o]:off[o + 1]].tolist(), ge[off[o]:off[o + 1]].tolist()))
        assert got == want, f"bed differs on contig {i}"
        beds.append(want)
    n50, nctg, lens, loff = ctx.score_terms(0, len(owners), n_iv, dp, fl)
    all_obs, all_ctg = [], 0
    for o, i in enumerate(owners):
        obs = O.complement_lengths(beds[o], int(lengths[i]), fl)
        new = O.complement_lengths(O.merge_close(beds[o], int(lengths[i]), dp, fl), int(lengths[i]), fl)
        assert lens[loff[o]:loff[o + 1]].tolist() == obs
        assert int(n50[o]) == O.n50(obs) and int(nctg[o]) == len(new)
        all_obs += obs
        all_ctg += len(new)
    assert int(n50[-1]) == O.n50(all_obs) and int(nctg[-1]) == all_ctg
    with_sums = ctx.score_terms(0, len(owners), n_iv, dp, fl, with_sums=True)[4]
    assert with_sums[:-1].tolist() == [int(want_d[i].sum()) for i in owners] and int(with_sums[-1]) == sum(
        int(want_d[i].sum()) for i in owners)


@pytest.mark.parametrize("seed", range(6))
def test_random_hifi_single_bam(ctx, seed):
    rng = np.random.default_rng(seed)
    lengths = [int(x) for x in rng.integers(20_000, 400_000, int(rng.integers(1, 5)))]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=float(rng.uniform(5, 40)), seed=100 + seed,
                                         read_mean=8000, read_min=1000, read_max=20000, hole_fraction=0.02))
    _check(ctx, d.contigs.names, lengths, [], [d.bam], d.n_reads, n_runs=d.n_runs,
           fl=int(rng.integers(0, 40)), ts=int(rng.integers(0, 4)))


@pytest.mark.parametrize("seed", range(4))
def test_random_ont_long_cigars(ctx, seed):
    """records of thousands of ops: every record spans several op tiles"""
    lengths = [300_000, 150_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=20, seed=200 + seed, read_mean=30000, read_sigma=0.6,
                                         read_min=2000, read_max=120000, events_per_base=0.05))
    assert d.bam.n_ops / max(1, d.bam.n_records) > 1000
    _check(ctx, d.contigs.names, lengths, [], [d.bam], d.n_reads)


@pytest.mark.parametrize("seed", range(4))
def test_random_multi_file_join(ctx, seed):
    lengths = [200_000, 120_000, 60_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=25, seed=300 + seed, read_mean=7000, read_min=1000,
                                         read_max=20000))
    b1 = synth.drop_reads(d.bam, 0.03, seed)
    b2 = synth.second_aligner(d, seed=seed + 50)
    b3 = synth.second_aligner(d, seed=seed + 60)
    p1 = synth.aln_to_paf(synth.second_aligner(d, seed=seed + 70))
    p2 = synth.aln_to_paf(synth.second_aligner(d, seed=seed + 80))
    combos = [([], [b1, b2]), ([p1], [b1]), ([p1, p2], [b1, b2, b3]), ([p1], [b2, b3])]
    pafs, bams = combos[seed % len(combos)]
    _check(ctx, d.contigs.names, lengths, pafs, bams, d.n_reads,
           params=dict(ovlp_percent=[0.9, 0.5, 0.95, 0.0][seed % 4], mq_cutoff=[50, 40, 60, 30][seed % 4]))


def test_chrs_selection(ctx):
    lengths = [100_000, 80_000, 50_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=15, seed=77, read_mean=6000, read_min=1000, read_max=15000))
    b2 = synth.second_aligner(d, seed=5)
    _check(ctx, d.contigs.names, lengths, [synth.aln_to_paf(b2)], [d.bam], d.n_reads, selected=[True, False, True])


def _bam(rows, n_contigs=1):
    out = []
    for i, r in enumerate(rows):
        d = dict(ref_id=0, mapq=60, flag=0, nm=0, read_id=i)
        d.update(r)
        ops = pack_cigar(d["cigar"])
        d.setdefault("qlen", int(sum((int(o) >> 4) for o in ops if (int(o) & 15) in (0, 1, 4, 7, 8))))
        out.append(d)
    return AlnTable.from_rows(out)


def test_edge_records(ctx):
    """zero-op records, 1-op records, slice wrap quirk, duplicate names, unsorted contig order"""
    rows = [dict(ref_start=0, cigar="10M"),                       # stop = -4 -> wraps
            dict(ref_start=100, cigar="200M", read_id=7), dict(ref_start=400, cigar="200M", read_id=7),
            dict(ref_start=500, cigar="*", flag=4), dict(ref_start=500, cigar="20M"),
            dict(ref_start=600, cigar="12S100M"), dict(ref_start=610, cigar="500H100M"),
            dict(ref_start=620, cigar="50M5I50M5D", nm=10), dict(ref_start=630, cigar="100M", nm=11),
            dict(ref_start=700, cigar="90=10X", nm=10), dict(ref_start=710, cigar="100M", mapq=29),
            dict(ref_start=720, cigar="100M", flag=0x100), dict(ref_start=730, cigar="100M", flag=0x800),
            dict(ref_start=900, cigar="100M"), dict(ref_start=980, cigar="20M")]
    t = _bam(rows)
    _check(ctx, ["c"], [1000], [], [t], 32)
    # contig order differs from file order: contig 1 records first
    rows = [dict(ref_id=1, ref_start=100, cigar="300M", read_id=3), dict(ref_id=0, ref_start=50, cigar="300M", read_id=3)]
    _check(ctx, ["a", "b"], [1000, 1000], [], [_bam(rows)], 8)
    # no records at all
    _check(ctx, ["a", "b"], [1000, 10], [], [_bam([])], 4)
    # a contig shorter than 2*fl and one exactly a tile long
    rows = [dict(ref_id=1, ref_start=0, cigar="8192M"), dict(ref_id=2, ref_start=8000, cigar="8384M")]
    _check(ctx, ["tiny", "tile", "tile2"], [20, 8192, 16384], [], [_bam(rows)], 4)


def test_reference_would_raise(ctx):
    from gci_b200._lib import ReferenceWouldRaise
    t = _bam([dict(ref_start=10, cigar="100M", nm=None)])
    with pytest.raises(ReferenceWouldRaise):
        _run_gpu(ctx, ["c"], [1000], [], [t], 4)
    t = _bam([dict(ref_start=10, cigar="100H")])
    with pytest.raises(ReferenceWouldRaise):
        _run_gpu(ctx, ["c"], [1000], [], [t], 4)
    with pytest.raises(O.ReferenceWouldRaise):
        O.filter_depth([], [t], ["c"], [1000])


def test_many_tiny_records_per_tile(ctx):
    """> 1024 records inside one 2048-op tile takes the global-atomics path"""
    n = 5000
    rows = [dict(ref_start=int(i * 3), cigar="40M", read_id=i) for i in range(n)]
    _check(ctx, ["c"], [30_000], [], [_bam(rows)], n)


def test_two_type_max_and_text(ctx):
    lengths = [90_000, 40_000]
    a = synth.make_reads(synth.SynthSpec(lengths, coverage=12, seed=1, read_mean=5000, read_min=500, read_max=12000))
    b = synth.make_reads(synth.SynthSpec(lengths, coverage=9, seed=2, read_mean=9000, read_min=500, read_max=30000))
    _run_gpu(ctx, a.contigs.names, lengths, [], [a.bam], a.n_reads)
    ctx.reads_begin(b.n_reads)
    ctx.upload_bam(b.bam)
    ctx.filter()
    ctx.depth(1, 15)
    ctx.merge_max(0, 1, 2, -1, 0)
    da, _ = O.filter_depth([], [a.bam], a.contigs.names, lengths)
    db, _ = O.filter_depth([], [b.bam], a.contigs.names, lengths)
    want = O.merge_two_types(da, db)
    for i in range(2):
        assert np.array_equal(ctx.fetch_depth(2, i).astype(np.int64), want[i])
        assert ctx.depth_text(2, i).tobytes().decode() == "".join(f"{int(v)}\n" for v in want[i])
    assert ctx.depth_sums(2).tolist() == [int(w.sum()) for w in want]
    ctx.scan(2, -1, 0, 15)
    gs, ge, off = ctx.fetch_intervals(2, 2)
    for i in range(2):
        assert list(zip(gs[off[i]:off[i + 1]].tolist(), ge[off[i]:off[i + 1]].tolist())) == \
            O.collapse_depth_range(want[i], -1, 0, 15, 0)


def test_regions_windows(ctx):
    rng = np.random.default_rng(3)
    lengths = [50_000, 20_000]
    depth = [(rng.random(l) < 0.97).astype(np.int32) * rng.integers(1, 5, l).astype(np.int32) for l in lengths]
    for d in depth:
        for _ in range(6):
            s = int(rng.integers(0, len(d) - 500))
            d[s:s + int(rng.integers(1, 400))] = 0
    ctx.set_contigs(lengths)
    for i, d in enumerate(depth):
        ctx.load_depth(0, i, d)
    wins = [(0, 0, 50_000), (0, 100, 5000), (1, 19_000, 20_000), (1, 5, 5), (0, 31, 64), (0, 49_990, 60_000)]
    for ts in (0, 2):
        n_iv = ctx.scan_windows(0, [w[0] for w in wins], [w[1] for w in wins], [w[2] for w in wins], -1, ts)
        gs, ge, off = ctx.fetch_intervals(0, len(wins))
        n50, nctg, lens, loff = ctx.score_terms(0, len(wins), n_iv, 0.005, 0)
        for k, (c, s, e) in enumerate(wins):
            sub = depth[c][s:e].astype(np.int64)
            want = O.collapse_depth_range(sub, -1, ts, 0, s)
            assert list(zip(gs[off[k]:off[k + 1]].tolist(), ge[off[k]:off[k + 1]].tolist())) == want
            obs = O.complement_lengths(want, e - s, s, s, e)
            merged = O.merge_close(want, e - s, 0.005, s, s, e)
            assert int(n50[k]) == O.n50(obs)
            assert int(nctg[k]) == len(O.complement_lengths(merged, e - s, s, s, e))


def test_scan_staged_and_walked_chunks(ctx):
    """run extraction: a 65 536-position chunk with at most 16 run starts / ends stages them while counting, one with
    more is walked a second time (csrc/scan.cu RUN_STAGE): both kinds side by side, a run across a chunk border, a run
    that starts on the last position of a chunk, flanks 0 and 15, thresholds 0 and 3"""
    L = 4 * 65536 + 1234
    d = np.full(L, 9, np.int32)
    for k in range(16):                                   # chunk 0: exactly 16 runs -> staged
        d[1000 + 300 * k:1000 + 300 * k + 7] = 0
    for k in range(17):                                   # chunk 1: 17 runs -> walked
        d[65536 + 500 + 200 * k:65536 + 500 + 200 * k + 3] = 2
    d[2 * 65536 - 40:2 * 65536 + 90] = 0                  # across the border of chunks 1 and 2
    d[3 * 65536 - 1:3 * 65536 + 5] = 1                    # starts on the last position of chunk 2
    d[L - 20:] = 0                                        # reaches the end of the contig
    ctx.set_contigs([L, 70_000])
    ctx.load_depth(0, 0, d)
    e = np.full(70_000, 5, np.int32)
    e[::2] = 0                                            # 35 000 runs in two chunks
    ctx.load_depth(0, 1, e)
    for fl in (0, 15):
        for hi in (0, 3):
            ctx.scan(0, -1, hi, fl)
            gs, ge, off = ctx.fetch_intervals(0, 2)
            for i, dep in enumerate((d, e)):
                got = list(zip(gs[off[i]:off[i + 1]].tolist(), ge[off[i]:off[i + 1]].tolist()))
                assert got == O.collapse_depth_range(dep.astype(np.int64), -1, hi, fl, 0), (fl, hi, i)


def test_full_size_chr19_properties(ctx):
    """BASELINE config 2 size (58 Mbp, 30x): size-independent properties instead of the slow oracle"""
    L = 58_000_000
    d = synth.make_reads(synth.SynthSpec([L], coverage=30, seed=20240634, contig_names=["chr19"]))
    n_surv = _run_gpu(ctx, ["chr19"], [L], [], [d.bam], d.n_reads)
    r, c, s, e = ctx.fetch_survivors()
    assert len(r) == n_surv and n_surv > 0.6 * d.n_reads
    # (1) sum of depth == sum of survivor slice lengths
    a = np.clip(s.astype(np.int64) + 15, 0, L)
    b = np.clip(e.astype(np.int64) - 15 + 1, 0, L)
    assert int(ctx.depth_sums(0)[0]) == int(np.maximum(0, b - a).sum())
    # (2) the depth array equals the numpy delta/cumsum construction from the survivors
    delta = np.zeros(L + 1, np.int64)
    np.add.at(delta, a[b > a], 1)
    np.add.at(delta, b[b > a], -1)
    want = np.cumsum(delta[:-1])
    got = ctx.fetch_depth(0, 0)
    assert np.array_equal(got, want.astype(np.int32))
    # (3) intervals == runs of zeros, and every synthetic hole is inside an issue interval
    ctx.scan(0, -1, 0, 15)
    gs, ge, off = ctx.fetch_intervals(0, 1)
    assert list(zip(gs.tolist(), ge.tolist())) == O.collapse_depth_range(want, -1, 0, 15, 0)
    for hs, he in d.holes[0]:
        k = np.searchsorted(gs, hs, side="right") - 1
        assert k >= 0 and gs[k] <= hs and ge[k] >= he


def test_narrow_fetch_widths(ctx):
    rng = np.random.default_rng(11)
    L = 10_007
    ctx.set_contigs([L, 33])
    for hi, want in ((200, np.uint8), (60_000, np.uint16), (70_000, np.int32)):
        d = rng.integers(0, hi + 1, L).astype(np.int32)
        d[-1] = hi
        ctx.load_depth(0, 0, d)
        ctx.load_depth(0, 1, np.arange(33, dtype=np.int32))
        got = ctx.fetch_depth_narrow(0, 0)
        assert got.dtype == want and np.array_equal(got.astype(np.int64), d)
        assert np.array_equal(ctx.fetch_depth_narrow(0, 1).astype(np.int64), np.arange(33))
    d = rng.integers(-5, 5, L).astype(np.int32)
    ctx.load_depth(0, 0, d)
    got = ctx.fetch_depth_narrow(0, 0)
    assert got.dtype == np.int32 and np.array_equal(got, d)


def test_gpu_gzip_members_decompress_to_reference_text(ctx):
    """.depth.gz produced on the GPU: gzip.decompress of the members == the reference's text stream"""
    import gzip
    rng = np.random.default_rng(5)
    L = 70_001
    d = np.zeros(L, np.int32)
    pos = 0
    while pos < L:                       # runs of equal values with every digit count, incl. 1-2 line runs
        run = int(rng.choice([1, 2, 3, 5, 86, 87, 129, 300, 5000]))
        d[pos:pos + run] = int(rng.choice([0, 7, 23, 100, 65535, 1234567, 2147483647, -5]))
        pos += run
    ctx.set_contigs([L, 5])
    ctx.load_depth(0, 0, d)
    ctx.load_depth(0, 1, np.array([1, 1, 2, 2, 2], np.int32))
    want = ">ctg one\n" + "".join(f"{v}\n" for v in d.tolist())
    got = ctx.depth_gzip(0, 0, header=b">ctg one\n").tobytes()
    assert gzip.decompress(got).decode() == want
    assert len(got) < len(want) / 20
    # sub-range without header, and a tiny contig
    got = ctx.depth_gzip(0, 0, 8190, 20_000).tobytes()
    assert gzip.decompress(got).decode() == "".join(f"{v}\n" for v in d[8190:28190].tolist())
    assert gzip.decompress(ctx.depth_gzip(0, 1, header=b">x\n").tobytes()) == b">x\n1\n1\n2\n2\n2\n"
    assert gzip.decompress(ctx.depth_gzip(0, 1, 0, 0, header=b">x\n").tobytes()) == b">x\n"


def test_gpu_gzip_whole_track_one_pass(ctx):
    """gci_depth_gzip_track: every selected contig in one pass (headers + members in contig order), per-contig byte
    offsets, an unselected contig in the middle, a zero-length contig, depth from the real pipeline."""
    import gzip
    lengths = [100_000, 40_000, 0, 8192, 8193, 30_000]
    sel = [True, False, True, True, True, True]
    d = synth.make_reads(synth.SynthSpec([100_000], coverage=25, seed=4, read_mean=6000, read_min=800, read_max=15000,
                                         hole_fraction=0.05))
    ctx.set_contigs(lengths, sel)
    ctx.reads_begin(d.n_reads)
    ctx.upload_bam(d.bam)
    ctx.filter()
    ctx.depth(0, 15, -1, 0)
    rng = np.random.default_rng(8)
    extra = {3: np.repeat(rng.integers(0, 3, 90), 100)[:8192].astype(np.int32),
             4: np.full(8193, 41, np.int32),
             5: np.repeat(rng.integers(-3, 2000, 300), 100).astype(np.int32)}
    for c, v in extra.items():
        ctx.load_depth(0, c, v)
    headers = [f">ctg{c} len={l}\n".encode() for c, l in enumerate(lengths)]
    blob, off = ctx.depth_gzip_track(0, headers)
    assert off[0] == 0 and off[-1] == len(blob) and off[1] == off[2]            # contig 1 is not selected
    depth0 = ctx.fetch_depth(0, 0)
    want = {0: depth0, 2: np.zeros(0, np.int32), **extra}
    whole = b""
    for c in (0, 2, 3, 4, 5):
        text = headers[c] + b"".join(b"%d\n" % int(v) for v in want[c])
        assert gzip.decompress(blob[off[c]:off[c + 1]].tobytes()) == text, c
        whole += text
    assert gzip.decompress(blob.tobytes()) == whole
    assert len(blob) < len(whole) / 15
    # the per-contig entry point produces the same members
    assert ctx.depth_gzip(0, 0, header=headers[0]).tobytes() == blob[off[0]:off[1]].tobytes()


def test_paf_election_known_cases(ctx):
    """PAF leg on the GPU: multi-block merge (touching blocks merge), longest target block with ties, equal scores
    decided by the contig NAME (not its index), query length of the first line, duplicates, two PAF files."""
    names = ["zeta", "alpha", "mid"]          # name order differs from index order
    lengths = [50_000, 50_000, 50_000]

    def line(read, ref, q0, q1, t0, t1, nmatch=None, alnlen=None, mapq=60, qlen=10_000):
        n = q1 - q0
        return dict(read_id=read, qlen=qlen, qstart=q0, qend=q1, ref_id=ref, tstart=t0, tend=t1,
                    nmatch=n if nmatch is None else nmatch, alnlen=n if alnlen is None else alnlen, mapq=mapq)

    p1 = PafTable.from_rows([
        # read 0: identical evidence on zeta (0) and alpha (1) -> equal scores -> 'zeta' > 'alpha' wins
        line(0, 0, 0, 5000, 1000, 6000), line(0, 1, 0, 5000, 2000, 7000),
        # read 1: three blocks on mid: [0,3000) and [3000,5000) touch -> merge; [7000,9000) separate;
        #         target blocks [100,3100),[3100,5100) merge to length 5000, [20000,22000) shorter
        line(1, 2, 0, 3000, 100, 3100), line(1, 2, 3000, 5000, 3100, 5100), line(1, 2, 7000, 9000, 20000, 22000),
        # read 2: two equally long target blocks -> the first in sorted order wins; duplicate line
        line(2, 1, 0, 2000, 30000, 32000), line(2, 1, 4000, 6000, 10000, 12000), line(2, 1, 4000, 6000, 10000, 12000),
        # read 3: low identity line dropped, low mapq line dropped, remaining one kept; qlen differs per line
        line(3, 0, 0, 4000, 500, 4500, nmatch=3000), line(3, 0, 0, 4000, 600, 4600, mapq=10),
        line(3, 2, 100, 4100, 700, 4700, qlen=8000),
        # read 4: only in the first PAF (leaks into the second one's table)
        line(4, 1, 0, 9000, 40000, 49000, mapq=40),
        # read 5: unknown contig
        line(5, -1, 0, 9000, 0, 9000),
    ])
    p2 = PafTable.from_rows([line(0, 1, 5000, 9000, 7000, 11000), line(1, 2, 0, 9000, 100, 9100, mapq=35),
                             line(6, 0, 0, 9000, 100, 9100, mapq=35)])
    rows = []
    for rid, (ref, st, ln) in enumerate([(0, 1000, 5000), (2, 100, 5000), (1, 10000, 2000), (2, 700, 4000),
                                         (1, 40000, 9000), (0, 0, 9000), (0, 100, 9000)]):
        rows.append(dict(ref_id=ref, ref_start=st, cigar=f"{ln}M", read_id=rid, mapq=45, qlen=ln))
    bam = _bam(rows)
    for pafs in ([p1], [p1, p2], [p2, p1]):
        _check(ctx, names, lengths, pafs, [bam], 8, params=dict(ovlp_percent=0.3))
    with pytest.raises(Exception):
        bad = PafTable.from_rows([line(0, 0, 0, 100, 0, 100, alnlen=0, nmatch=0)])
        _run_gpu(ctx, names, lengths, [bad], [bam], 8)


@pytest.mark.parametrize("seed", range(3))
def test_fused_pipeline_equals_staged_calls(ctx, seed):
    """gci_pipeline (one synchronisation) == gci_filter + gci_depth + gci_scan + gci_score_terms_sums, also when
    the interval buffers are too small for the result (the retry path)"""
    from gci_b200._lib import Context
    lengths = [120_000, 70_000, 30_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=[25, 3, 12][seed], seed=700 + seed, read_mean=5000,
                                         read_min=800, read_max=12000, hole_fraction=0.03, hole_mean=300))
    files = [d.bam] + ([synth.second_aligner(d, seed=seed)] if seed else [])
    pafs = [synth.aln_to_paf(synth.second_aligner(d, seed=40))] if seed == 2 else []
    c2 = Context(0)     # fresh context: first pipeline call runs with the minimum interval capacity
    try:
        for c in (ctx, c2):
            c.set_contigs(lengths)
            c.set_name_rank(_name_rank(d.contigs.names))
            c.reads_begin(d.n_reads)
            for t in pafs:
                c.upload_paf(t)
            for t in files:
                c.upload_bam(t)
        ts = 1
        n_surv = ctx.filter()
        ctx.depth(0, 15, -1, ts)
        n_iv = ctx.scan(0, -1, ts, 15)
        a50, actg, _, _, asum = ctx.score_terms(0, 3, n_iv, 0.005, 15, with_sums=True)
        for rep in range(2):
            got = c2.pipeline(0, 3, hi=ts)
            assert got[0] == n_surv and got[1] == n_iv
            assert (got[2] == a50).all() and (got[3] == actg).all() and (got[4] == asum).all()
            gs, ge, off = c2.fetch_intervals(0, 3)
            ws, we, woff = ctx.fetch_intervals(0, 3)
            assert (gs == ws).all() and (ge == we).all() and (off == woff).all()
            for i in range(3):
                assert np.array_equal(c2.fetch_depth(0, i), ctx.fetch_depth(0, i))
        # more issue intervals than the initial capacity (4096) of a fresh context: the fused call redoes the scan
        c3 = Context(0)
        try:
            n = 6000
            many = _bam([dict(ref_start=i * 100, cigar="60M", read_id=i) for i in range(n)])
            L = n * 100 + 50
            for c in (ctx, c3):
                c.set_contigs([L])
                c.reads_begin(n)
                c.upload_bam(many)
            ctx.filter()
            ctx.depth(0, 5, -1, 0)
            n_iv = ctx.scan(0, -1, 0, 5)
            assert n_iv > 4096
            a50, actg, _, _, asum = ctx.score_terms(0, 1, n_iv, 0.005, 5, with_sums=True)
            got = c3.pipeline(0, 1, flank_len=5, hi=0)
            assert got[1] == n_iv and (got[2] == a50).all() and (got[3] == actg).all() and (got[4] == asum).all()
            assert all((x == y).all() for x, y in zip(c3.fetch_intervals(0, 1), ctx.fetch_intervals(0, 1)))
            got = c3.pipeline(0, 1, flank_len=5, hi=0)      # second call: buffers are large enough now
            assert got[1] == n_iv and (got[2] == a50).all()
        finally:
            c3.close()
    finally:
        c2.close()


def test_refilter_same_read_set_with_other_cutoffs(ctx):
    """gci_filter twice on one uploaded read set with different gates: the second run must not inherit the
    high-quality set (or anything else) of the first"""
    lengths = [120_000, 60_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=20, seed=4242, read_mean=6000, read_min=1000, read_max=15000))
    b2 = synth.second_aligner(d, seed=3)
    paf = synth.aln_to_paf(synth.second_aligner(d, seed=8))
    ctx.set_contigs(lengths)
    ctx.set_name_rank(_name_rank(d.contigs.names))
    ctx.reads_begin(d.n_reads)
    ctx.upload_paf(paf)
    ctx.upload_bam(d.bam)
    ctx.upload_bam(b2)
    for p in (dict(mq_cutoff=20, map_qual=10, iden_percent=0.8), dict(mq_cutoff=60, map_qual=30, iden_percent=0.95),
              dict(mq_cutoff=35, ovlp_percent=0.5)):
        kw = dict(map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9)
        kw.update(p)
        n = ctx.filter(**kw)
        want_d, want_s = O.filter_depth([paf], [d.bam, b2], d.contigs.names, lengths, **kw)
        assert n == len(want_s), p
        ctx.depth(0, 15)
        for i in range(2):
            assert np.array_equal(ctx.fetch_depth(0, i).astype(np.int64), want_d[i]), p


@pytest.mark.parametrize("kind", ["long", "mixed", "edges", "short"])
def test_cigar_stats_match_op_sums(ctx, kind):
    """gci_fetch_cigar_stats (what pysam's get_cigar_stats / reference_end give the reference, GCI.py:157-166)
    against a numpy per-op sum: record lengths that put record ends on every 128- / 512- / 2048-op border of the
    streaming kernel, tiles mixing many short and a few long records, rare ops (N, S, H, P) in the middle."""
    rng = np.random.default_rng({"long": 1, "mixed": 2, "edges": 3, "short": 4}[kind])
    if kind == "long":
        n_ops = rng.integers(1500, 30_000, 60)
    elif kind == "mixed":
        n_ops = np.concatenate([rng.integers(0, 60, 700), rng.integers(1000, 9000, 40), rng.integers(100, 400, 60)])
        rng.shuffle(n_ops)
    elif kind == "edges":
        n_ops = rng.choice([0, 1, 2, 127, 128, 129, 383, 384, 511, 512, 513, 1023, 1024, 2047, 2048, 2049, 4096,
                            6143, 6144], 300)
    else:
        n_ops = rng.integers(1, 50, 4000)
    off = np.concatenate([[0], np.cumsum(n_ops)]).astype(np.uint64)
    c = int(off[-1])
    op = rng.choice([0, 0, 0, 0, 1, 2, 7, 8], c)
    rare = rng.random(c) < (0.002 if kind != "short" else 0.05)
    op[rare] = rng.choice([3, 4, 5, 6], int(rare.sum()))
    ln = rng.integers(1, 60, c)
    big = rng.random(c) < 0.001
    ln[big] = rng.integers(1 << 16, 1 << 20, int(big.sum()))          # lengths beyond 16 bits
    cigar = ((ln.astype(np.uint32) << 4) | op.astype(np.uint32)).astype(np.uint32)
    a = len(n_ops)
    t = AlnTable(np.zeros(a, np.int32), rng.integers(0, 1000, a).astype(np.int32), np.full(a, 60, np.uint8),
                 np.full(a, 4, np.uint16), np.zeros(a, np.int32), np.full(a, 100, np.int32),
                 np.arange(a, dtype=np.uint32), off, cigar)   # flag 0x4: the statistics are computed, the gates skipped
    ctx.set_contigs([2_000_000_000])
    ctx.reads_begin(a)
    ctx.upload_bam(t)
    assert ctx.filter() == 0
    got, got_end = ctx.fetch_cigar_stats(0, a)
    sums = t.op_sums()
    want = np.stack([sums[:, 0] + sums[:, 7] + sums[:, 8], sums[:, 1], sums[:, 2], sums[:, 3], sums[:, 4]], axis=1)
    assert np.array_equal(got.astype(np.int64), want & 0xFFFFFFFF)
    rlen = want[:, 0] + want[:, 2] + want[:, 3]
    assert np.array_equal(got_end.astype(np.int64), t.ref_start.astype(np.int64) + np.maximum(rlen, 1))


def test_pipeline_graph_replay_follows_the_data(ctx):
    """stage timers off: gci_pipeline records the step into a CUDA graph on its second run and replays it; the
    replay must see new record CONTENTS of the same shape, and any change of shape / arguments must drop it"""
    from gci_b200._lib import Context
    lengths = [150_000, 60_000]
    d = synth.make_reads(synth.SynthSpec(lengths, coverage=20, seed=909, read_mean=5000, read_min=800, read_max=12000,
                                         hole_fraction=0.02, hole_mean=300))
    b2 = synth.second_aligner(d, seed=4)
    rng = np.random.default_rng(5)

    def variant(t, k):
        if k == 0:
            return t
        if k == 3:
            # same record and op counts, other record borders: the CIGARs change hands, so the tile / span
            # lists built at upload differ while every size in the step's signature may stay the same
            perm = rng.permutation(t.n_records)
            off = t.cigar_off.astype(np.int64)
            n = (off[1:] - off[:-1])[perm]
            new_off = np.concatenate([[0], np.cumsum(n)])
            src = np.repeat(off[perm] - new_off[:-1], n) + np.arange(int(new_off[-1]), dtype=np.int64)
            return AlnTable(t.ref_id, t.ref_start, t.mapq, t.flag, np.full(t.n_records, 0, np.int32), t.qlen,
                            t.read_id, new_off.astype(np.uint64), t.cigar[src])
        mapq = t.mapq.copy()
        mapq[rng.random(len(mapq)) < 0.2 * k] = 5          # same shape, other survivors
        return AlnTable(t.ref_id, t.ref_start, mapq, t.flag, t.nm, t.qlen, t.read_id, t.cigar_off, t.cigar)

    def staged(files, ts, fl):
        ctx.set_contigs(lengths)
        ctx.reads_begin(d.n_reads)
        for t in files:
            ctx.upload_bam(t)
        n_surv = ctx.filter()
        ctx.depth(0, fl, -1, ts)
        n_iv = ctx.scan(0, -1, ts, fl)
        a50, actg, _, _, asum = ctx.score_terms(0, 2, n_iv, 0.005, fl, with_sums=True)
        return n_surv, n_iv, a50, actg, asum, [ctx.fetch_depth(0, i) for i in range(2)], ctx.fetch_intervals(0, 2)

    c = Context(0)
    try:
        c.set_timing(False)
        c.set_contigs(lengths)
        replays = 0
        for k in range(4):                      # four read sets of one shape, four steps each
            files = [variant(d.bam, k), b2]
            want = staged(files, 1, 15)
            for rep in range(4):
                c.reads_begin(d.n_reads)
                for t in files:
                    c.upload_bam(t)
                got = c.pipeline(0, 2, hi=1)
                assert got[0] == want[0] and got[1] == want[1], (k, rep)
                assert all((x == y).all() for x, y in zip(got[2:5], want[2:5])), (k, rep)
                assert all(np.array_equal(c.fetch_depth(0, i), want[5][i]) for i in range(2)), (k, rep)
                assert all((x == y).all() for x, y in zip(c.fetch_intervals(0, 2), want[6])), (k, rep)
        replays = c.graph_replays
        assert replays >= 12, replays           # everything after the first two steps ran as a graph
        # other arguments: eager again, then a new graph
        want = staged([variant(d.bam, 0), b2], 3, 20)
        c.reads_begin(d.n_reads)
        for t in (d.bam, b2):
            c.upload_bam(t)
        for rep in range(3):
            got = c.pipeline(0, 2, hi=3, flank_len=20)
            assert got[0] == want[0] and got[1] == want[1]
            assert all((x == y).all() for x, y in zip(got[2:5], want[2:5]))
        # other shape (one file instead of two)
        want = staged([d.bam], 1, 15)
        for rep in range(3):
            c.reads_begin(d.n_reads)
            c.upload_bam(d.bam)
            got = c.pipeline(0, 2, hi=1)
            assert got[0] == want[0] and got[1] == want[1]
            assert all(np.array_equal(c.fetch_depth(0, i), want[5][i]) for i in range(2))
        # an entry point that touches the track between two steps drops the graph, results stay right
        c.scan_windows(0, [0], [1000], [50_000], lo=-1, hi=1)
        got = c.pipeline(0, 2, hi=1)
        assert got[1] == want[1] and all((x == y).all() for x, y in zip(c.fetch_intervals(0, 2), want[6]))
        assert c.graph_replays > replays
    finally:
        c.close()
