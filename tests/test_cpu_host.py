"""CPU-only checks: the C-ABI library loads and exports every symbol of include/gci_cuda.h, fails loudly
without a GPU, and the host-side file formats round-trip."""
import os
import re

import numpy as np
import pytest

from gci_b200 import io as gio, synth
from gci_b200.records import AlnTable, PafTable, pack_cigar, unpack_cigar

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from gci_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib


def test_library_exports_every_header_symbol():
    _lib = _ensure_built()
    hdr = open(os.path.join(ROOT, "include", "gci_cuda.h")).read()
    declared = set(re.findall(r"\b(gci_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gci_cuda.h but not exported"
    assert declared == set(_lib.exported_symbols())
    assert lib.gci_version() >= 100


def test_home_mapping_host_matches_library():
    """the PAF dealing on the host (sharded.home_rank / home_local) and the kernels (gci_shard_home: the same inline
    functions the dispatch kernel uses) agree, and the home-local ids of every rank are 0 .. count-1 without holes"""
    import ctypes as C
    from gci_b200 import sharded
    lib = _ensure_built().load_library()
    ids = np.concatenate([np.arange(0, 3000), np.arange(2**31 - 200, 2**31 + 200), [2**32 - 1]]).astype(np.uint32)
    for world in (1, 2, 3, 8):
        hr, hl = sharded.home_rank(ids, world), sharded.home_local(ids, world)
        for q, r_want, l_want in zip(ids.tolist(), hr.tolist(), hl.tolist()):
            r, l = C.c_int32(-1), C.c_uint32(0)
            assert lib.gci_shard_home(q, world, C.byref(r), C.byref(l)) == 0
            assert (r.value, l.value) == (r_want, l_want), (q, world)
        n = 1000
        for rank in range(world):
            loc = np.sort(hl[:n][hr[:n] == rank])
            assert np.array_equal(loc, np.arange(len(loc))), (world, rank)
    assert lib.gci_shard_home(1, 0, None, None) != 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib = _ensure_built()
    with pytest.raises(_lib.GciError):
        _lib.Context(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gci_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_cigar_pack_roundtrip():
    for txt in ["10S100M", "500H100M5I3D2N7=8X", "*"]:
        assert unpack_cigar(pack_cigar(txt)) == txt


def test_bam_roundtrip(tmp_path):
    d = synth.make_reads(synth.SynthSpec([60_000, 30_000], coverage=8, seed=9, read_mean=4000, read_min=500,
                                         read_max=9000))
    p = str(tmp_path / "x.bam")
    gio.write_bam(p, d.contigs.names, d.contigs.lengths, d.bam)
    names, lengths = gio.read_bam_header(p)
    assert names == d.contigs.names and lengths == [int(x) for x in d.contigs.lengths]
    intern = {}
    _, _, t = gio.read_bam_py(p, intern)
    for col in ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "cigar_off", "cigar"):
        assert np.array_equal(getattr(t, col), getattr(d.bam, col)), col
    # interned ids follow first appearance; names map back to the synthetic ids
    back = {v: int(k.decode()[4:]) for k, v in intern.items()}
    assert [back[int(i)] for i in t.read_id] == [int(i) for i in d.bam.read_id]
    # the native decoder (libgci_io.so, parallel inflate) gives the same table
    for threads in (1, 4):
        _, _, tn = gio.read_bam(p, gio.NameTable(native=True), threads)
        for col in ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id", "cigar_off", "cigar"):
            assert np.array_equal(getattr(tn, col), getattr(t, col)), col


def test_bam_streaming_windows(tmp_path, monkeypatch):
    """the native decoder streams the file through a bounded window: with windows of a single BGZF block every
    other record is cut by a window border (carried tail), the header of a many-contig assembly is longer than
    a window, and the table still equals the pure-Python decode and the one-window decode"""
    n_ctg = 3000
    names = [f"scaffold_{i:05d}_of_a_fragmented_assembly" for i in range(n_ctg)]
    lengths = [2000 + i for i in range(n_ctg)]
    d = synth.make_reads(synth.SynthSpec([40_000, 25_000], coverage=10, seed=19, read_mean=3000, read_min=500,
                                         read_max=8000))
    tab = d.bam
    tab.ref_id[:] = (np.arange(tab.n_records) * 7) % n_ctg          # any contig of the big header
    p = str(tmp_path / "w.bam")
    gio.write_bam(p, names, lengths, tab)
    _, _, want = gio.read_bam_py(p, {})
    monkeypatch.delenv("GCI_IO_WINDOW_BYTES", raising=False)
    hn, hl, one = gio.read_bam(p, gio.NameTable(native=True), 3)
    assert hn == names and hl == lengths
    for window in ("1", "70000", "300000"):
        monkeypatch.setenv("GCI_IO_WINDOW_BYTES", window)
        for threads in (1, 4):
            hn, hl, got = gio.read_bam(p, gio.NameTable(native=True), threads)
            assert hn == names and hl == lengths
            for col in ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id", "cigar_off", "cigar"):
                assert np.array_equal(getattr(got, col), getattr(want, col)), (window, threads, col)
                assert np.array_equal(getattr(got, col), getattr(one, col)), (window, threads, col)


def test_paf_parallel_chunks(tmp_path, monkeypatch):
    """the native PAF parser cuts the file into one chunk per thread at line starts: same table and same read ids
    as the line-by-line Python reader for any thread count, and the first bad line is reported by its number"""
    rng = np.random.default_rng(3)
    n = 30_000
    reads = rng.integers(0, 9000, n)                               # repeated names: ids follow first appearance
    lines = []
    for i in range(n):
        ctg = ["chr1", "chr2", "unplaced"][int(rng.integers(0, 3))]
        s0 = int(rng.integers(0, 10**6))
        lines.append("\t".join(map(str, [f"read/{int(reads[i])}/ccs", 20000, 5, 19990, "+-"[i & 1], ctg, 10**7, s0,
                                          s0 + 19985, 19000 + i % 900, 19985, int(rng.integers(0, 61))])) +
                     ("\ttp:A:P\tcm:i:5" if i % 3 else ""))
    p = str(tmp_path / "big.paf")
    open(p, "w").write("\n".join(lines) + "\n")
    want = gio.read_paf_py(p, {"chr1": 0, "chr2": 1})
    for threads in ("1", "3", "8"):
        monkeypatch.setenv("GCI_IO_THREADS", threads)
        got = gio.read_paf(p, ["chr1", "chr2"], gio.NameTable(native=True))
        for col in ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq"):
            assert np.array_equal(getattr(got, col), getattr(want, col)), (threads, col)
    bad = lines[:]
    bad[25_000] = bad[25_000].replace("\t20000\t", "\t20k\t", 1)
    bad[29_000] = "too\tshort"
    open(p, "w").write("\n".join(bad))                            # and no newline at the end of the file
    for threads in ("1", "8"):
        monkeypatch.setenv("GCI_IO_THREADS", threads)
        with pytest.raises(ValueError, match="PAF line 25001: invalid integer"):
            gio.read_paf(p, ["chr1", "chr2"], gio.NameTable(native=True))
    # Python's int() accepts blanks around a field; a value outside int32 is reported, not truncated
    ok = lines[:50]
    ok[7] = ok[7].replace("\t20000\t", "\t 20000 \t", 1)
    open(p, "w").write("\n".join(ok) + "\n")
    want = gio.read_paf_py(p, {"chr1": 0, "chr2": 1})
    got = gio.read_paf(p, ["chr1", "chr2"], gio.NameTable(native=True))
    assert np.array_equal(got.qlen, want.qlen) and int(got.qlen[7]) == 20000
    ok[9] = ok[9].replace("\t20000\t", "\t3000000000\t", 1)
    open(p, "w").write("\n".join(ok) + "\n")
    with pytest.raises(ValueError, match="PAF line 10: integer outside the int32 range"):
        gio.read_paf(p, ["chr1", "chr2"], gio.NameTable(native=True))


def test_bam_long_cigar_cg_tag(tmp_path):
    ops = np.array([(3 << 4) | 0, (1 << 4) | 1] * 40000, dtype=np.uint32)     # 80000 ops > 65535
    t = AlnTable([0], [5], [60], [0], [40000], [160000], [0], np.array([0, len(ops)], np.uint64), ops)
    p = str(tmp_path / "long.bam")
    gio.write_bam(p, ["c"], [500000], t)
    for native in (False, True):
        _, _, back = gio.read_bam(p, gio.NameTable(native=native))
        assert np.array_equal(back.cigar, ops) and int(back.qlen[0]) == 160000 and int(back.nm[0]) == 40000


def test_paf_fasta_depth_roundtrip(tmp_path):
    d = synth.make_reads(synth.SynthSpec([50_000], coverage=5, seed=4, read_mean=4000, read_min=500, read_max=9000))
    paf = synth.aln_to_paf(d.bam)
    p = str(tmp_path / "x.paf")
    gio.write_paf(p, paf, d.contigs.names, d.contigs.lengths)
    for native in (False, True):
        back = gio.read_paf(p, ["chr1"], gio.NameTable(native=native))
        for col in ("qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq"):
            assert np.array_equal(getattr(back, col), getattr(paf, col)), (native, col)
    fa = str(tmp_path / "r.fa")
    gio.write_fasta(fa, ["chr1", "chr2"], [50_000, 3000], [[(100, 200), (4000, 4010)], []])
    for reader in (gio.read_fasta_gaps_py, gio.read_fasta_gaps):
        ids, gaps = reader(fa)
        assert ids == ["chr1", "chr2"] and gaps == {"chr1": [(100, 200), (4000, 4010)]}
    dz = str(tmp_path / "d.depth.gz")
    gio.write_depth_gz(dz, [b">c1\n", b"0\n1\n2\n", b"3\n", b">c2\n7\n"], threads=3)
    got = gio.read_depth_gz(dz)
    assert list(got) == ["c1", "c2"] and got["c1"].tolist() == [0, 1, 2, 3] and got["c2"].tolist() == [7]


@pytest.mark.parametrize("seed", range(6))
def test_fasta_scanner_on_messy_files(tmp_path, seed):
    """native line-based N-run scanner == the Python reader on CRLF / blank lines / trailing blanks / lower-case n /
    runs crossing line ends / unwrapped records longer than the read buffer / text before the first header /
    no final newline / gzip"""
    import gzip as _gz
    rng = np.random.default_rng(seed)
    out = [b"; a comment before the first record\n"] if seed % 2 else []
    n_rec = int(rng.integers(1, 6))
    for r in range(n_rec):
        L = int(rng.integers(0, 30_000)) if seed != 5 else int(rng.integers(5_000_000, 6_000_000))
        seq = rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), L)
        for _ in range(int(rng.integers(0, 6))):
            a = int(rng.integers(0, max(1, L)))
            b = min(L, a + int(rng.integers(1, 400)))
            seq[a:b] = rng.choice(np.frombuffer(b"Nn", np.uint8), b - a)
        if L and rng.random() < 0.5:
            seq[:int(rng.integers(1, 10))] = ord("N")                # run at the very start
        if L and rng.random() < 0.5:
            seq[L - int(rng.integers(1, min(10, L) + 1)):] = ord("n")   # and at the very end
        eol = b"\r\n" if seed == 1 else b"\n"
        out.append(b">" + (b"  " if seed == 2 else b"") + f"ctg{r}".encode() + (b" some description" if r % 2 else b"") + eol)
        width = 60 if seed != 3 else max(1, L)                         # seed 3: unwrapped
        txt = seq.tobytes()
        for i in range(0, L, width):
            out.append(txt[i:i + width] + (b" \t" if seed == 4 and i % 7 == 0 else b"") + eol)
            if seed == 4 and i % 11 == 0:
                out.append(eol)                                        # blank line inside a record
    data = b"".join(out)
    if seed % 3 == 0 and data.endswith(b"\n"):
        data = data[:-1]                                               # no newline at the end of the file
    plain, zipped = str(tmp_path / "m.fa"), str(tmp_path / "m.fa.gz")
    open(plain, "wb").write(data)
    with _gz.open(zipped, "wb") as f:
        f.write(data)
    want = gio.read_fasta_gaps_py(plain)
    assert gio.read_fasta_gaps(plain) == want
    assert gio.read_fasta_gaps(zipped) == want


def test_depth_gz_native_reader(tmp_path, monkeypatch):
    """native .depth.gz reader == the pure-Python one on multi-member files, tiny windows (lines cut by the window
    border), CRLF, a header holding several '>' (the reference keeps the text after the last one), no final
    newline; malformed lines are refused"""
    rng = np.random.default_rng(11)
    a = rng.integers(0, 3, 70_000)
    a[1000:1200] = rng.integers(90, 2_000_000, 200)                # several digits
    b = rng.integers(0, 60, 12_345)
    parts = [b">ctgA\n", "".join(f"{v}\n" for v in a[:30_000]).encode(), "".join(f"{v}\n" for v in a[30_000:]).encode(),
             b">old>ctgB\r\n", "".join(f"{v}\r\n" for v in b).encode()[:-2]]
    p = str(tmp_path / "x.depth.gz")
    gio.write_depth_gz(p, parts, threads=3)
    want = gio.read_depth_gz_py(p)
    assert list(want) == ["ctgA", "ctgB"] and np.array_equal(want["ctgA"], a) and np.array_equal(want["ctgB"], b)
    for window in (None, "1000", "70000"):
        if window:
            monkeypatch.setenv("GCI_IO_WINDOW_BYTES", window)
        for threads in (1, 5):
            got = gio.read_depth_gz(p, threads)
            assert list(got) == ["ctgA", "ctgB"]                  # item.split('>')[-1], utility/GCI_score.py:31
            assert np.array_equal(got["ctgA"], a) and np.array_equal(got["ctgB"], b)
    monkeypatch.delenv("GCI_IO_WINDOW_BYTES", raising=False)
    for bad in (b">c\n1\nx2\n", b"5\n>c\n1\n", b">c\n1\n99999999999\n"):
        gio.write_depth_gz(p, [bad], threads=1)
        with pytest.raises(ValueError):
            gio.read_depth_gz(p)
    gio.write_depth_gz(p, [b">empty\n", b">c\n7\n"], threads=1)
    got = gio.read_depth_gz(p)
    assert list(got) == ["empty", "c"] and len(got["empty"]) == 0 and got["c"].tolist() == [7]


@pytest.mark.skipif(not os.path.exists("/root/reference/example/MH63.depth.gz"),
                    reason="the reference's example is only mounted in the build container")
def test_depth_gz_native_reader_on_the_reference_example():
    """the reference's own example/MH63.depth.gz (395 765 488 depth lines) through the native reader equals the
    committed run-length fixture of the same file"""
    from helpers import mh63_depths
    names, lengths, depths = mh63_depths()
    got = gio.read_depth_gz("/root/reference/example/MH63.depth.gz")
    assert list(got) == names
    for n, d in zip(names, depths):
        assert np.array_equal(got[n], d), n


def test_native_io_exports_every_header_symbol():
    from gci_b200 import io_native
    hdr = open(os.path.join(ROOT, "include", "gci_io.h")).read()
    declared = set(re.findall(r"\b(gci_[a-z0-9_]+)\s*\(", hdr))
    L = io_native.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(io_native.exported_symbols())


def test_native_nm_types_and_missing(tmp_path):
    """NM stored as c / C / s / S / i / I and a record without NM"""
    import struct
    names, lengths = ["c"], [1000]
    t = AlnTable([0] * 3, [10, 20, 30], [60] * 3, [0] * 3, [5, 300, -(2 ** 31)], [50] * 3, [0, 1, 2],
                 np.array([0, 1, 2, 3], np.uint64), np.array([(50 << 4)] * 3, np.uint32))
    p = str(tmp_path / "nm.bam")
    gio.write_bam(p, names, lengths, t)
    for native in (False, True):
        _, _, back = gio.read_bam(p, gio.NameTable(native=native))
        assert back.nm.tolist() == [5, 300, -(2 ** 31)]


def test_host_n50_matches_oracle():
    from gci_b200.pipeline import compute_n50
    from oracle import gci_oracle as O
    rng = np.random.default_rng(0)
    for _ in range(200):
        v = rng.integers(1, 50, int(rng.integers(0, 12))).tolist()
        assert compute_n50(v) == O.n50(v)


def test_name_table_is_shared_between_files(tmp_path):
    """BAM and PAF of one read type must intern names into the same table (ids line up across files)"""
    d = synth.make_reads(synth.SynthSpec([40_000], coverage=6, seed=2, read_mean=3000, read_min=500, read_max=8000))
    bam = synth.drop_reads(d.bam, 0.3, 1)
    paf = synth.aln_to_paf(d.bam)
    gio.write_bam(str(tmp_path / "a.bam"), d.contigs.names, d.contigs.lengths, bam)
    gio.write_paf(str(tmp_path / "a.paf"), paf, d.contigs.names, d.contigs.lengths)
    for native in (False, True):
        nt = gio.NameTable(native=native)
        _, _, b = gio.read_bam(str(tmp_path / "a.bam"), nt)
        n_after_bam = len(nt)
        p = gio.read_paf(str(tmp_path / "a.paf"), d.contigs.names, nt)
        assert n_after_bam == len(np.unique(bam.read_id))
        assert len(nt) == len(np.union1d(bam.read_id, paf.read_id))
        # the same read gets the same id in both files: synthetic name -> id maps agree
        m_b = dict(zip(bam.read_id.tolist(), b.read_id.tolist()))
        m_p = dict(zip(paf.read_id.tolist(), p.read_id.tolist()))
        common = set(m_b) & set(m_p)
        assert len(common) > 20 and all(m_p[k] == m_b[k] for k in common)


def test_later_bam_with_permuted_sq_order_is_remapped_by_name():
    """the reference addresses contigs by NAME (fetch(contig=target), GCI.py:150-151); a later BAM whose
    @SQ order differs must land on the first BAM's contig indices, unknown contigs on -1."""
    from gci_b200 import pipeline as P
    a = AlnTable.from_rows([dict(ref_id=0, ref_start=5, mapq=60, flag=0, nm=0, qlen=10, read_id=0, cigar="10M"),
                            dict(ref_id=1, ref_start=7, mapq=60, flag=0, nm=0, qlen=10, read_id=1, cigar="10M")])
    a.contig_names, a.contig_lengths = ["chr1", "chr2"], [100, 200]
    b = AlnTable.from_rows([dict(ref_id=0, ref_start=5, mapq=60, flag=0, nm=0, qlen=10, read_id=1, cigar="10M"),
                            dict(ref_id=1, ref_start=7, mapq=60, flag=0, nm=0, qlen=10, read_id=0, cigar="10M"),
                            dict(ref_id=2, ref_start=1, mapq=60, flag=0, nm=0, qlen=10, read_id=2, cigar="10M"),
                            dict(ref_id=-1, ref_start=0, mapq=0, flag=4, nm=0, qlen=10, read_id=3, cigar="*")])
    b.contig_names, b.contig_lengths = ["chr2", "chr1", "chrUn"], [200, 100, 50]
    names, lengths, pafs, bams, n_reads = P._load_inputs([], [a, b])
    assert names == ["chr1", "chr2"] and lengths == [100, 200] and n_reads == 4
    assert bams[0].ref_id.tolist() == [0, 1]
    assert bams[1].ref_id.tolist() == [1, 0, -1, -1]
    assert bams[1].ref_start.tolist() == [5, 7, 1, 0]


def test_records_the_gates_drop_first_are_pruned_before_upload():
    """unmapped / secondary / supplementary / below -mq / never-fetched-contig records are dropped on the
    host (GCI.py:151-156 skips them before it reads anything else); their CIGARs go with them; a table that keeps
    almost everything is passed on untouched"""
    from gci_b200 import pipeline as P
    rows = []
    for i in range(40):
        flag = [0, 16, 0x100, 0x800, 4, 0x900][i % 6]
        rows.append(dict(ref_id=[0, 1, 2, -1][i % 4], ref_start=i, mapq=[60, 10][i % 2 if i % 5 == 0 else 0], flag=flag,
                         nm=None if i % 7 == 0 else 1, qlen=10 + i, read_id=i, cigar=f"{5 + i}M{1 + i % 3}I2S"))
    t = AlnTable.from_rows(rows)
    sel = [True, False, True]
    got = P._prune_for_gates(t, sel, 30)
    want = [i for i, r in enumerate(rows) if r["ref_id"] in (0, 2) and not (r["flag"] & 0x904) and r["mapq"] >= 30]
    assert 0 < len(want) < 36 and got.read_id.tolist() == want
    for k, i in enumerate(want):
        a, b = int(t.cigar_off[i]), int(t.cigar_off[i + 1])
        c, d = int(got.cigar_off[k]), int(got.cigar_off[k + 1])
        assert np.array_equal(got.cigar[c:d], t.cigar[a:b])
        assert int(got.nm[k]) == int(t.nm[i]) and int(got.qlen[k]) == rows[i]["qlen"]
    clean = AlnTable.from_rows([dict(ref_id=0, ref_start=i, mapq=60, flag=0, nm=0, qlen=10, read_id=i, cigar="10M")
                                for i in range(20)])
    assert P._prune_for_gates(clean, [True], 30) is clean
    empty = clean.take(np.zeros(0, np.int64))
    assert P._prune_for_gates(empty, [True], 30).n_records == 0
