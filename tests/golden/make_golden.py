#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (yeeus/GCI @ 455e19c7).

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What it does
  * imports /root/reference/GCI.py with `pysam`, `Bio.SeqIO` and `matplotlib`
    replaced by the thin shims below (the recipe of SURVEY.md §4.3).  The pysam
    shim serves records from in-memory `AlnTable`s through exactly the surface
    GCI.py touches (GCI.py:150-168, :201-208, :963-976); PAF and FASTA inputs are
    real text files;
  * runs the reference's own `GCI(**args)` driver on seeded synthetic inputs and
    stores inputs + every output file (depth streams as arrays, BED / .gci /
    regions.gci / gaps.bed as text) in `filter_cases.npz`;
  * derives `mh63_depth_rle.npz` (run-length encoding of the reference's
    `example/MH63.depth.gz`) and copies the two small example outputs
    `MH63.0.depth.bed`, `MH63.gci` next to it.

Nothing here is imported by the product; tests only read the generated files.
"""
from __future__ import annotations

import gzip
import io
import json
import os
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gci_b200.records import AlnTable, PafTable  # noqa: E402
from gci_b200 import synth  # noqa: E402

REF = "/root/reference"

# ------------------------------------------------------------------------------------
# shims
# ------------------------------------------------------------------------------------
BAM_REGISTRY: dict = {}   # path -> (AlnTable, names, lengths, read_name_fn)


class _Segment:
    __slots__ = ("t", "i", "names")

    def __init__(self, t, i, names):
        self.t, self.i, self.names = t, i, names

    @property
    def is_mapped(self):
        return not (int(self.t.flag[self.i]) & 0x4)

    @property
    def is_unmapped(self):
        return bool(int(self.t.flag[self.i]) & 0x4)

    @property
    def is_secondary(self):
        return bool(int(self.t.flag[self.i]) & 0x100)

    @property
    def is_supplementary(self):
        return bool(int(self.t.flag[self.i]) & 0x800)

    @property
    def mapping_quality(self):
        return int(self.t.mapq[self.i])

    def _ops(self):
        o = self.t.cigar_off
        return self.t.cigar[int(o[self.i]):int(o[self.i + 1])]

    def get_cigar_stats(self):
        base = [0] * 11
        blocks = [0] * 11
        for op in self._ops():
            base[int(op) & 15] += int(op) >> 4
            blocks[int(op) & 15] += 1
        nm = int(self.t.nm[self.i])
        if nm != -(2 ** 31):
            base[10] = nm
        return base, blocks

    def get_tag(self, tag):
        assert tag == "NM"
        nm = int(self.t.nm[self.i])
        if nm == -(2 ** 31):
            raise KeyError("tag 'NM' not present")
        return nm

    @property
    def query_name(self):
        return f"read{int(self.t.read_id[self.i])}"

    @property
    def reference_name(self):
        return self.names[int(self.t.ref_id[self.i])]

    @property
    def reference_start(self):
        return int(self.t.ref_start[self.i])

    @property
    def reference_end(self):
        if self.is_unmapped or len(self._ops()) == 0:
            return None
        rlen = sum(int(op) >> 4 for op in self._ops() if (int(op) & 15) in (0, 2, 3, 7, 8))
        return self.reference_start + (rlen if rlen else 1)

    @property
    def query_length(self):
        return int(self.t.qlen[self.i])


class AlignmentFile:
    def __init__(self, path, mode="rb", threads=1):
        self.t, self.references, self.lengths = BAM_REGISTRY[os.path.abspath(path)]
        self.references = tuple(self.references)
        self.lengths = tuple(int(x) for x in self.lengths)
        self._ends = self.t.ref_start.astype(np.int64) + np.maximum(1, self.t.ref_len())

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def fetch(self, contig=None, start=None, stop=None, multiple_iterators=False):
        rid = self.references.index(contig)
        t = self.t
        sel = np.flatnonzero((t.ref_id == rid) & (t.ref_start < stop) & (self._ends > start))
        for i in sel:
            yield _Segment(t, int(i), self.references)


class _Rec:
    def __init__(self, rid, seq):
        self.id, self.seq = rid, seq


def _fasta_parse(path, fmt):
    assert fmt == "fasta"
    rid, chunks = None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if rid is not None:
                    yield _Rec(rid, "".join(chunks))
                rid, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.strip())
    if rid is not None:
        yield _Rec(rid, "".join(chunks))


def import_reference():
    """The unmodified reference as a module (under its own name: the repo's drop-in CLI is a `GCI` module too).
    The shims only have to be in sys.modules while the reference's `import` statements run."""
    import importlib.util
    shim_names = ["pysam", "Bio", "Bio.SeqIO", "matplotlib", "matplotlib.pyplot", "matplotlib.lines", "matplotlib.ticker"]
    saved = {n: sys.modules.get(n) for n in shim_names}
    for n in shim_names:
        sys.modules[n] = types.ModuleType(n)
    sys.modules["pysam"].AlignmentFile = AlignmentFile
    sys.modules["Bio.SeqIO"].parse = _fasta_parse
    sys.modules["Bio"].SeqIO = sys.modules["Bio.SeqIO"]
    sys.modules["matplotlib.ticker"].AutoMinorLocator = object
    try:
        spec = importlib.util.spec_from_file_location("gci_reference_unmodified", os.path.join(REF, "GCI.py"))
        ref = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = ref   # multiprocessing pickles read_sam by module name (GCI.py:267-270)
        spec.loader.exec_module(ref)   # the unmodified reference, read where it lies
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return ref


# ------------------------------------------------------------------------------------
# inputs
# ------------------------------------------------------------------------------------

def write_paf(path, tab: PafTable, names, lengths):
    with open(path, "w") as f:
        for i in range(tab.n_records):
            t = int(tab.ref_id[i])
            f.write("\t".join(map(str, [
                f"read{int(tab.read_id[i])}", int(tab.qlen[i]), int(tab.qstart[i]), int(tab.qend[i]), "+",
                names[t], int(lengths[t]), int(tab.tstart[i]), int(tab.tend[i]), int(tab.nmatch[i]),
                int(tab.alnlen[i]), int(tab.mapq[i])])) + "\ttp:A:P\n")


def write_fasta(path, names, lengths, n_runs, seed=1):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for c, (n, l) in enumerate(zip(names, lengths)):
            seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(l))
            for k, (s, e) in enumerate(n_runs[c] or ()):
                seq[s:e] = ord("n") if k % 2 else ord("N")
            f.write(f">{n} synthetic contig {c}\n")
            txt = seq.tobytes().decode()
            for i in range(0, len(txt), 80):
                f.write(txt[i:i + 80] + "\n")


def parse_depth_stream(path):
    """decompressed .depth.gz -> {name: int64 array} in file order."""
    data = gzip.open(path, "rb").read()
    out = {}
    pos = 0
    parts = data.split(b">")
    for p in parts[1:]:
        nl = p.index(b"\n")
        name = p[:nl].decode()
        body = p[nl + 1:]
        out[name] = np.array(body.split(), dtype=np.int64) if body.strip() else np.zeros(0, np.int64)
    return out


def tab_to_dict(prefix, t, store):
    cols = (("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id", "cigar_off", "cigar")
            if isinstance(t, AlnTable) else
            ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq"))
    for c in cols:
        store[f"{prefix}.{c}"] = getattr(t, c)


def make_case(ref, name, store, meta, *, lengths, seed, hifi=("bam",), nano=None, args=None, chrs=None,
              regions=None, threads=1, coverage=30.0):
    """hifi / nano: tuples of 'bam' | 'paf' giving the files of that read type in CLI order."""
    args = dict(args or {})
    work = tempfile.mkdtemp(prefix=f"gci_case_{name}_")
    names = [f"ctg{chr(65 + i)}" for i in range(len(lengths))]
    case = {"name": name, "lengths": [int(x) for x in lengths], "names": names, "args": args, "chrs": chrs,
            "regions": regions, "threads": threads, "files": {}}
    cli = {}
    n_runs = None
    for rtype, kinds in (("hifi", hifi), ("nano", nano)):
        if not kinds:
            continue
        if rtype == "hifi":
            spec = synth.SynthSpec(lengths, coverage=coverage, seed=seed, contig_names=names,
                                   read_mean=6000, read_min=1500, read_max=15000, hole_mean=1200,
                                   hole_fraction=0.02)
        else:
            spec = synth.SynthSpec(lengths, coverage=coverage, seed=seed + 1000, contig_names=names,
                                   read_mean=9000, read_min=1500, read_max=25000, read_sigma=0.6,
                                   events_per_base=0.04, hole_mean=800, hole_fraction=0.01)
        d = synth.make_reads(spec)
        if n_runs is None:
            n_runs = d.n_runs
        paths = []
        case["files"][rtype] = []
        for k, kind in enumerate(kinds):
            if k == 0:
                bam = synth.drop_reads(d.bam, 0.02, seed + 5)
            else:
                bam = synth.second_aligner(d, seed=seed + 11 * k)
            key = f"{name}.{rtype}{k}"
            if kind == "bam":
                p = os.path.join(work, f"{rtype}{k}.bam")
                BAM_REGISTRY[os.path.abspath(p)] = (bam, names, lengths)
                open(p, "wb").close()
                tab_to_dict(key, bam, store)
            else:
                p = os.path.join(work, f"{rtype}{k}.paf")
                paf = synth.aln_to_paf(bam)
                write_paf(p, paf, names, lengths)
                tab_to_dict(key, paf, store)
            case["files"][rtype].append({"kind": kind, "key": key})
            paths.append(p)
        cli[rtype] = paths
    fasta = os.path.join(work, "ref.fa")
    write_fasta(fasta, names, lengths, n_runs)
    case["n_runs"] = [[list(map(int, iv)) for iv in (r or [])] for r in n_runs]
    reg_path = None
    if regions:
        reg_path = os.path.join(work, "regions.bed")
        with open(reg_path, "w") as f:
            for c, s, e in regions:
                f.write(f"{c}\t{s}\t{e}\n")
    out_dir = os.path.join(work, "out")
    kw = dict(hifi=cli.get("hifi"), nano=cli.get("nano"), directory=out_dir, prefix="T", reference=fasta,
              regions=reg_path, chrs=chrs, threads=threads, force=True)
    kw.update(args)
    saved = sys.stdout
    sys.stdout = io.StringIO()
    try:
        ref.GCI(**kw)
    finally:
        log = sys.stdout.getvalue()
        sys.stdout = saved
    outs = {}
    for fn in sorted(os.listdir(out_dir)):
        p = os.path.join(out_dir, fn)
        if fn.endswith(".depth.gz"):
            dd = parse_depth_stream(p)
            outs[fn] = {"contigs": list(dd.keys())}
            for cn, arr in dd.items():
                store[f"{name}.out.{fn}.{cn}"] = arr.astype(np.int32)
        else:
            outs[fn] = {"text": open(p).read()}
    case["outputs"] = outs
    case["stdout"] = log.replace(work, "<WORK>")
    meta["cases"].append(case)
    shutil.rmtree(work)
    print(f"case {name}: outputs {list(outs)}")


def make_filter_cases(ref):
    store, meta = {}, {"cases": [], "reference_commit": "455e19c7"}
    L3 = [120_000, 80_000, 50_000]
    make_case(ref, "hifi_bam", store, meta, lengths=L3, seed=101, hifi=("bam",))
    make_case(ref, "hifi_bam_t4", store, meta, lengths=L3, seed=101, hifi=("bam",), threads=4)
    make_case(ref, "hifi_bam_paf", store, meta, lengths=L3, seed=102, hifi=("bam", "paf"))
    make_case(ref, "hifi_paf_bam_bam", store, meta, lengths=L3, seed=103, hifi=("paf", "bam", "bam"))
    make_case(ref, "hifi_2paf_bam", store, meta, lengths=L3, seed=104, hifi=("bam", "paf", "paf"))
    make_case(ref, "dual", store, meta, lengths=L3, seed=105, hifi=("bam", "paf"), nano=("bam", "bam"))
    make_case(ref, "dual_args", store, meta, lengths=L3, seed=106, hifi=("bam", "bam"), nano=("bam", "paf"),
              args=dict(map_qual=20, mq_cutoff=40, iden_percent=0.95, clip_percent=0.05, ovlp_percent=0.8,
                        flank_len=10, threshold=3, dist_percent=0.01))
    make_case(ref, "chrs_regions", store, meta, lengths=L3, seed=107, hifi=("bam", "paf"), nano=("bam",),
              chrs="ctgA,ctgC",
              regions=[("ctgA", 1000, 60000), ("ctgC", 0, 50000), ("ctgA", 70000, 119000), ("ctgC", 20000, 20000)])
    make_case(ref, "nano_only_ts", store, meta, lengths=L3, seed=108, hifi=None, nano=("bam",), args=dict(threshold=5, flank_len=0))
    np.savez_compressed(os.path.join(HERE, "filter_cases.npz"), **store)
    with open(os.path.join(HERE, "filter_cases.json"), "w") as f:
        json.dump(meta, f, indent=1)


def make_random_cases(ref, n_cases=16, seed=20240633, prefix="rand", save=True):
    """Seeded random small cases over the whole argument space (gates, flank, threshold, -dp, file mixes, --chrs,
    -R regions, contig lengths next to tile borders and below 2 * flank) -> filter_cases_rand.{npz,json}."""
    store, meta = {}, {"cases": [], "reference_commit": "455e19c7"}
    rng = np.random.default_rng(seed)
    combos = [("bam",), ("bam", "bam"), ("bam", "paf"), ("paf", "bam"), ("bam", "paf", "bam"), ("bam", "bam", "bam"),
              ("paf", "paf", "bam")]
    special = [1023, 1024, 1025, 2047, 2048, 2049, 3072, 4000, 40, 90]
    for k in range(n_cases):
        n_ctg = int(rng.integers(2, 5))
        lengths = [int(rng.integers(6_000, 45_000)) for _ in range(n_ctg)]
        lengths[int(rng.integers(0, n_ctg))] = int(rng.choice(special)) + (5000 if rng.random() < 0.5 else 0)
        args = dict(map_qual=int(rng.choice([0, 10, 30, 45])), mq_cutoff=int(rng.choice([20, 50, 60])),
                    iden_percent=float(rng.choice([0.8, 0.9, 0.99, 0.999])),
                    clip_percent=float(rng.choice([0.0, 0.05, 0.1, 0.3])),
                    ovlp_percent=float(rng.choice([0.0, 0.5, 0.9, 1.0])),
                    flank_len=int(rng.choice([0, 1, 15, 15, 40, 200])), threshold=int(rng.choice([0, 0, 1, 3, 10])),
                    dist_percent=float(rng.choice([0.0, 0.001, 0.005, 0.05, 0.5])))
        hifi = combos[int(rng.integers(0, len(combos)))] if rng.random() < 0.85 else None
        nano = combos[int(rng.integers(0, len(combos)))] if (hifi is None or rng.random() < 0.4) else None
        names = [f"ctg{chr(65 + i)}" for i in range(n_ctg)]
        chrs = None
        if rng.random() < 0.3:
            keep = sorted(rng.choice(n_ctg, int(rng.integers(1, n_ctg + 1)), replace=False).tolist())
            chrs = ",".join(names[i] for i in keep)
        regions = None
        if rng.random() < 0.4:
            ok = [i for i in range(n_ctg) if chrs is None or names[i] in chrs.split(",")]
            regions = []
            for _ in range(int(rng.integers(1, 5))):
                c = int(rng.choice(ok))
                a, b = sorted(int(x) for x in rng.integers(0, lengths[c] + 1, 2))
                regions.append((names[c], a, b))
        try:
            make_case(ref, f"{prefix}{k:02d}", store, meta, lengths=lengths, seed=(0 if seed == 20240633 else seed % 100000) + 500 + k, hifi=hifi, nano=nano, args=args,
                      chrs=chrs, regions=regions, coverage=float(rng.choice([6.0, 12.0, 25.0])))
        except (Exception, SystemExit) as e:   # the reference itself gives up on this input: not a parity case
            print(f"case {prefix}{k:02d}: reference raised {type(e).__name__}: {e}")
            for key in [x for x in store if x.startswith(f"{prefix}{k:02d}.")]:
                del store[key]
    if save:
        np.savez_compressed(os.path.join(HERE, "filter_cases_rand.npz"), **store)
        with open(os.path.join(HERE, "filter_cases_rand.json"), "w") as f:
            json.dump(meta, f, indent=1)
    return store, meta


def make_sliding_window_kats(ref, n_cases=24):
    """known answers of the reference's sliding_window_average_depth (GCI.py:660-705) -> sliding_window_kat.json"""
    import contextlib
    rng = np.random.default_rng(660705)
    cases = []
    for k in range(n_cases):
        n = int(rng.integers(0, 300))
        w = int(rng.choice([1, 2, 3, 7, 50, 1000]))
        d = rng.integers(1, 40, n)
        d[rng.random(n) < rng.choice([0.0, 0.05, 0.3, 0.9])] = 0
        md = [3, 10.5, 100.0, 7][k % 4]
        st = int(rng.integers(0, 10**7))
        err = io.StringIO()
        with contextlib.redirect_stderr(err):
            pos, val = ref.sliding_window_average_depth(d.tolist(), w, md, st, "ctgA")
        cases.append({"depths": d.tolist(), "window_size": w, "max_depth": md, "start": st, "target": "ctgA",
                      "positions": pos, "values": val.tolist(), "dtype": str(val.dtype), "stderr": err.getvalue()})
    with open(os.path.join(HERE, "sliding_window_kat.json"), "w") as f:
        json.dump({"reference_commit": "455e19c7", "cases": cases}, f)
    print("sliding window KATs:", len(cases))


def make_mh63():
    data = gzip.open(os.path.join(REF, "example", "MH63.depth.gz"), "rb").read()
    names, vals, runs, lens = [], [], [], []
    for p in data.split(b">")[1:]:
        nl = p.index(b"\n")
        names.append(p[:nl].decode())
        d = np.array(p[nl + 1:].split(), dtype=np.int64)
        lens.append(len(d))
        brk = np.flatnonzero(np.diff(d)) + 1
        st = np.concatenate([[0], brk])
        vals.append(d[st].astype(np.int32))
        runs.append(np.diff(np.concatenate([st, [len(d)]])).astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "mh63_depth_rle.npz"), names=np.array(names), lengths=np.array(lens),
                        run_counts=np.array([len(v) for v in vals]), values=np.concatenate(vals),
                        run_lengths=np.concatenate(runs))
    shutil.copy(os.path.join(REF, "example", "MH63.0.depth.bed"), os.path.join(HERE, "mh63.0.depth.bed"))
    shutil.copy(os.path.join(REF, "example", "MH63.gci"), os.path.join(HERE, "mh63.gci"))
    print("mh63:", names, sum(lens), sum(len(v) for v in vals), "runs")


if __name__ == "__main__":
    ref = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "rand":
        make_random_cases(ref)
    elif len(sys.argv) > 1 and sys.argv[1] == "window":
        make_sliding_window_kats(ref)
    else:
        make_sliding_window_kats(ref)
        make_filter_cases(ref)
        make_random_cases(ref)
        make_mh63()
