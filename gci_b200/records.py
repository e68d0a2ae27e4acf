"""Columnar alignment records — the layout the GPU owns.

One `AlnTable` per BAM file and one `PafTable` per PAF file.  The columns are
exactly what the reference touches per record (GCI.py:153-166 for BAM,
GCI.py:218-229 for PAF); everything else in the file is dropped at decode time.

Read names are interned on the host into dense `read_id`s shared by all files
of one read type (exact — no hash collisions), and contig names
into `ref_id`s in BAM-header order of the first BAM (GCI.py:201-207).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import re

import numpy as np

# BAM CIGAR op codes (SAM spec §4.2): MIDNSHP=XB
CIGAR_OPS = "MIDNSHP=XB"
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X, OP_B = range(10)

NM_MISSING = np.int32(-(2**31))  # record has no NM aux tag (reference raises KeyError, GCI.py:163)

FLAG_UNMAP = 0x4
FLAG_SECONDARY = 0x100
FLAG_SUPPLEMENTARY = 0x800

_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=XB])")


def pack_cigar(text: str) -> np.ndarray:
    """SAM CIGAR text -> BAM packed u32 ops (`len << 4 | op`)."""
    if text == "*" or text == "":
        return np.zeros(0, dtype=np.uint32)
    out = [(int(n) << 4) | CIGAR_OPS.index(c) for n, c in _CIGAR_RE.findall(text)]
    return np.asarray(out, dtype=np.uint32)


def unpack_cigar(ops: np.ndarray) -> str:
    return "".join(f"{int(o) >> 4}{CIGAR_OPS[int(o) & 15]}" for o in ops) or "*"


@dataclass
class AlnTable:
    """Records of one BAM file, in file order (coordinate sorted)."""

    ref_id: np.ndarray      # i32 [A]   index into the contig table, -1 = none / not selected
    ref_start: np.ndarray   # i32 [A]   0-based leftmost position (BAM `pos`)
    mapq: np.ndarray        # u8  [A]
    flag: np.ndarray        # u16 [A]
    nm: np.ndarray          # i32 [A]   NM aux value, NM_MISSING if absent
    qlen: np.ndarray        # i32 [A]   BAM `l_seq` (pysam query_length)
    read_id: np.ndarray     # u32 [A]   dense interned read name
    cigar_off: np.ndarray   # u64 [A+1] offsets into `cigar`
    cigar: np.ndarray       # u32 [C]   packed ops
    kind: str = "bam"

    def __post_init__(self):
        self.ref_id = np.ascontiguousarray(self.ref_id, dtype=np.int32)
        self.ref_start = np.ascontiguousarray(self.ref_start, dtype=np.int32)
        self.mapq = np.ascontiguousarray(self.mapq, dtype=np.uint8)
        self.flag = np.ascontiguousarray(self.flag, dtype=np.uint16)
        self.nm = np.ascontiguousarray(self.nm, dtype=np.int32)
        self.qlen = np.ascontiguousarray(self.qlen, dtype=np.int32)
        self.read_id = np.ascontiguousarray(self.read_id, dtype=np.uint32)
        self.cigar_off = np.ascontiguousarray(self.cigar_off, dtype=np.uint64)
        self.cigar = np.ascontiguousarray(self.cigar, dtype=np.uint32)
        a = len(self.ref_id)
        for name in ("ref_start", "mapq", "flag", "nm", "qlen", "read_id"):
            if len(getattr(self, name)) != a:
                raise ValueError(f"AlnTable column {name} has wrong length")
        if len(self.cigar_off) != a + 1 or (a + 1 and int(self.cigar_off[-1]) != len(self.cigar)):
            raise ValueError("AlnTable cigar_off does not describe cigar")

    @property
    def n_records(self) -> int:
        return len(self.ref_id)

    @property
    def n_ops(self) -> int:
        return len(self.cigar)

    def nbytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in
                   ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen", "read_id", "cigar_off", "cigar"))

    def ref_len(self) -> np.ndarray:
        """Reference bases consumed per record (M, D, N, =, X)."""
        op = self.cigar & 15
        ln = (self.cigar >> 4).astype(np.int64)
        consumes = (op == OP_M) | (op == OP_D) | (op == OP_N) | (op == OP_EQ) | (op == OP_X)
        w = np.where(consumes, ln, 0)
        cs = np.concatenate([[0], np.cumsum(w)])
        off = self.cigar_off.astype(np.int64)
        return cs[off[1:]] - cs[off[:-1]]

    def op_sums(self) -> np.ndarray:
        """[A,10] int64 bases per CIGAR op code (host-side helper for generators / decoders)."""
        a = self.n_records
        out = np.zeros((a, 10), dtype=np.int64)
        if len(self.cigar):
            off = self.cigar_off.astype(np.int64)
            rec = np.repeat(np.arange(a, dtype=np.int64), off[1:] - off[:-1])
            np.add.at(out, (rec, (self.cigar & 15).astype(np.int64)), (self.cigar >> 4).astype(np.int64))
        return out

    def take(self, idx: np.ndarray) -> "AlnTable":
        """Sub-table of the given record indices (in that order)."""
        idx = np.asarray(idx, dtype=np.int64)
        off = self.cigar_off.astype(np.int64)
        n = off[idx + 1] - off[idx]
        new_off = np.concatenate([[0], np.cumsum(n)])
        # gather ops
        total = int(new_off[-1])
        src = np.repeat(off[idx] - new_off[:-1], n) + np.arange(total, dtype=np.int64)
        return AlnTable(self.ref_id[idx], self.ref_start[idx], self.mapq[idx], self.flag[idx],
                        self.nm[idx], self.qlen[idx], self.read_id[idx],
                        new_off.astype(np.uint64), self.cigar[src])

    @staticmethod
    def from_rows(rows) -> "AlnTable":
        """rows: iterable of dicts with ref_id, ref_start, mapq, flag, nm (or None), qlen, read_id, cigar (text or u32 array)."""
        rows = list(rows)
        cig = [pack_cigar(r["cigar"]) if isinstance(r["cigar"], str) else np.asarray(r["cigar"], np.uint32) for r in rows]
        off = np.concatenate([[0], np.cumsum([len(c) for c in cig])]).astype(np.uint64)
        return AlnTable(
            [r["ref_id"] for r in rows], [r["ref_start"] for r in rows], [r["mapq"] for r in rows],
            [r["flag"] for r in rows],
            [NM_MISSING if r.get("nm") is None else r["nm"] for r in rows],
            [r["qlen"] for r in rows], [r["read_id"] for r in rows], off,
            np.concatenate(cig) if cig else np.zeros(0, np.uint32))


@dataclass
class PafTable:
    """Lines of one PAF file, in file order (columns 0,1,2,3,5,7,8,9,10,11 — GCI.py:218-229)."""

    read_id: np.ndarray   # u32
    qlen: np.ndarray      # i32  col 1
    qstart: np.ndarray    # i32  col 2
    qend: np.ndarray      # i32  col 3
    ref_id: np.ndarray    # i32  col 5 interned, -1 if the target is not a selected contig
    tstart: np.ndarray    # i32  col 7
    tend: np.ndarray      # i32  col 8
    nmatch: np.ndarray    # i32  col 9
    alnlen: np.ndarray    # i32  col 10
    mapq: np.ndarray      # i32  col 11
    kind: str = "paf"

    def __post_init__(self):
        self.read_id = np.ascontiguousarray(self.read_id, dtype=np.uint32)
        for name in ("qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.int32))
            if len(getattr(self, name)) != len(self.read_id):
                raise ValueError(f"PafTable column {name} has wrong length")

    @property
    def n_records(self) -> int:
        return len(self.read_id)

    def nbytes(self) -> int:
        return sum(getattr(self, n).nbytes for n in
                   ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq"))

    @staticmethod
    def from_rows(rows) -> "PafTable":
        rows = list(rows)
        cols = ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")
        return PafTable(*[[r[c] for r in rows] for c in cols])


@dataclass
class ContigTable:
    names: list
    lengths: np.ndarray  # i32/i64 [n]

    def __post_init__(self):
        self.names = list(self.names)
        self.lengths = np.ascontiguousarray(self.lengths, dtype=np.int64)
        if len(self.names) != len(self.lengths):
            raise ValueError("ContigTable names/lengths mismatch")

    def __len__(self):
        return len(self.names)

    def name_rank(self) -> np.ndarray:
        """Rank of every contig name under Python str ordering (PAF tie-break, GCI.py:252)."""
        order = sorted(range(len(self.names)), key=lambda i: self.names[i])
        rank = np.empty(len(self.names), dtype=np.int32)
        rank[order] = np.arange(len(self.names), dtype=np.int32)
        return rank

    def as_dict(self) -> dict:
        return {n: int(l) for n, l in zip(self.names, self.lengths)}
