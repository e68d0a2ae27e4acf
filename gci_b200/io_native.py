"""ctypes binding of libgci_io.so (include/gci_io.h): native BAM / PAF / FASTA decoders."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .records import AlnTable, PafTable

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgci_io.so")
_lib = None
_p = C.c_void_p

_SIG = {
    "gci_io_last_error": (C.c_char_p, []),
    "gci_interner_create": (_p, []),
    "gci_interner_destroy": (None, [_p]),
    "gci_interner_size": (C.c_int64, [_p]),
    "gci_bam_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_p)]),
    "gci_bam_close": (None, [_p]),
    "gci_bam_n_refs": (C.c_int32, [_p]),
    "gci_bam_ref_name": (C.c_char_p, [_p, C.c_int32]),
    "gci_bam_ref_len": (C.c_int64, [_p, C.c_int32]),
    "gci_bam_n_records": (C.c_int64, [_p]),
    "gci_bam_n_ops": (C.c_int64, [_p]),
    "gci_bam_fill": (C.c_int, [_p] * 11),
    "gci_paf_open": (C.c_int, [C.c_char_p, _p, C.c_int32, _p, C.POINTER(_p)]),
    "gci_paf_n_lines": (C.c_int64, [_p]),
    "gci_paf_fill": (C.c_int, [_p] * 11),
    "gci_paf_close": (None, [_p]),
    "gci_fasta_open": (C.c_int, [C.c_char_p, C.POINTER(_p)]),
    "gci_fasta_n_records": (C.c_int32, [_p]),
    "gci_fasta_id": (C.c_char_p, [_p, C.c_int32]),
    "gci_fasta_n_runs": (C.c_int64, [_p]),
    "gci_fasta_runs": (C.c_int, [_p, _p, _p, _p]),
    "gci_fasta_close": (None, [_p]),
    "gci_depth_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_p)]),
    "gci_depth_n_contigs": (C.c_int32, [_p]),
    "gci_depth_name": (C.c_char_p, [_p, C.c_int32]),
    "gci_depth_len": (C.c_int64, [_p, C.c_int32]),
    "gci_depth_fill": (C.c_int, [_p, C.c_int32, _p]),
    "gci_depth_close": (None, [_p]),
}


def exported_symbols():
    return sorted(_SIG)


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError(f"{LIB_PATH} is missing: run gci_b200/csrc/build.sh")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _err():
    return lib().gci_io_last_error().decode(errors="replace")


def _ptr(a):
    return a.ctypes.data_as(_p)


class Interner:
    """Read-name table shared by all files of one read type (exact, dense ids in first-seen order)."""

    def __init__(self):
        self._h = _p(lib().gci_interner_create())

    def __len__(self):
        return int(lib().gci_interner_size(self._h))

    def __del__(self):
        try:
            if self._h:
                lib().gci_interner_destroy(self._h)
                self._h = None
        except Exception:
            pass


def read_bam(path, interner: Interner, threads=1):
    """-> (names, lengths, AlnTable)"""
    L = lib()
    h = _p()
    if L.gci_bam_open(os.fsencode(path), int(threads), C.byref(h)) != 0:
        raise ValueError(f"{path}: {_err()}")
    try:
        names = [L.gci_bam_ref_name(h, i).decode() for i in range(L.gci_bam_n_refs(h))]
        lengths = [int(L.gci_bam_ref_len(h, i)) for i in range(len(names))]
        n, c = int(L.gci_bam_n_records(h)), int(L.gci_bam_n_ops(h))
        cols = dict(ref_id=np.empty(n, np.int32), ref_start=np.empty(n, np.int32), mapq=np.empty(n, np.uint8),
                    flag=np.empty(n, np.uint16), nm=np.empty(n, np.int32), qlen=np.empty(n, np.int32),
                    read_id=np.empty(n, np.uint32), cigar_off=np.empty(n + 1, np.uint64), cigar=np.empty(c, np.uint32))
        if L.gci_bam_fill(h, interner._h, *[_ptr(cols[k]) for k in ("ref_id", "ref_start", "mapq", "flag", "nm", "qlen",
                                                                     "read_id", "cigar_off", "cigar")]) != 0:
            raise ValueError(f"{path}: {_err()}")
    finally:
        L.gci_bam_close(h)
    tab = AlnTable(**cols)
    tab.contig_names, tab.contig_lengths = names, lengths
    return names, lengths, tab


def read_bam_header(path):
    # the header lives in the first BGZF blocks; the pure-Python reader stops after them
    from . import io as gio
    return gio.read_bam_header_py(path)


def read_paf(path, contig_names, interner: Interner):
    L = lib()
    h = _p()
    enc = [n.encode() for n in contig_names]
    arr = (C.c_char_p * max(1, len(enc)))(*enc)
    if L.gci_paf_open(os.fsencode(path), interner._h, len(enc), arr, C.byref(h)) != 0:
        raise ValueError(f"{path}: {_err()}")
    try:
        n = int(L.gci_paf_n_lines(h))
        rid = np.empty(n, np.uint32)
        cols = [np.empty(n, np.int32) for _ in range(9)]
        L.gci_paf_fill(h, _ptr(rid), *[_ptr(c) for c in cols])
    finally:
        L.gci_paf_close(h)
    return PafTable(rid, *cols)


def read_fasta_gaps(path):
    L = lib()
    h = _p()
    if L.gci_fasta_open(os.fsencode(path), C.byref(h)) != 0:
        raise ValueError(f"{path}: {_err()}")
    try:
        ids = [L.gci_fasta_id(h, i).decode() for i in range(L.gci_fasta_n_records(h))]
        k = int(L.gci_fasta_n_runs(h))
        rec, s, e = np.empty(k, np.int32), np.empty(k, np.int64), np.empty(k, np.int64)
        if k:
            L.gci_fasta_runs(h, _ptr(rec), _ptr(s), _ptr(e))
    finally:
        L.gci_fasta_close(h)
    gaps = {}
    for r, a, b in zip(rec.tolist(), s.tolist(), e.tolist()):
        gaps.setdefault(ids[r], []).append((a, b))
    return ids, gaps


def read_depth_gz(path, threads=0):
    """utility/GCI_score.py:11-39 — {name: int32 array} in file order, parsed by libgci_io.so"""
    L = lib()
    h = _p()
    if L.gci_depth_open(os.fsencode(path), int(threads) or (os.cpu_count() or 1), C.byref(h)) != 0:
        raise ValueError(f"{path}: {_err()}")
    try:
        out = {}
        for i in range(L.gci_depth_n_contigs(h)):
            a = np.empty(int(L.gci_depth_len(h, i)), np.int32)
            if L.gci_depth_fill(h, i, _ptr(a)) != 0:
                raise ValueError(f"{path}: {_err()}")
            out[L.gci_depth_name(h, i).decode()] = a
        return out
    finally:
        L.gci_depth_close(h)
