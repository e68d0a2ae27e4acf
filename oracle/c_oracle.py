"""ctypes wrapper of oracle/gci_oracle.c — TEST INFRASTRUCTURE ONLY (see oracle/gci_oracle.py header).

`hot_path()` runs BAM gates -> dedup -> join -> depth -> collapse on the CPU with pthreads; it is the
timed "port" baseline of bench.py and a second checker in tests."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None
_p = C.c_void_p


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_collapse.restype = C.c_int64
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_depth_hash.restype = C.c_uint64
        _lib.orc_write_depth.restype = C.c_int64
    return _lib


def _ptr(a):
    return a.ctypes.data_as(_p)


def max_threads():
    return int(lib().orc_max_threads())


class RaisesLikeReference(Exception):
    pass


def bam_leg(tab, selected, n_reads, map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, threads=1):
    """-> per-read table (contig or -1, start, end, qlen) + highq marks, like gci_oracle.bam_leg"""
    L = lib()
    n = tab.n_records
    passed = np.zeros(max(1, n), np.int8)
    ref_end = np.zeros(max(1, n), np.int32)
    sel = np.ascontiguousarray(selected, dtype=np.uint8)
    bad = L.orc_gate(C.c_int64(n), _ptr(tab.ref_id), _ptr(tab.ref_start), _ptr(tab.mapq), _ptr(tab.flag),
                     _ptr(tab.nm), _ptr(tab.cigar_off), _ptr(tab.cigar), _ptr(sel), C.c_int32(len(sel)),
                     C.c_int32(map_qual), C.c_double(iden_percent), C.c_double(clip_percent), _ptr(passed),
                     _ptr(ref_end), C.c_int(threads))
    if bad:
        raise RaisesLikeReference(f"the reference raises on this input (mask {bad})")
    win = np.empty(max(1, n_reads), np.int64)
    highq = np.zeros(max(1, n_reads), np.uint8)
    L.orc_dedup(C.c_int64(n), _ptr(tab.ref_id), _ptr(tab.mapq), _ptr(tab.read_id), _ptr(passed),
                C.c_uint32(n_reads), C.c_int32(mq_cutoff), _ptr(win), _ptr(highq))
    win = win[:n_reads]
    idx = (win & 0xffffffff).astype(np.int64)
    present = win >= 0
    idx[~present] = 0
    if n == 0:
        z = np.zeros(n_reads, np.int32)
        return np.full(n_reads, -1, np.int32), z, z.copy(), z.copy(), highq[:n_reads]
    c = np.where(present, tab.ref_id[idx], -1).astype(np.int32)
    return c, tab.ref_start[idx].astype(np.int32), ref_end[idx].astype(np.int32), tab.qlen[idx].astype(np.int32), \
        highq[:n_reads]


def join(tables, highq, ovlp_percent=0.9, threads=1):
    L = lib()
    f = len(tables)
    n_reads = len(highq)
    arr = lambda k: (C.c_void_p * f)(*[_ptr(np.ascontiguousarray(t[k], np.int32)) for t in tables])
    keep = [[np.ascontiguousarray(t[k], np.int32) for t in tables] for k in range(4)]
    ptrs = [(C.c_void_p * f)(*[_ptr(a) for a in keep[k]]) for k in range(4)]
    oc, os_, oe = (np.empty(max(1, n_reads), np.int32) for _ in range(3))
    hq = np.ascontiguousarray(highq, np.uint8)
    bad = L.orc_join(C.c_int32(f), C.c_uint32(n_reads), ptrs[0], ptrs[1], ptrs[2], ptrs[3], _ptr(hq),
                     C.c_double(ovlp_percent), _ptr(oc), _ptr(os_), _ptr(oe), C.c_int(threads))
    if bad:
        raise RaisesLikeReference("ZeroDivisionError at GCI.py:292")
    return oc[:n_reads], os_[:n_reads], oe[:n_reads]


def depth(sc, ss, se, lengths, flank_len=15, threads=1):
    L = lib()
    lengths = np.ascontiguousarray(lengths, np.int64)
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    d = np.zeros(int(off[-1]), np.int64)
    L.orc_depth(C.c_uint32(len(sc)), _ptr(np.ascontiguousarray(sc, np.int32)), _ptr(np.ascontiguousarray(ss, np.int32)),
                _ptr(np.ascontiguousarray(se, np.int32)), C.c_int32(flank_len), C.c_int32(len(lengths)),
                _ptr(lengths), _ptr(off), _ptr(d), C.c_int(threads))
    return [d[off[i]:off[i + 1]] for i in range(len(lengths))]


def collapse(depth_arr, leftmost=-1, rightmost=0, flank_len=15, start_pos=0):
    L = lib()
    d = np.ascontiguousarray(depth_arr, np.int64)
    cap = 1024
    while True:
        s, e = np.empty(cap, np.int64), np.empty(cap, np.int64)
        n = L.orc_collapse(_ptr(d), C.c_int64(len(d)), C.c_int64(leftmost), C.c_int64(rightmost),
                           C.c_int64(flank_len), C.c_int64(start_pos), _ptr(s), _ptr(e), C.c_int64(cap))
        if n <= cap:
            return list(zip(s[:n].tolist(), e[:n].tolist()))
        cap = int(n)


def paf_leg(pafs, selected, name_rank, n_reads, map_qual=30, mq_cutoff=50, iden_percent=0.9, threads=1):
    """GCI.py:211-254 over all PAF files of a read type -> ([per file (contig or -1, start, end, qlen)], highq)"""
    L = lib()
    f = len(pafs)
    off = np.concatenate([[0], np.cumsum([t.n_records for t in pafs])]).astype(np.int64)
    cat = lambda k, dt: np.ascontiguousarray(np.concatenate([getattr(t, k) for t in pafs]) if f else np.zeros(0, dt), dt)
    cols = [cat("read_id", np.uint32)] + [cat(k, np.int32) for k in
                                          ("qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")]
    sel = np.ascontiguousarray(selected, dtype=np.uint8)
    rank = np.ascontiguousarray(name_rank, dtype=np.int32)
    nr = max(1, n_reads)
    oc, os_, oe, oq = (np.zeros(max(1, f) * nr, np.int32) for _ in range(4))
    highq = np.zeros(nr, np.uint8)
    bad = L.orc_paf_leg(C.c_int32(f), _ptr(off), *[_ptr(c) for c in cols], _ptr(sel), C.c_int32(len(sel)), _ptr(rank),
                        C.c_uint32(n_reads), C.c_int32(map_qual), C.c_int32(mq_cutoff), C.c_double(iden_percent),
                        _ptr(oc), _ptr(os_), _ptr(oe), _ptr(oq), _ptr(highq), C.c_int(threads))
    if bad:
        raise RaisesLikeReference(f"the reference raises on this input (mask {bad})")
    tabs = [tuple(x[i * nr:i * nr + n_reads] for x in (oc, os_, oe, oq)) for i in range(f)]
    return tabs, highq[:n_reads]


def depth_hash(depth_arr, threads=1):
    """order-independent 64-bit checksum of one depth array; gci_depth_hash computes the same on the GPU"""
    d = np.ascontiguousarray(depth_arr, np.int64)
    return int(lib().orc_depth_hash(_ptr(d), C.c_int64(len(d)), C.c_int(threads)))


def write_depth(depth_arr, name, threads=1, level=9, workers=None, fetch=False):
    """write_depth of one contig (GCI.py:99-143): `threads` gzip members (level 9 like gzip.open) of ">name" + one
    decimal per line, formatted and deflated on `workers` host threads.  -> number of bytes, or the bytes (fetch)."""
    d = np.ascontiguousarray(depth_arr, dtype=np.int64)
    workers = int(workers or threads)
    nm = str(name).encode()
    n = int(lib().orc_write_depth(_ptr(d), C.c_int64(len(d)), C.c_char_p(nm), int(threads), int(level), workers, None,
                                  C.c_int64(0)))
    if n < 0:
        raise RuntimeError("orc_write_depth failed")
    if not fetch:
        return n
    out = np.zeros(max(1, n), np.uint8)
    m = int(lib().orc_write_depth(_ptr(d), C.c_int64(len(d)), C.c_char_p(nm), int(threads), int(level), workers,
                                  _ptr(out), C.c_int64(n)))
    assert m == n
    return out[:n].tobytes()


def name_rank(names):
    order = sorted(range(len(names)), key=lambda i: names[i])
    rank = np.empty(len(names), np.int32)
    rank[order] = np.arange(len(names), dtype=np.int32)
    return rank


def survivors(bams, lengths, n_reads, selected=None, map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1,
              ovlp_percent=0.9, threads=1, pafs=(), names=None):
    """PAF election + BAM gates -> dedup -> join (PAFs first, GCI.py:272): per read (contig or -1, start, end)"""
    if selected is None:
        selected = np.ones(len(lengths), bool)
    tables, hq = [], np.zeros(n_reads, np.uint8)
    if len(pafs):
        if names is None:
            names = [f"chr{i + 1}" for i in range(len(lengths))]
        ptabs, phq = paf_leg(list(pafs), selected, name_rank(names), n_reads, map_qual, mq_cutoff, iden_percent, threads)
        tables += ptabs
        hq |= phq
    for t in bams:
        c, s, e, q, h = bam_leg(t, selected, n_reads, map_qual, mq_cutoff, iden_percent, clip_percent, threads)
        tables.append((c, s, e, q))
        hq |= h
    return join(tables, hq, ovlp_percent, threads)


def hot_path(bams, lengths, n_reads, selected=None, map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1,
             ovlp_percent=0.9, flank_len=15, threshold=0, threads=1, pafs=(), names=None):
    """Whole hot path on the CPU: filter -> depth -> collapse.  Returns (depths, beds, n_survivors)."""
    if selected is None:
        selected = np.ones(len(lengths), bool)
    sc, ss, se = survivors(bams, lengths, n_reads, selected, map_qual, mq_cutoff, iden_percent, clip_percent,
                           ovlp_percent, threads, pafs, names)
    lens = [int(l) if selected[i] else 0 for i, l in enumerate(lengths)]
    depths = depth(sc, ss, se, lens, flank_len, threads)
    beds = [collapse(d, -1, threshold, flank_len, 0) for d in depths]
    return depths, beds, int((sc >= 0).sum())


def hot_path_summary(bams, lengths, n_reads, selected=None, flank_len=15, threshold=0, threads=1, pafs=(), names=None,
                     with_hash=True, with_write=False, **gates):
    """The same path with one contig's depth array alive at a time (a 3.1 Gbp genome is 25 GB of int64 otherwise):
    -> (beds, depth checksums (orc_depth_hash) or None, depth sums, n_survivors, seconds spent on the checksums).
    with_write: every contig's depth also goes through write_depth (GCI.py:99-143, `threads` members per contig,
    level 9, in memory); the byte count is left in hot_path_summary.gz_bytes."""
    import time
    if selected is None:
        selected = np.ones(len(lengths), bool)
    sc, ss, se = survivors(bams, lengths, n_reads, selected, threads=threads, pafs=pafs, names=names, **gates)
    order = np.argsort(sc, kind="stable")
    bounds = np.searchsorted(sc[order], np.arange(len(lengths) + 1))
    beds, hashes, sums, t_hash = [], [], [], 0.0
    gz_bytes = 0
    for c, L in enumerate(lengths):
        if not selected[c]:
            beds.append([]); hashes.append(0); sums.append(0)
            continue
        idx = order[bounds[c]:bounds[c + 1]]
        d = depth(np.zeros(len(idx), np.int32), ss[idx], se[idx], [int(L)], flank_len, threads)[0]
        beds.append(collapse(d, -1, threshold, flank_len, 0))
        if with_write:
            gz_bytes += write_depth(d, names[c] if names is not None else f"contig{c}", threads, 9, threads)
        t0 = time.perf_counter()
        hashes.append(depth_hash(d, threads) if with_hash else 0)
        sums.append(int(d.sum()))
        t_hash += time.perf_counter() - t0
    hot_path_summary.gz_bytes = gz_bytes
    return beds, (hashes if with_hash else None), sums, int((sc >= 0).sum()), t_hash
