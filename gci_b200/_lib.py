"""ctypes binding of libgci_cuda.so (include/gci_cuda.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, constructing a `Context` raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgci_cuda.so")

INT32_MIN = -(2 ** 31)
NO_FLAGS = INT32_MIN

STAGES = {"h2d": 0, "cigar": 1, "gate": 2, "join": 3, "bucket": 4, "depth": 5, "flags": 6, "runs": 7,
          "max": 8, "mask": 9, "d2h": 10, "paf": 11, "text": 12, "score": 13, "xdispatch": 14, "xwait": 15,
          "xconsume": 16}

TRACK_HIFI, TRACK_NANO, TRACK_MERGED = 0, 1, 2


class GciError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgci_cuda error {code}: {msg}")
        self.code = code


class ReferenceWouldRaise(GciError):
    """The reference raises (ZeroDivisionError / KeyError) on this input."""


_lib = None

_p = C.c_void_p
_i32, _i64, _u32, _f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_double

_SIGNATURES = {
    "gci_version": (C.c_int, []),
    "gci_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "gci_destroy": (None, [_p]),
    "gci_last_error": (C.c_char_p, [_p]),
    "gci_set_stream": (C.c_int, [_p, _p]),
    "gci_set_timing": (C.c_int, [_p, _i32]),
    "gci_sync": (C.c_int, [_p]),
    "gci_host_alloc": (_p, [C.c_uint64]),
    "gci_host_free": (None, [_p]),
    "gci_stage_reset": (C.c_int, [_p]),
    "gci_stage_ms": (C.c_int, [_p, C.c_int, C.POINTER(_f64), C.POINTER(_i64)]),
    "gci_kernel_launches": (_i64, [_p]),
    "gci_device_bytes": (_i64, [_p]),
    "gci_graph_replays": (_i64, [_p]),
    "gci_set_contigs": (C.c_int, [_p, _i32, _p, _p]),
    "gci_set_n_runs": (C.c_int, [_p, _i64, _p, _p, _p]),
    "gci_reads_begin": (C.c_int, [_p, _u32]),
    "gci_upload_bam": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64]),
    "gci_set_name_rank": (C.c_int, [_p, _p]),
    "gci_upload_paf": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gci_upload_table": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p]),
    "gci_filter": (C.c_int, [_p, _i32, _i32, _f64, _f64, _f64, C.POINTER(_i64)]),
    "gci_fetch_cigar_stats": (C.c_int, [_p, _i32, _i64, _p, _p]),
    "gci_fetch_survivors": (C.c_int, [_p, _i64, _p, _p, _p, _p, C.POINTER(_i64)]),
    "gci_fetch_file_table": (C.c_int, [_p, _i32, _i64, _p, _p, _p, _p, _p, _p, C.POINTER(_i64)]),
    "gci_depth": (C.c_int, [_p, _i32, _i32, _i32, _i32]),
    "gci_mask_gaps": (C.c_int, [_p, _i32]),
    "gci_merge_max": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32]),
    "gci_load_depth": (C.c_int, [_p, _i32, _i32, _p, _i64]),
    "gci_fetch_depth": (C.c_int, [_p, _i32, _i32, _p, _i64]),
    "gci_fetch_depth_narrow": (C.c_int, [_p, _i32, _i32, _p, _i64, _i32, C.POINTER(_i32)]),
    "gci_depth_sums": (C.c_int, [_p, _i32, _p]),
    "gci_depth_hash": (C.c_int, [_p, _i32, _p]),
    "gci_depth_text": (C.c_int, [_p, _i32, _i32, _i64, _i64, _p, _i64, C.POINTER(_i64)]),
    "gci_depth_gzip": (C.c_int, [_p, _i32, _i32, _i64, _i64, C.c_char_p, _i32, _p, _i64, C.POINTER(_i64)]),
    "gci_depth_gzip_track": (C.c_int, [_p, _i32, C.c_char_p, _p, _p, _i64, C.POINTER(_i64), _p]),
    "gci_scan": (C.c_int, [_p, _i32, _i32, _i32, _i32, C.POINTER(_i64)]),
    "gci_scan_windows": (C.c_int, [_p, _i32, _i32, _i32, _i64, _p, _p, _p, C.POINTER(_i64)]),
    "gci_fetch_intervals": (C.c_int, [_p, _i32, _i64, _p, _p, _p, C.POINTER(_i64)]),
    "gci_load_intervals": (C.c_int, [_p, _i32, _i64, _p, _p, _p, _p]),
    "gci_score_terms": (C.c_int, [_p, _i32, _f64, _i32, _p, _p, _i64, _p, _p]),
    "gci_pipeline": (C.c_int, [_p, _i32, _i32, _i32, _f64, _f64, _f64, _i32, _i32, _i32, _f64, C.POINTER(_i64),
                               C.POINTER(_i64), _p, _p, _p]),
    "gci_pipeline_row": (C.c_int, [_p, _i32, _i32, _i32, _f64, _f64, _f64, _i32, _i32, _i32, _f64, C.POINTER(_i64),
                                   C.POINTER(_i64), _p, _p, _p, _i64, _i64, _p]),
    "gci_comm_unique_id": (C.c_int, [_p]),
    "gci_comm_init": (C.c_int, [_p, _p, _i32, _i32]),
    "gci_comm_p2p_alloc": (C.c_int, [_p, _i64, _p]),
    "gci_comm_p2p_open": (C.c_int, [_p, _p]),
    "gci_comm_p2p_disable": (C.c_int, [_p]),
    "gci_genome_row": (C.c_int, [_p, _i32, _f64, _i32, _i64, _i64, _p, _p, _p, _p]),
    "gci_score_terms_sums": (C.c_int, [_p, _i32, _f64, _i32, _p, _p, _i64, _p, _p, _p]),
    "gci_sliding_window": (C.c_int, [_p, _i32, _i32, _i64, _i64, _i64, _i64, _p, _p, _p, _p, C.POINTER(_i64)]),
    "gci_shard_home": (C.c_int, [_u32, _i32, C.POINTER(_i32), C.POINTER(_u32)]),
    "gci_shard_config": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "gci_shard_alloc": (C.c_int, [_p, _u32, _i32, _p]),
    "gci_shard_open": (C.c_int, [_p, _p]),
    "gci_shard_close": (C.c_int, [_p]),
    "gci_shard_area": (_p, [_p]),
    "gci_shard_attach": (C.c_int, [_p, _p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load_library():
    """Load libgci_cuda.so (loudly failing if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` (gci_b200/csrc/build.sh).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_p)


def _arr(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class PinnedPool:
    """Pinned host arrays (cudaHostAlloc) so that H2D/D2H copies run at PCIe speed."""

    def __init__(self):
        self._lib = load_library()
        self._ptrs = []

    def empty(self, n, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(16, int(n) * dtype.itemsize)
        p = self._lib.gci_host_alloc(nbytes)
        if not p:
            raise MemoryError("gci_host_alloc failed")
        self._ptrs.append(p)
        buf = (C.c_char * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    def copy(self, a):
        out = self.empty(a.size, a.dtype)
        out[...] = a.reshape(-1)
        return out

    def close(self):
        for p in self._ptrs:
            self._lib.gci_host_free(p)
        self._ptrs = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One GPU context: contig table, one read set at a time, up to three depth tracks."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        h = _p()
        rc = self._lib.gci_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise GciError(rc, f"gci_create(device={device}) failed — no usable CUDA device; "
                               f"libgci_cuda has no CPU fallback")
        self._h = h
        self.device = device
        self.n_contigs = 0
        self.lengths = None
        self.selected = None

    # ---- plumbing ----
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gci_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.gci_last_error(self._h).decode(errors="replace")
            if rc == -4:
                raise ReferenceWouldRaise(rc, msg)
            raise GciError(rc, msg)

    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.gci_set_stream(self._h, _p(cuda_stream_handle)))

    def set_timing(self, on):
        """Stage timers on (default): every stage is bracketed by a CUDA event pair and `pipeline` launches its
        kernels one by one.  Off: `pipeline` / `pipeline_row` replay a CUDA graph of the step once it has been
        seen twice with the same arguments and read-set shape (stage_report() then stays empty)."""
        self._check(self._lib.gci_set_timing(self._h, 1 if on else 0))

    def sync(self):
        self._check(self._lib.gci_sync(self._h))

    def stage_reset(self):
        self._check(self._lib.gci_stage_reset(self._h))

    def stage_ms(self, stage):
        ms, k = _f64(), _i64()
        self._check(self._lib.gci_stage_ms(self._h, STAGES[stage] if isinstance(stage, str) else stage,
                                           C.byref(ms), C.byref(k)))
        return ms.value, k.value

    def stage_report(self):
        return {name: self.stage_ms(name) for name in STAGES}

    @property
    def kernel_launches(self):
        return int(self._lib.gci_kernel_launches(self._h))

    @property
    def graph_replays(self):
        return int(self._lib.gci_graph_replays(self._h))

    @property
    def device_bytes(self):
        return int(self._lib.gci_device_bytes(self._h))

    # ---- contigs ----
    def set_contigs(self, lengths, selected=None):
        lengths = _arr(lengths, np.int64)
        sel = None if selected is None else _arr(np.asarray(selected, dtype=bool), np.uint8)
        self._check(self._lib.gci_set_contigs(self._h, len(lengths), _ptr(lengths), _ptr(sel)))
        self.n_contigs = len(lengths)
        self.lengths = lengths.copy()
        self.selected = np.ones(len(lengths), bool) if sel is None else sel.astype(bool)

    def set_name_rank(self, rank):
        rank = _arr(rank, np.int32)
        assert len(rank) == self.n_contigs
        self._check(self._lib.gci_set_name_rank(self._h, _ptr(rank)))

    def set_n_runs(self, contig, start, end):
        contig, start, end = _arr(contig, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        self._check(self._lib.gci_set_n_runs(self._h, len(contig), _ptr(contig), _ptr(start), _ptr(end)))

    # ---- filter ----
    def reads_begin(self, n_reads):
        self._check(self._lib.gci_reads_begin(self._h, int(n_reads)))

    def upload_bam(self, t):
        """t: AlnTable (or anything with the same column attributes)."""
        cols = [_arr(t.ref_id, np.int32), _arr(t.ref_start, np.int32), _arr(t.mapq, np.uint8),
                _arr(t.flag, np.uint16), _arr(t.nm, np.int32), _arr(t.qlen, np.int32),
                _arr(t.read_id, np.uint32), _arr(t.cigar_off, np.uint64), _arr(t.cigar, np.uint32)]
        n = len(cols[0])
        self._check(self._lib.gci_upload_bam(self._h, n, *[_ptr(c) for c in cols], len(cols[8])))

    def upload_paf(self, t):
        cols = [_arr(t.read_id, np.uint32)] + [_arr(getattr(t, c), np.int32) for c in
                                               ("qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch",
                                                "alnlen", "mapq")]
        self._check(self._lib.gci_upload_paf(self._h, len(cols[0]), *[_ptr(c) for c in cols]))

    def upload_table(self, read_id, ref_id, start, end, qlen, highq=None):
        cols = [_arr(read_id, np.uint32), _arr(ref_id, np.int32), _arr(start, np.int32), _arr(end, np.int32),
                _arr(qlen, np.int32)]
        hq = None if highq is None else _arr(highq, np.uint8)
        self._check(self._lib.gci_upload_table(self._h, len(cols[0]), *[_ptr(c) for c in cols], _ptr(hq)))

    def filter(self, map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1, ovlp_percent=0.9):
        n = _i64()
        self._check(self._lib.gci_filter(self._h, int(map_qual), int(mq_cutoff), float(iden_percent),
                                         float(clip_percent), float(ovlp_percent), C.byref(n)))
        return n.value

    def fetch_cigar_stats(self, bam_idx, n_records):
        """-> (uint32[n][5] bases in (M+=+X, I, D, N, S), int32[n] reference_end) of one BAM upload"""
        st = np.zeros((int(n_records), 5), np.uint32)
        end = np.zeros(int(n_records), np.int32)
        self._check(self._lib.gci_fetch_cigar_stats(self._h, bam_idx, int(n_records), _ptr(st), _ptr(end)))
        return st, end

    def fetch_survivors(self):
        n = _i64()
        self._check(self._lib.gci_fetch_survivors(self._h, 0, None, None, None, None, C.byref(n)))
        k = n.value
        r, c, s, e = (np.empty(k, np.uint32), np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.int32))
        if k:
            self._check(self._lib.gci_fetch_survivors(self._h, k, _ptr(r), _ptr(c), _ptr(s), _ptr(e), C.byref(n)))
        return r, c, s, e

    def fetch_file_table(self, file_idx):
        n = _i64()
        self._check(self._lib.gci_fetch_file_table(self._h, file_idx, 0, None, None, None, None, None, None,
                                                   C.byref(n)))
        k = n.value
        r = np.empty(k, np.uint32)
        c, s, e, q = (np.empty(k, np.int32) for _ in range(4))
        h = np.empty(k, np.uint8)
        if k:
            self._check(self._lib.gci_fetch_file_table(self._h, file_idx, k, _ptr(r), _ptr(c), _ptr(s), _ptr(e),
                                                       _ptr(q), _ptr(h), C.byref(n)))
        return r, c, s, e, q, h

    # ---- depth ----
    def depth(self, track, flank_len=15, lo=NO_FLAGS, hi=NO_FLAGS):
        self._check(self._lib.gci_depth(self._h, track, int(flank_len), int(lo), int(hi)))

    def mask_gaps(self, track):
        self._check(self._lib.gci_mask_gaps(self._h, track))

    def merge_max(self, a, b, out, lo=NO_FLAGS, hi=NO_FLAGS):
        self._check(self._lib.gci_merge_max(self._h, a, b, out, int(lo), int(hi)))

    def load_depth(self, track, contig, values):
        values = _arr(values, np.int32)
        self._check(self._lib.gci_load_depth(self._h, track, contig, _ptr(values), len(values)))

    def fetch_depth(self, track, contig, out=None):
        n = int(self.lengths[contig])
        if out is None:
            out = np.empty(n, np.int32)
        assert out.dtype == np.int32 and out.size == n and out.flags.c_contiguous
        self._check(self._lib.gci_fetch_depth(self._h, track, contig, _ptr(out), n))
        return out

    def fetch_depth_narrow(self, track, contig, out8=None, out16=None):
        """Depth of one contig in the narrowest exact dtype (uint8 -> uint16 -> int32): the conversion runs on
        the GPU, so 4x / 2x fewer bytes cross PCIe.  out8 / out16: optional preallocated (pinned) buffers."""
        n = int(self.lengths[contig])
        ovf = _i32()
        for width, dt, buf in ((1, np.uint8, out8), (2, np.uint16, out16)):
            out = buf if buf is not None else np.empty(n, dt)
            assert out.dtype == dt and out.size == n
            self._check(self._lib.gci_fetch_depth_narrow(self._h, track, contig, _ptr(out), n, width, C.byref(ovf)))
            if ovf.value == 0:
                return out
        return self.fetch_depth(track, contig)

    def depth_sums(self, track):
        out = np.zeros(self.n_contigs, np.int64)
        self._check(self._lib.gci_depth_sums(self._h, track, _ptr(out)))
        return out

    def depth_hash(self, track):
        """per contig: sum (depth + 1) * splitmix64(position) mod 2^64 (oracle: c_oracle.depth_hash)"""
        out = np.zeros(self.n_contigs, np.uint64)
        self._check(self._lib.gci_depth_hash(self._h, track, _ptr(out)))
        return out

    def depth_text(self, track, contig, first=0, count=None):
        """ASCII `"%d\\n"` lines of depth[first:first+count], formatted on the GPU."""
        if count is None:
            count = int(self.lengths[contig]) - first
        n = _i64()
        self._check(self._lib.gci_depth_text(self._h, track, contig, first, count, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        if n.value:
            self._check(self._lib.gci_depth_text(self._h, track, contig, first, count, _ptr(buf), n.value,
                                                 C.byref(n)))
        return buf

    def depth_gzip(self, track, contig, first=0, count=None, header=b""):
        """gzip members (compressed on the GPU) of `header` + the "%d\n" lines of depth[first:first+count]."""
        if count is None:
            count = int(self.lengths[contig]) - first
        n = _i64()
        self._check(self._lib.gci_depth_gzip(self._h, track, contig, first, count, header, len(header), None, 0,
                                             C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._check(self._lib.gci_depth_gzip(self._h, track, contig, first, count, header, len(header), _ptr(buf),
                                             n.value, C.byref(n)))
        return buf

    def depth_gzip_track(self, track, headers, out=None):
        """`.depth.gz` bytes of every selected contig of the track in one pass: headers[c] (bytes, e.g. b">chr1\n")
        then the contig's depth lines, as gzip members in contig order.  out: optional (pinned) uint8 buffer.
        -> (uint8 view of the bytes, byte offset of every contig [n_contigs + 1])"""
        assert len(headers) == self.n_contigs
        blob = b"".join(headers)
        off = np.zeros(self.n_contigs + 1, np.int64)
        off[1:] = np.cumsum([len(h) for h in headers])
        cb = np.zeros(self.n_contigs + 1, np.int64)
        n = _i64()
        if out is None:
            self._check(self._lib.gci_depth_gzip_track(self._h, track, blob, _ptr(off), None, 0, C.byref(n), None))
            out = np.empty(max(1, n.value), np.uint8)
        rc = self._lib.gci_depth_gzip_track(self._h, track, blob, _ptr(off), _ptr(out), out.size, C.byref(n), _ptr(cb))
        self._check(rc)
        return out[:n.value], cb

    # ---- scan / score ----
    def scan(self, track, lo=-1, hi=0, flank_len=15):
        n = _i64()
        self._check(self._lib.gci_scan(self._h, track, int(lo), int(hi), int(flank_len), C.byref(n)))
        return n.value

    def scan_windows(self, track, contig, start, end, lo=-1, hi=0):
        contig, start, end = _arr(contig, np.int32), _arr(start, np.int64), _arr(end, np.int64)
        n = _i64()
        self._check(self._lib.gci_scan_windows(self._h, track, int(lo), int(hi), len(contig), _ptr(contig),
                                               _ptr(start), _ptr(end), C.byref(n)))
        return n.value

    def fetch_intervals(self, track, n_owners):
        n = _i64()
        self._check(self._lib.gci_fetch_intervals(self._h, track, 0, None, None, None, C.byref(n)))
        k = n.value
        s, e = np.empty(k, np.int32), np.empty(k, np.int32)
        off = np.zeros(n_owners + 1, np.int64)
        self._check(self._lib.gci_fetch_intervals(self._h, track, k, _ptr(s), _ptr(e), _ptr(off), C.byref(n)))
        return s, e, off

    def pipeline(self, track, n_owners, map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1,
                 ovlp_percent=0.9, flank_len=15, lo=-1, hi=0, dist_percent=0.005):
        """filter -> depth -> scan -> score terms in one call, one host synchronisation.
        -> (n_survivors, n_intervals, n50[owners+1], n_ctg[owners+1], depth_sums[owners+1])"""
        ns, ni = _i64(), _i64()
        n50, nctg, sums = (np.zeros(n_owners + 1, np.int64) for _ in range(3))
        self._check(self._lib.gci_pipeline(self._h, track, int(map_qual), int(mq_cutoff), float(iden_percent),
                                           float(clip_percent), float(ovlp_percent), int(flank_len), int(lo), int(hi),
                                           float(dist_percent), C.byref(ns), C.byref(ni), _ptr(n50), _ptr(nctg),
                                           _ptr(sums)))
        return ns.value, ni.value, n50, nctg, sums

    def pipeline_row(self, track, n_owners, sum_len, cap=2048, map_qual=30, mq_cutoff=50, iden_percent=0.9,
                     clip_percent=0.1, ovlp_percent=0.9, flank_len=15, lo=-1, hi=0, dist_percent=0.005):
        """`pipeline` + the multi-GPU genome row (one ncclAllGather) in the same single synchronisation.
        -> (n_survivors, n_intervals, n50, n_ctg, depth_sums, mean_depth, total curated contigs, all lengths)"""
        world = self.comm_world
        ns, ni = _i64(), _i64()
        n50, nctg, sums = (np.zeros(n_owners + 1, np.int64) for _ in range(3))
        rows = np.zeros(world * (4 + cap), np.int64)
        self._check(self._lib.gci_pipeline_row(self._h, track, int(map_qual), int(mq_cutoff), float(iden_percent),
                                               float(clip_percent), float(ovlp_percent), int(flank_len), int(lo),
                                               int(hi), float(dist_percent), C.byref(ns), C.byref(ni), _ptr(n50),
                                               _ptr(nctg), _ptr(sums), int(sum_len), cap, _ptr(rows)))
        o = rows.reshape(world, 4 + cap)
        head = o[:, :4].sum(axis=0)
        all_len = o[:, 4:][np.arange(cap)[None, :] < o[:, 3:4]]
        mean = float(head[0]) / float(head[1]) if head[1] else float("nan")
        return ns.value, ni.value, n50, nctg, sums, mean, int(head[2]), all_len

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        buf = np.zeros(128, np.uint8)
        if load_library().gci_comm_unique_id(_ptr(buf)) != 0:
            raise GciError(-1, "ncclGetUniqueId failed (libnccl not loadable?)")
        return buf

    def comm_init(self, unique_id, rank, world):
        uid = _arr(unique_id, np.uint8)
        assert uid.size == 128
        self._check(self._lib.gci_comm_init(self._h, _ptr(uid), int(rank), int(world)))
        self.comm_world = int(world)

    def comm_p2p_alloc(self, cap=2048):
        """-> 64-byte CUDA IPC handle of this rank's receive area for the peer-memory genome row"""
        h = np.zeros(64, np.uint8)
        self._check(self._lib.gci_comm_p2p_alloc(self._h, int(cap), _ptr(h)))
        return h

    def comm_p2p_open(self, handles):
        """handles: uint8[world, 64], row r = what rank r got from comm_p2p_alloc"""
        h = _arr(handles, np.uint8)
        assert h.size == 64 * self.comm_world
        self._check(self._lib.gci_comm_p2p_open(self._h, _ptr(h)))

    def comm_p2p_disable(self):
        self._check(self._lib.gci_comm_p2p_disable(self._h))

    def sliding_window(self, track, contig, start, end, window_size):
        """points of GCI.py:660-705 over depth[start:end]: -> (idx, num, den, kind) int64 x3 + uint8, position order"""
        n = _i64()
        self._check(self._lib.gci_sliding_window(self._h, track, contig, int(start), int(end), int(window_size), 0,
                                                 None, None, None, None, C.byref(n)))
        k = n.value
        idx, num, den = (np.empty(k, np.int64) for _ in range(3))
        kind = np.empty(k, np.uint8)
        if k:
            self._check(self._lib.gci_sliding_window(self._h, track, contig, int(start), int(end), int(window_size), k,
                                                     _ptr(idx), _ptr(num), _ptr(den), _ptr(kind), C.byref(n)))
        return idx, num, den, kind

    # ---- read sets sharded over ranks (contig owners, read homes) ----
    def shard_config(self, rank, world, contig_owner, gate_selected=None):
        owner = _arr(contig_owner, np.int32)
        assert len(owner) == self.n_contigs
        gs = None if gate_selected is None else _arr(np.asarray(gate_selected, dtype=bool), np.uint8)
        self._check(self._lib.gci_shard_config(self._h, int(rank), int(world), _ptr(owner), _ptr(gs)))
        self.shard_rank, self.shard_world = int(rank), int(world)

    def shard_alloc(self, max_reads, max_bam_files=2, ipc=True):
        """-> 64-byte CUDA IPC handle of this rank's exchange area (ipc=False: no handle, for contexts of one process)"""
        h = np.zeros(64, np.uint8)
        self._check(self._lib.gci_shard_alloc(self._h, int(max_reads), int(max_bam_files), _ptr(h) if ipc else None))
        return h

    def shard_open(self, handles):
        h = _arr(handles, np.uint8)
        assert h.size == 64 * self.shard_world
        self._check(self._lib.gci_shard_open(self._h, _ptr(h)))

    def shard_close(self):
        self._check(self._lib.gci_shard_close(self._h))

    def shard_area(self):
        return int(self._lib.gci_shard_area(self._h) or 0)

    def shard_attach(self, areas):
        arr = (C.c_void_p * len(areas))(*[C.c_void_p(a) for a in areas])
        self._check(self._lib.gci_shard_attach(self._h, arr))

    def genome_row(self, track, n_owners, sum_len, dist_percent=0.005, flank_len=15, cap=2048):
        """Score terms of this rank's contigs + one NCCL all-gather of every rank's genome-row terms.
        -> (n50, n_ctg, depth_sums, mean_depth, total curated contigs, all curated lengths)"""
        world = self.comm_world
        n50, nctg, sums = (np.zeros(n_owners + 1, np.int64) for _ in range(3))
        rows = np.zeros(world * (4 + cap), np.int64)
        self._check(self._lib.gci_genome_row(self._h, track, float(dist_percent), int(flank_len), int(sum_len), cap,
                                             _ptr(n50), _ptr(nctg), _ptr(sums), _ptr(rows)))
        o = rows.reshape(world, 4 + cap)
        head = o[:, :4].sum(axis=0)
        all_len = o[:, 4:][np.arange(cap)[None, :] < o[:, 3:4]]
        mean = float(head[0]) / float(head[1]) if head[1] else float("nan")
        return n50, nctg, sums, mean, int(head[2]), all_len

    def load_intervals(self, track, contig, owner_off, start, end):
        contig, owner_off = _arr(contig, np.int32), _arr(owner_off, np.int64)
        start, end = _arr(start, np.int32), _arr(end, np.int32)
        self._check(self._lib.gci_load_intervals(self._h, track, len(contig), _ptr(contig), _ptr(owner_off),
                                                 _ptr(start), _ptr(end)))

    def score_terms(self, track, n_owners, n_intervals, dist_percent=0.005, flank_len=15, with_sums=False):
        """-> (n50[owners+1], n_ctg[owners+1], lengths, lengths_off[owners+1]); the last n50 / n_ctg entry is
        over all owners together (the Genome / All_regions row).  with_sums: also the per-contig depth sums
        (+ total) from the same device->host copy."""
        n50 = np.zeros(n_owners + 1, np.int64)
        nctg = np.zeros(n_owners + 1, np.int64)
        off = np.zeros(n_owners + 1, np.int64)
        lengths = np.zeros(int(n_intervals) + n_owners + 1, np.int64)
        if with_sums:
            sums = np.zeros(n_owners + 1, np.int64)
            self._check(self._lib.gci_score_terms_sums(self._h, track, float(dist_percent), int(flank_len), _ptr(n50),
                                                       _ptr(nctg), len(lengths), _ptr(lengths), _ptr(off), _ptr(sums)))
            return n50, nctg, lengths[:int(off[-1])], off, sums
        self._check(self._lib.gci_score_terms(self._h, track, float(dist_percent), int(flank_len), _ptr(n50),
                                              _ptr(nctg), len(lengths), _ptr(lengths), _ptr(off)))
        return n50, nctg, lengths[:int(off[-1])], off
