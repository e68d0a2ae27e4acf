"""File formats either side of the hot path (pure-Python host code; the native decoder is the
"next" row of SURVEY.md §8f): BAM/BGZF, PAF and FASTA readers into the columnar schema, a small
BAM writer for tests, and the `.depth.gz` / BED / `.gci` writers.

pysam / htslib / Biopython are not available in the image, so the readers restate the SAM/BAM
spec for exactly the fields the reference touches (GCI.py:150-166, :201-208, :218-229, :30-35).
"""
from __future__ import annotations

import gzip
import re
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .records import AlnTable, PafTable, NM_MISSING

# ------------------------------------------------------------------------------------------------
# BGZF / BAM
# ------------------------------------------------------------------------------------------------
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_decompress(path) -> bytes:
    """Concatenated payload of all BGZF blocks (a BGZF file is a multi-member gzip)."""
    with open(path, "rb") as f:
        raw = f.read()
    out = []
    pos = 0
    n = len(raw)
    while pos < n:
        if raw[pos:pos + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError(f"{path}: not a BGZF block at byte {pos}")
        xlen = struct.unpack_from("<H", raw, pos + 10)[0]
        extra = raw[pos + 12:pos + 12 + xlen]
        bsize = None
        k = 0
        while k + 4 <= len(extra):
            si1, si2, slen = extra[k], extra[k + 1], struct.unpack_from("<H", extra, k + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", extra, k + 4)[0] + 1
            k += 4 + slen
        if bsize is None:
            raise ValueError(f"{path}: BGZF block without BSIZE")
        cdata = raw[pos + 12 + xlen:pos + bsize - 8]
        out.append(zlib.decompress(cdata, -15))
        pos += bsize
    return b"".join(out)


def _bgzf_blocks(payload: bytes, level=6):
    for i in range(0, max(len(payload), 1), 0xff00):
        chunk = payload[i:i + 0xff00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = c.compress(chunk) + c.flush()
        bsize = len(cdata) + 25
        yield (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + cdata +
               struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    yield _BGZF_EOF


def read_bam_header_py(path):
    """(names, lengths) of the BAM header (GCI.py:201-207, :963-965): inflates only the leading BGZF blocks."""
    data = b""
    with open(path, "rb") as f:
        while True:
            head = f.read(18)
            if len(head) < 18:
                break
            xlen = struct.unpack_from("<H", head, 10)[0]
            extra = head[12:] + f.read(max(0, xlen - 6))
            bsize = None
            k = 0
            while k + 4 <= xlen:
                slen = struct.unpack_from("<H", extra, k + 2)[0]
                if extra[k] == 66 and extra[k + 1] == 67 and slen == 2:
                    bsize = struct.unpack_from("<H", extra, k + 4)[0] + 1
                k += 4 + slen
            if bsize is None:
                raise ValueError(f"{path}: BGZF block without BSIZE")
            rest = f.read(bsize - 12 - xlen)
            data += zlib.decompress(rest[:-8], -15)
            try:
                names, lengths, _ = _parse_bam_header(data)
                return names, lengths
            except (struct.error, IndexError):
                continue
    names, lengths, _ = _parse_bam_header(data)
    return names, lengths


read_bam_header = read_bam_header_py


def _parse_bam_header(data):
    if data[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", data, 4)[0]
    pos = 8 + l_text
    n_ref = struct.unpack_from("<i", data, pos)[0]
    pos += 4
    names, lengths = [], []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, pos)[0]
        names.append(data[pos + 4:pos + 4 + l_name - 1].decode())
        lengths.append(struct.unpack_from("<i", data, pos + 4 + l_name)[0])
        pos += 8 + l_name
    return names, lengths, pos


_AUX_FIXED = {ord("A"): 1, ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2, ord("i"): 4, ord("I"): 4, ord("f"): 4}
_AUX_FMT = {ord("c"): "<b", ord("C"): "<B", ord("s"): "<h", ord("S"): "<H", ord("i"): "<i", ord("I"): "<I"}
_B_SIZE = {ord("c"): 1, ord("C"): 1, ord("s"): 2, ord("S"): 2, ord("i"): 4, ord("I"): 4, ord("f"): 4}


def _scan_aux(data, pos, end):
    """-> (NM value or None, CG ops array or None)"""
    nm, cg = None, None
    while pos + 3 <= end:
        tag = data[pos:pos + 2]
        typ = data[pos + 2]
        pos += 3
        if typ in _AUX_FIXED:
            if tag == b"NM" and typ in _AUX_FMT:
                nm = struct.unpack_from(_AUX_FMT[typ], data, pos)[0]
            pos += _AUX_FIXED[typ]
        elif typ in (ord("Z"), ord("H")):
            pos = data.index(b"\x00", pos) + 1
        elif typ == ord("B"):
            sub = data[pos]
            cnt = struct.unpack_from("<i", data, pos + 1)[0]
            if tag == b"CG" and sub == ord("I"):
                cg = np.frombuffer(data, dtype="<u4", count=cnt, offset=pos + 5)
            pos += 5 + cnt * _B_SIZE[sub]
        else:
            raise ValueError(f"unknown aux type {chr(typ)}")
    return nm, cg


def read_bam_py(path, intern=None):
    """Decode a BAM file into (names, lengths, AlnTable).  `intern` maps read names to dense ids and
    is shared by all files of one read type (it is updated in place)."""
    if intern is None:
        intern = {}
    data = bgzf_decompress(path)
    names, lengths, pos = _parse_bam_header(data)
    n = len(data)
    ref_id, start, mapq, flag, nm, qlen, rid, n_ops = [], [], [], [], [], [], [], []
    cig = []
    unpack = struct.Struct("<iiiBBHHHi").unpack_from
    while pos + 4 <= n:
        block_size, r, p, l_name, mq, _bin, n_cig, fl, l_seq = unpack(data, pos)
        rec_end = pos + 4 + block_size
        q0 = pos + 36
        qname = data[q0:q0 + l_name - 1]
        c0 = q0 + l_name
        ops = np.frombuffer(data, dtype="<u4", count=n_cig, offset=c0)
        aux0 = c0 + 4 * n_cig + (l_seq + 1) // 2 + l_seq
        nmv, cg = _scan_aux(data, aux0, rec_end)
        if cg is not None and n_cig == 2 and (int(ops[0]) & 15) == 4 and (int(ops[0]) >> 4) == l_seq and \
                (int(ops[1]) & 15) == 3:
            ops = cg                                    # long CIGAR stored in the CG:B,I tag (SAM spec §4.2.2)
        ref_id.append(r)
        start.append(p)
        mapq.append(mq)
        flag.append(fl)
        nm.append(NM_MISSING if nmv is None else nmv)
        qlen.append(l_seq)
        rid.append(intern.setdefault(qname, len(intern)))
        n_ops.append(len(ops))
        cig.append(ops)
        pos = rec_end
    off = np.concatenate([[0], np.cumsum(n_ops)]).astype(np.uint64) if n_ops else np.zeros(1, np.uint64)
    cigar = np.concatenate(cig).astype(np.uint32) if cig else np.zeros(0, np.uint32)
    tab = AlnTable(ref_id, start, mapq, flag, nm, qlen, rid, off, cigar)
    return names, lengths, tab


class NameTable:
    """Read-name interning shared by the files of one read type: native table when libgci_io.so is built,
    a dict otherwise (the pure-Python decoders are kept as a cross-check of the native ones)."""

    def __init__(self, native=None):
        from . import io_native
        self.native = io_native.available() if native is None else native
        self.table = io_native.Interner() if self.native else {}

    def __len__(self):
        return len(self.table)


def read_bam(path, names: "NameTable" = None, threads=1):
    """Decode a BAM file -> (contig names, lengths, AlnTable)."""
    names = NameTable() if names is None else names
    if names.native:
        from . import io_native
        return io_native.read_bam(path, names.table, threads)
    n, l, tab = read_bam_py(path, names.table)
    tab.contig_names, tab.contig_lengths = n, l
    return n, l, tab


def read_paf(path, contig_names, names: "NameTable" = None) -> PafTable:
    names = NameTable() if names is None else names
    if names.native:
        from . import io_native
        return io_native.read_paf(path, list(contig_names), names.table)
    return read_paf_py(path, {c: i for i, c in enumerate(contig_names)}, names.table)


def read_fasta_gaps(path):
    from . import io_native
    if io_native.available():
        return io_native.read_fasta_gaps(path)
    return read_fasta_gaps_py(path)


def write_bam(path, names, lengths, tab: AlnTable, read_names=None, level=1, long_cigar_as_cg=True):
    """Minimal BAM writer (tests and synthetic twins): SEQ/QUAL are written as '*' unless needed for
    l_seq, in which case zero bytes are emitted."""
    out = [b"BAM\x01"]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(names, lengths))
    tb = text.encode()
    out.append(struct.pack("<i", len(tb)) + tb + struct.pack("<i", len(names)))
    for n, l in zip(names, lengths):
        nb = n.encode() + b"\x00"
        out.append(struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(l)))
    off = tab.cigar_off.astype(np.int64)
    for i in range(tab.n_records):
        name = (read_names[int(tab.read_id[i])] if read_names is not None else f"read{int(tab.read_id[i])}")
        nb = name.encode() + b"\x00"
        ops = tab.cigar[off[i]:off[i + 1]].astype("<u4")
        l_seq = int(tab.qlen[i])
        aux = b""
        nmv = int(tab.nm[i])
        if nmv != int(NM_MISSING):
            aux += b"NMi" + struct.pack("<i", nmv)
        if len(ops) > 65535 and long_cigar_as_cg:
            rlen = int(sum(int(o) >> 4 for o in ops if (int(o) & 15) in (0, 2, 3, 7, 8)))
            aux += b"CGBI" + struct.pack("<i", len(ops)) + ops.tobytes()
            ops = np.array([(l_seq << 4) | 4, (rlen << 4) | 3], dtype="<u4")
        body = struct.pack("<iiBBHHHiiii", int(tab.ref_id[i]), int(tab.ref_start[i]), len(nb), int(tab.mapq[i]),
                           4680, len(ops), int(tab.flag[i]), l_seq, -1, -1, 0)
        body += nb + ops.tobytes() + b"\x00" * ((l_seq + 1) // 2) + b"\xff" * l_seq + aux
        out.append(struct.pack("<i", len(body)) + body)
    payload = b"".join(out)
    with open(path, "wb") as f:
        for blk in _bgzf_blocks(payload, level):
            f.write(blk)


# ------------------------------------------------------------------------------------------------
# PAF / FASTA / BED
# ------------------------------------------------------------------------------------------------

def read_paf_py(path, contig_index, intern=None) -> PafTable:
    """Columns 0,1,2,3,5,7,8,9,10,11 of a PAF file (GCI.py:218-229); `line.strip().split('\\t')`."""
    if intern is None:
        intern = {}
    cols = [[] for _ in range(10)]
    with open(path, "r") as f:
        for line in f:
            p = line.strip().split("\t")
            cols[0].append(intern.setdefault(p[0].encode(), len(intern)))
            cols[1].append(int(p[1]))
            cols[2].append(int(p[2]))
            cols[3].append(int(p[3]))
            cols[4].append(contig_index.get(p[5], -1))
            cols[5].append(int(p[7]))
            cols[6].append(int(p[8]))
            cols[7].append(int(p[9]))
            cols[8].append(int(p[10]))
            cols[9].append(int(p[11]))
    return PafTable(*cols)


def write_paf(path, tab: PafTable, names, lengths, read_names=None):
    with open(path, "w") as f:
        for i in range(tab.n_records):
            t = int(tab.ref_id[i])
            q = read_names[int(tab.read_id[i])] if read_names is not None else f"read{int(tab.read_id[i])}"
            f.write("\t".join(map(str, [q, int(tab.qlen[i]), int(tab.qstart[i]), int(tab.qend[i]), "+", names[t],
                                        int(lengths[t]), int(tab.tstart[i]), int(tab.tend[i]), int(tab.nmatch[i]),
                                        int(tab.alnlen[i]), int(tab.mapq[i])])) + "\n")


_N_RUN = re.compile(rb"[Nn]+")


def read_fasta_gaps_py(path):
    """(record ids in file order, {id: [(start, end), ...]} of N/n runs) — GCI.py:28-35, :939-941.
    The id is the first whitespace-delimited token of the header like Biopython's `record.id`."""
    ids, gaps = [], {}
    opener = gzip.open if str(path).endswith(".gz") else open
    name, chunks = None, []

    def flush():
        if name is None:
            return
        seq = b"".join(chunks)
        runs = [(m.start(), m.end()) for m in _N_RUN.finditer(seq)]
        if runs:
            gaps.setdefault(name, []).extend(runs)

    with opener(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                flush()
                tok = line[1:].split()
                name = tok[0].decode() if tok else ""
                ids.append(name)
                chunks = []
            else:
                chunks.append(line.strip().replace(b" ", b"").replace(b"\r", b""))   # like SimpleFastaParser
    flush()
    return ids, gaps


def write_fasta(path, names, lengths, n_runs, seed=1, width=80):
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for c, (n, l) in enumerate(zip(names, lengths)):
            seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(l))
            for k, (s, e) in enumerate(n_runs[c] or ()):
                seq[s:e] = ord("n") if k % 2 else ord("N")
            f.write(f">{n} synthetic\n")
            txt = seq.tobytes().decode()
            for i in range(0, len(txt), width):
                f.write(txt[i:i + width] + "\n")


def read_regions_bed(path):
    """GCI.py:905-910: dict contig -> [(start, end)], insertion ordered."""
    regions = {}
    with open(path, "r") as f:
        for line in f:
            target, start, end = line.strip().split("\t")
            regions.setdefault(target, []).append((int(start), int(end)))
    return regions


# ------------------------------------------------------------------------------------------------
# .depth.gz
# ------------------------------------------------------------------------------------------------

def _gzip_member(data: bytes, level: int) -> bytes:
    return gzip.compress(data, compresslevel=level, mtime=0)


def write_depth_gz(path, pieces, threads=1, level=6):
    """`pieces` yields byte strings (">name\\n" headers and blocks of "%d\\n" lines) in file order;
    each piece becomes one gzip member — the reference also writes a multi-member file
    (GCI.py:99-143), so parity is defined on the decompressed stream."""
    with open(path, "wb") as f, ThreadPoolExecutor(max(1, int(threads))) as pool:
        pending = []
        for piece in pieces:
            pending.append(pool.submit(_gzip_member, bytes(piece), level))
            while len(pending) > 2 * max(1, int(threads)):
                f.write(pending.pop(0).result())
        for fut in pending:
            f.write(fut.result())


def read_depth_gz(path, threads=0):
    """utility/GCI_score.py:11-39 — {name: int32 array} in file order (native parser; `read_depth_gz_py` is the
    pure-Python cross-check)."""
    from . import io_native
    return io_native.read_depth_gz(path, threads)


_DEPTH_HEADER = re.compile(rb"^[ \t\r\f\v]*>[^\n]*$", re.M)


def read_depth_gz_py(path):
    """utility/GCI_score.py:11-39 without the native library: a line whose stripped text starts with '>' opens the
    target `item.split('>')[-1]`, every other line is one integer.  (Blank lines, on which the reference raises,
    are skipped.)  Independent cross-check of the native parser; the tests read the product's files with it."""
    with gzip.open(path, "rb") as f:
        data = f.read()
    out, cur, pos = {}, None, 0

    def numbers(body):
        tok = body.split()
        if tok:
            out[cur].append(np.array(tok, dtype=np.int64))      # KeyError on numbers before the first header

    for m in _DEPTH_HEADER.finditer(data):
        numbers(data[pos:m.start()])
        cur = m.group().strip().split(b">")[-1].decode()
        out[cur] = []
        pos = m.end()
    numbers(data[pos:])
    return {k: (np.concatenate(v) if v else np.zeros(0, np.int64)).astype(np.int32) for k, v in out.items()}
