"""Host-side mirror of the reference's hot-path interface (GCI.py) on top of libgci_cuda.so.

Same function names, argument meaning, printed progress lines, output files and `sys.exit`
texts as the reference, so a `GCI.py` user can switch over:

    filter()                 GCI.py:172-312      -> gci_filter + gci_depth
    write_depth()            GCI.py:99-143       -> gci_depth_text + gzip members
    merge_gaps_depths()      GCI.py:315-329      -> gci_mask_gaps
    merge_two_type_depth()   GCI.py:332-353      -> gci_merge_max
    collapse_depth_range()   GCI.py:356-390      -> gci_scan / gci_scan_windows
    merge_depth()            GCI.py:393-419
    compute_index()          GCI.py:522-657      -> gci_score_terms (+ log2/round on the host)
    GCI()                    GCI.py:897-1028

Depth dictionaries are `DeviceDepths` mappings whose arrays live on the GPU and are fetched on
access.  File arguments may be paths (decoded by gci_b200.io) or already-decoded
`AlnTable` / `PafTable` objects.  Nothing here computes on the CPU what the library computes on
the GPU, and nothing imports `oracle/`.
"""
from __future__ import annotations

import os
import sys
from collections.abc import Mapping
from math import log2

import numpy as np

from . import io as gio
from ._lib import Context, NO_FLAGS, TRACK_HIFI, TRACK_NANO, TRACK_MERGED
from .records import AlnTable, PafTable



# ------------------------------------------------------------------------------------------------
# session
# ------------------------------------------------------------------------------------------------
class Session:
    """One GPU context plus the contig table it was configured with."""

    def __init__(self, device: int = 0):
        self.ctx = Context(device)
        self.names = None
        self.lengths = None
        self.selected = None
        self.scan_hint = None          # (lo, hi): thresholds the driver will scan with -> fused flags
        self._n_runs_key = None
        self.plan = None               # gci_b200.sharded.ShardPlan of a multi-GPU run (torchrun), else None
        self.run_selected = None

    def close(self):
        self.ctx.close()

    def configure(self, names, lengths, chrs_list):
        """contig table of the run.  Under torchrun (WORLD_SIZE > 1) the selected contigs are dealt to the ranks
        (gci_b200.sharded): `selected` then means selected AND owned by this rank, `run_selected` selected anywhere."""
        names = list(names)
        lengths = [int(x) for x in lengths]
        run_selected = [(n in chrs_list) if len(chrs_list) > 0 else True for n in names]
        plan = None
        if _world() > 1:
            from . import sharded
            plan = sharded.make_plan(_rank(), _world(), lengths, run_selected)
            if sum(run_selected) < _world():
                sys.exit(f'ERROR!!! {_world()} GPUs for {sum(run_selected)} chromosomes: use at most one GPU per chromosome')
        selected = run_selected if plan is None else [bool(x) for x in plan.owned]
        if self.names == names and self.lengths == lengths and self.selected == selected:
            return
        self.ctx.set_contigs(lengths, selected)
        order = sorted(range(len(names)), key=lambda i: names[i])
        rank = np.empty(len(names), np.int32)
        rank[order] = np.arange(len(names), dtype=np.int32)
        self.ctx.set_name_rank(rank)
        if plan is not None:
            self.ctx.shard_config(plan.rank, plan.world, plan.owner, plan.selected)
            self._exchange = (0, 0)
        self.names, self.lengths, self.selected, self.run_selected, self.plan = names, lengths, selected, run_selected, plan
        self._n_runs_key = None

    def ensure_exchange(self, n_reads, n_bam_files):
        """(re)allocate the ranks' exchange areas when a read set needs more room; collective"""
        have_reads, have_files = getattr(self, "_exchange", (0, 0))
        if n_reads <= have_reads and n_bam_files <= have_files:
            return
        from . import sharded
        want = (max(n_reads, have_reads), max(n_bam_files, have_files, 1))
        self.ctx.shard_close()             # nobody may still map an area that is about to be reallocated
        _dist().barrier()
        handle = self.ctx.shard_alloc(want[0], want[1])
        sharded.open_over_process_group(self.ctx, self.plan, handle)
        self._exchange = want

    def run_selected_names(self):
        sel = getattr(self, "run_selected", None) or self.selected
        return [n for n, s in zip(self.names, sel) if s]

    @property
    def index(self):
        return {n: i for i, n in enumerate(self.names)}

    def selected_names(self):
        return [n for n, s in zip(self.names, self.selected) if s]

    def set_n_runs(self, Ns_bed):
        key = repr(sorted((k, tuple(v)) for k, v in Ns_bed.items()))
        if key == self._n_runs_key:
            return
        idx = self.index
        c, s, e = [], [], []
        for target, segments in Ns_bed.items():
            if target in idx:
                for seg in segments:
                    c.append(idx[target]); s.append(seg[0]); e.append(seg[1])
        self.ctx.set_n_runs(c, s, e)
        self._n_runs_key = key


def _world():
    return int(os.environ.get("WORLD_SIZE", "1"))


def _rank():
    return int(os.environ.get("RANK", "0"))


def _dist():
    """torch.distributed helpers of a multi-GPU run (one process per GPU under torchrun), initialised on first use"""
    from . import dist as D
    if _world() > 1 and not D.is_dist():
        D.init()
    return D


def _by_rank_merge(local: dict, order):
    """per-contig results of every rank -> one dict in `order` (every contig lives on exactly one rank)"""
    merged = {}
    for part in _dist().gather_objects(local):
        merged.update(part)
    return {k: merged[k] for k in order if k in merged}


def _open_out(path, mode='w'):
    """output files are written by rank 0 of a multi-GPU run; the other ranks compute the same text and drop it"""
    return open(path if _rank() == 0 else os.devnull, mode)


def _exists(path):
    return _rank() == 0 and os.path.exists(path)


_default_session = None


def default_session() -> Session:
    global _default_session
    if _default_session is None:
        _default_session = Session(int(os.environ.get("GCI_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
    return _default_session


def set_default_session(s):
    global _default_session
    _default_session = s


class DeviceDepths(Mapping):
    """dict[contig -> np.ndarray(int64)] whose values live in a GPU depth track."""

    def __init__(self, session: Session, track: int):
        self.session, self.track = session, track
        self._names = session.selected_names()

    def __getitem__(self, name):
        if name not in self._names:
            raise KeyError(name)
        return self.session.ctx.fetch_depth_narrow(self.track, self.session.index[name]).astype(np.int64)

    def fetch_i32(self, name):
        return self.session.ctx.fetch_depth(self.track, self.session.index[name])

    def order_all(self):
        """contigs of the run in output order (in a multi-GPU run the mapping itself only holds the owned ones)"""
        return getattr(self, "_all_names", None) or list(self._names)

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def mean_depth(self):
        """np.mean over all contigs (GCI.py:862-868) from the per-contig sums kept on the GPU."""
        sums = self.session.ctx.depth_sums(self.track)
        tot_len = sum(l for l, s in zip(self.session.lengths, self.session.selected) if s)
        return float(sums.sum()) / tot_len if tot_len else float("nan")


class DeviceBed(dict):
    """dict[contig -> [(start, end), ...]] that remembers which GPU track holds the same intervals."""

    session = None
    track = None
    scan_id = None


def _adopt(depths, session=None, track=TRACK_HIFI) -> DeviceDepths:
    """Host dict of depth arrays (e.g. parsed from a .depth.gz) -> GPU track."""
    if isinstance(depths, DeviceDepths):
        return depths
    session = session or default_session()
    names = list(depths.keys())
    lengths = [len(depths[n]) for n in names]
    session.configure(names, lengths, [])
    for i, n in enumerate(names):
        session.ctx.load_depth(track, i, np.ascontiguousarray(depths[n], dtype=np.int32))
    return DeviceDepths(session, track)


# ------------------------------------------------------------------------------------------------
# L3: depth file
# ------------------------------------------------------------------------------------------------
def write_depth(directory='.', prefix='GCI', depths={}, threads=1):
    """GCI.py:99-143: `>name` then one decimal per line, gzip (multi-member like the reference's).
    Text formatting, DEFLATE and CRC-32 all run on the GPU (gci_depth_gzip_track: one pass over the track); only
    compressed bytes reach the host."""
    depths = _adopt(depths)
    session, ctx = depths.session, depths.session.ctx
    idx = session.index
    heads = {target: f'>{target}\n'.encode('utf-8') for target in depths.keys()}
    if session.plan is not None:
        # every rank compresses the contigs it owns; rank 0 writes the members in the reference's contig order
        D = _dist()
        blob, off = ctx.depth_gzip_track(depths.track, [heads.get(n, b'') for n in session.names])
        parts = D.allgather_varlen(blob)
        offs = D.gather_objects(off.tolist())
        if _rank() == 0:
            with open(f'{directory}/{prefix}.depth.gz', 'wb') as f:
                for target in depths.order_all():
                    c = idx[target]
                    r = session.plan.owner[c]
                    f.write(parts[r][offs[r][c]:offs[r][c + 1]].tobytes())
        return
    with open(f'{directory}/{prefix}.depth.gz', 'wb') as f:
        if list(depths.keys()) == session.selected_names() and all(len(h) <= 4096 for h in heads.values()):
            blob, _ = ctx.depth_gzip_track(depths.track, [heads.get(n, b'') for n in session.names])
            f.write(blob.tobytes())
            return
        for target in depths.keys():         # another contig order than the session's (or an over-long name)
            c = idx[target]
            n = int(session.lengths[c])
            head = heads[target]
            if len(head) > 4096:
                f.write(gio._gzip_member(head, 6))
                head = b''
            f.write(ctx.depth_gzip(depths.track, c, 0, n, head).tobytes())


# ------------------------------------------------------------------------------------------------
# L2 + L3: filter
# ------------------------------------------------------------------------------------------------
def _remap_by_name(t, names, base_names):
    """A later file's `ref_id`s live in that file's own @SQ order: bring them onto `base_names` by NAME, the way
    the reference addresses contigs (`fetch(contig=target)`, GCI.py:150-151).  Contigs the base does not know are
    never fetched by the reference (-> -1)."""
    names = list(names)
    if names == list(base_names):
        return t
    base = {n: i for i, n in enumerate(base_names)}
    lut = np.full(len(names) + 1, -1, np.int32)             # last slot: ref_id outside the header stays -1
    for i, n in enumerate(names):
        lut[i] = base.get(n, -1)
    rid = np.asarray(t.ref_id, np.int64)
    rid = np.where((rid < 0) | (rid >= len(names)), len(names), rid)
    out = AlnTable(lut[rid], t.ref_start, t.mapq, t.flag, t.nm, t.qlen, t.read_id, t.cigar_off, t.cigar)
    return out


def _remap_paf_by_name(t, names, base_names):
    if list(names) == list(base_names):
        return t
    base = {n: i for i, n in enumerate(base_names)}
    lut = np.full(len(names) + 1, -1, np.int32)
    for i, n in enumerate(names):
        lut[i] = base.get(n, -1)
    rid = np.asarray(t.ref_id, np.int64)
    rid = np.where((rid < 0) | (rid >= len(names)), len(names), rid)
    return PafTable(t.read_id, t.qlen, t.qstart, t.qend, lut[rid], t.tstart, t.tend, t.nmatch, t.alnlen, t.mapq)


def _prune_for_gates(t, selected, map_qual):
    """Records the gates drop before they look at anything else never need to reach the GPU: unmapped / secondary /
    supplementary records and records below `-mq` (GCI.py:153-156: skipped before the NM tag is read, so they cannot
    raise either) and records on contigs that are never fetched (`--chrs`, :151, :202-207).  Their CIGARs are most of
    what a BAM with many supplementary alignments uploads.  A table that keeps more than 90 % of its records is
    passed on as it is (the copy would cost more than it saves)."""
    n = t.n_records
    if n == 0:
        return t
    sel = np.asarray(selected, bool)
    rid = np.asarray(t.ref_id, np.int64)
    ok = (rid >= 0) & (rid < len(sel))
    ok &= sel[np.clip(rid, 0, len(sel) - 1)]
    ok &= (np.asarray(t.flag, np.int64) & (0x4 | 0x100 | 0x800)) == 0
    ok &= np.asarray(t.mapq, np.int64) >= int(map_qual)
    kept = int(ok.sum())
    if kept > 0.9 * n:
        return t
    return t.take(np.flatnonzero(ok))


def _load_inputs(paf_files, bam_files, threads=1):
    """Decode / collect the files of one read type.  Returns (names, lengths, paf tables, bam tables, n_reads).
    Contig names and lengths are the first BAM's (GCI.py:201); every later BAM is remapped onto them by name."""
    intern = gio.NameTable()
    bams, pafs = [], []
    names = lengths = None
    for f in bam_files:
        if isinstance(f, AlnTable):
            t = f
            n, l = getattr(f, "contig_names", None), getattr(f, "contig_lengths", None)
            if n is None:
                raise ValueError("in-memory AlnTable needs .contig_names / .contig_lengths")
        else:
            n, l, t = gio.read_bam(f, intern, threads)
        if names is None:
            names, lengths = list(n), [int(x) for x in l]     # header of the first BAM (GCI.py:201)
        else:
            t = _remap_by_name(t, n, names)
        bams.append(t)
    for f in paf_files:
        pafs.append(f if isinstance(f, PafTable) else gio.read_paf(f, names, intern))
    n_reads = len(intern)
    for t in bams + pafs:
        if t.n_records:
            n_reads = max(n_reads, int(t.read_id.max()) + 1)
    return names, lengths, pafs, bams, n_reads


def filter(paf_files=[], bam_files=[], prefix='GCI', map_qual=30, mq_cutoff=50, iden_percent=0.9, clip_percent=0.1,
           ovlp_percent=0.9, flank_len=15, directory='.', force=False, log_reads_type='', chrs_list=[], threads=1,
           session=None, write=True):
    """GCI.py:172-312.  Returns (depths, targets_length)."""
    if _exists(f'{directory}/{prefix}.depth.gz') and force == False:
        sys.exit(f'ERROR!!! The file "{directory}/{prefix}.depth.gz" exists\nPlease use "-f" or "--force" to rewrite')
    print(f'Filtering {log_reads_type} alignment files ...')
    session = session or default_session()
    ctx = session.ctx
    names, lengths, pafs, bams, n_reads = _load_inputs(paf_files, bam_files, threads)
    out_order = None
    if session.names is not None and names != session.names and sorted(names) == sorted(session.names) \
            and dict(zip(names, lengths)) == dict(zip(session.names, session.lengths)):
        # same contigs in another @SQ order (e.g. the ONT BAM of a HiFi + ONT run): the tracks already held by the
        # session stay valid, the records move onto its order by name; the outputs keep THIS file's order, like
        # the reference's `depths` dict (GCI.py:201-207)
        bams = [_remap_by_name(t, names, session.names) for t in bams]
        pafs = [_remap_paf_by_name(t, names, session.names) for t in pafs]
        out_order = list(names)
        names, lengths = list(session.names), list(session.lengths)
    session.configure(names, lengths, chrs_list)
    if out_order is None:
        out_order = names
    sel = dict(zip(names, session.run_selected if session.plan is not None else session.selected))
    targets_length = {n: l for n, l in zip(out_order, (dict(zip(names, lengths))[x] for x in out_order)) if sel[n]}
    track = TRACK_NANO if log_reads_type == 'ONT' else TRACK_HIFI

    gate_sel = session.run_selected if session.plan is not None else session.selected
    bams = [_prune_for_gates(t, gate_sel, map_qual) for t in bams]
    lo, hi = session.scan_hint if session.scan_hint is not None else (NO_FLAGS, NO_FLAGS)
    if session.plan is None:
        ctx.reads_begin(n_reads)
        for t in pafs:                       # files = paf_lines + samfile_dicts (GCI.py:272)
            ctx.upload_paf(t)
        for t in bams:
            ctx.upload_bam(t)
        ctx.filter(map_qual, mq_cutoff, iden_percent, clip_percent, ovlp_percent)
        ctx.depth(track, flank_len, lo, hi)
        depths = DeviceDepths(session, track)
        depths._names = list(targets_length.keys())
    else:
        # multi-GPU: this rank holds the records of its contigs and the PAF lines of its home reads; the library
        # moves winners and survivors between the ranks inside the call (csrc/shard.cu)
        from . import sharded
        session.ensure_exchange(n_reads, len(bams))
        ctx.reads_begin(n_reads)
        for t in pafs:
            ctx.upload_paf(sharded.shard_paf(t, session.plan))
        for t in bams:
            ctx.upload_bam(sharded.shard_bam(t, session.plan))
        flo, fhi = (lo, hi) if lo != NO_FLAGS else (-1, 0)
        ctx.pipeline(track, sum(session.selected), map_qual, mq_cutoff, iden_percent, clip_percent, ovlp_percent,
                     flank_len, flo, fhi, 0.005)
        depths = DeviceDepths(session, track)
        depths._all_names = list(targets_length.keys())
        depths._names = [n for n in depths._all_names if session.selected[session.index[n]]]

    print(f'Filtering {log_reads_type} alignment files done!!!')
    if write:
        print(f'Writing depths into "{directory}/{prefix}.depth.gz" ...')
        write_depth(directory, prefix, depths, threads)
        print(f'Writing depths done!!!\n\n')
    return depths, targets_length


def merge_gaps_depths(depths={}, Ns_bed=None):
    """GCI.py:315-329."""
    if Ns_bed != None:
        depths = _adopt(depths)
        depths.session.set_n_runs(Ns_bed)
        depths.session.ctx.mask_gaps(depths.track)
    return depths


def merge_two_type_depth(hifi_depths={}, nano_depths={}, prefix='GCI_two_type', directory='.', force=False, threads=1,
                         write=True):
    """GCI.py:332-353."""
    print('Merging HiFi and ONT depth file ...')
    if _exists(f'{directory}/{prefix}.depth.gz') and force == False:
        sys.exit(f'ERROR!!! The file "{directory}/{prefix}.depth.gz" exists\nPlease use "-f" or "--force" to rewrite')
    if not (isinstance(hifi_depths, DeviceDepths) and isinstance(nano_depths, DeviceDepths)
            and hifi_depths.session is nano_depths.session):
        raise TypeError("merge_two_type_depth expects the depth mappings returned by filter() of one session")
    session = hifi_depths.session
    lo, hi = session.scan_hint if session.scan_hint is not None else (NO_FLAGS, NO_FLAGS)
    session.ctx.merge_max(hifi_depths.track, nano_depths.track, TRACK_MERGED, lo, hi)
    merged = DeviceDepths(session, TRACK_MERGED)
    merged._names = list(hifi_depths.keys())
    merged._all_names = getattr(hifi_depths, "_all_names", None)
    if write:
        write_depth(directory, prefix, merged, threads)
    print('Merging HiFi and ONT depth file done!!!\n\n')
    return merged


# ------------------------------------------------------------------------------------------------
# L4: gap scan
# ------------------------------------------------------------------------------------------------
_scan_counter = [0]


def collapse_depth_range(depths={}, leftmost=-1, rightmost=0, flank_len=15, start_pos=0):
    """GCI.py:356-390 over every contig of `depths` (start_pos must be 0 here; the regions variant
    goes through `_collapse_regions`)."""
    if start_pos != 0:
        raise ValueError("use regions scoring for start_pos != 0")
    depths = _adopt(depths)
    session, ctx = depths.session, depths.session.ctx
    ctx.scan(depths.track, leftmost, rightmost, flank_len)
    owners = session.selected_names()
    s, e, off = ctx.fetch_intervals(depths.track, len(owners))
    bed = DeviceBed()
    sl, el = s.tolist(), e.tolist()
    by_name = {}
    for o, name in enumerate(owners):
        a, b = int(off[o]), int(off[o + 1])
        by_name[name] = list(zip(sl[a:b], el[a:b]))
    if session.plan is not None:         # every rank scanned its contigs: the BED of the run, on every rank
        by_name = _by_rank_merge(by_name, depths.order_all())
    for name in depths.order_all():      # the depth mapping's own order (a permuted @SQ order keeps its file's)
        bed[name] = by_name[name]
    _scan_counter[0] += 1
    bed.session, bed.track, bed.scan_id = session, depths.track, _scan_counter[0]
    session.__dict__.setdefault("_last_scan", {})[depths.track] = (bed.scan_id, len(s))
    return bed


def merge_depth(depths={}, prefix='GCI', threshold=0, flank_len=15, directory='.', force=False, log_reads_type=''):
    """GCI.py:393-419."""
    print(f'Getting {log_reads_type} issues bed file detected by GCI ...')
    if _exists(f'{directory}/{prefix}.{threshold}.depth.bed') and force == False:
        sys.exit(f'ERROR!!! The file "{directory}/{prefix}.{threshold}.depth.bed" exists\nPlease use "-f" or "--force" to rewrite')
    merged_depths_bed = collapse_depth_range(depths, -1, threshold, flank_len, 0)
    if _rank() == 0:
        with open(f'{directory}/{prefix}.{threshold}.depth.bed', 'w') as f:
            for target, segments in merged_depths_bed.items():
                f.write(''.join(f'{target}\t{s}\t{e}\n' for s, e in segments))
    print(f'Getting {log_reads_type} issues bed file done!!!\n\n')
    return merged_depths_bed


# ------------------------------------------------------------------------------------------------
# L5: score
# ------------------------------------------------------------------------------------------------
def compute_n50(lengths=[]):
    """GCI.py:465-480 for the tiny host-side lists (contig lengths, region lengths)."""
    n50 = 0
    lengths = sorted((int(x) for x in lengths), reverse=True)
    total = sum(lengths)
    cum = 0
    for x in lengths:
        cum += x
        if cum >= total / 2:
            n50 = x
            break
    return n50


def _gci(obs_n50, exp_n50, obs_num_ctg, exp_num_ctg):
    if obs_num_ctg == 0:                                          # GCI.py:601-604
        return 0
    return round(100 * log2(obs_n50 / exp_n50 + 1) / log2(obs_num_ctg / exp_num_ctg + 1), 4)


def _terms_for_bed(bed, targets_length, dist_percent, flank_len, session):
    """(n50[], n_ctg[]) per contig of targets_length + the all-contigs entry, from the GPU."""
    session = getattr(bed, "session", None) or session or default_session()
    ctx = session.ctx
    owners = list(targets_length.keys())
    last = session.__dict__.get("_last_scan", {})
    if session.plan is not None:
        # multi-GPU: every rank scores the contigs it owns (the intervals of its last scan are still on its GPU);
        # the genome row needs every curated length (GCI.py:572-587): they are few and travel as Python lists
        mine = session.selected_names()
        n_iv = last[bed.track][1]
        n50, nctg, lens, off = ctx.score_terms(bed.track, len(mine), n_iv, dist_percent, flank_len)
        local = {name: (int(n50[o]), int(nctg[o]), lens[off[o]:off[o + 1]].tolist()) for o, name in enumerate(mine)}
        full = _by_rank_merge(local, owners)
        all_len = [x for name in owners for x in full[name][2]]
        return (np.array([full[name][0] for name in owners] + [compute_n50(all_len)], np.int64),
                np.array([full[name][1] for name in owners] + [sum(full[name][1] for name in owners)], np.int64))
    if isinstance(bed, DeviceBed) and bed.track is not None and last.get(bed.track, (None,))[0] == bed.scan_id \
            and owners == session.selected_names():
        track, n_iv = bed.track, last[bed.track][1]
    else:
        # intervals from elsewhere (a BED file): load them next to track 0's depth
        if session.names is None or any(t not in session.index for t in owners):
            session.configure(owners, [targets_length[t] for t in owners], [])
        idx = session.index
        off = np.concatenate([[0], np.cumsum([len(bed[t]) for t in owners])])
        flat = [seg for t in owners for seg in bed[t]]
        track, n_iv = TRACK_HIFI, len(flat)
        ctx.load_intervals(track, [idx[t] for t in owners], off, [s for s, _ in flat], [e for _, e in flat])
        session.__dict__.setdefault("_last_scan", {})[track] = (None, n_iv)
    n50, nctg, _, _ = ctx.score_terms(track, len(owners), n_iv, dist_percent, flank_len)
    return n50, nctg


def compute_index(targets_length={}, prefix='GCI', directory='.', force=False, merged_depths_bed_list=[], type_list=[],
                  flank_len=15, dist_percent=0.005, regions_bed={}, depths_list=[], threshold=0, chrs_list=[],
                  session=None):
    """GCI.py:522-657."""
    if _exists(f'{directory}/{prefix}.gci') and force == False:
        sys.exit(f'ERROR!!! The file "{directory}/{prefix}.gci" exists\nPlease use "-f" or "--force" to rewrite')
    with _open_out(f'{directory}/{prefix}.gci', 'w') as f:
        pass
    if len(regions_bed) > 0:
        if _exists(f'{directory}/{prefix}.regions.gci') and force == False:
            sys.exit(f'ERROR!!! The file "{directory}/{prefix}.regions.gci" exists\nPlease use "-f" or "--force" to rewrite')
        with _open_out(f'{directory}/{prefix}.regions.gci', 'w') as f:
            f.write('Chromosome\tStart\tEnd\t' + '\t'.join(type_list) + '\n')

    print('Computing Theoretical minimum N50 and contigs number ...')
    all_label = 'Genome' if len(chrs_list) == 0 else 'All_chromosomes'
    exp_lengths = [length for length in targets_length.values()]
    exp_n50_all = compute_n50(exp_lengths)
    print('Computing Theoretical minimum N50 and contigs number done!!!')

    for i, merged_depths_bed in enumerate(merged_depths_bed_list):
        print(f'Computing Curated N50 and contigs number for {type_list[i]} ...')
        n50, nctg = _terms_for_bed(merged_depths_bed, targets_length, dist_percent, flank_len, session)
        print(f'Computing Curated N50 and contigs number for {type_list[i]} done!!!')

        print(f'Writing results to {directory}/{prefix}.gci ...')
        with _open_out(f'{directory}/{prefix}.gci', 'a') as f:
            f.write(f'{type_list[i]}:\n')
            f.write('Chromosome\tTheoretical maximum N50\tCurated N50\tTheoretical minimum contigs number\tCurated contigs number\tGCI score\n')
            for o, (target, length) in enumerate(targets_length.items()):
                obs_n50, obs_num_ctg = int(n50[o]), int(nctg[o])
                f.write(f'{target}\t{length}\t{obs_n50}\t1\t{obs_num_ctg}\t{_gci(obs_n50, length, obs_num_ctg, 1)}\n')
            obs_n50, obs_num_ctg = int(n50[-1]), int(nctg[-1])
            gci = _gci(obs_n50, exp_n50_all, obs_num_ctg, len(exp_lengths))
            f.write(f'{all_label}\t{exp_n50_all}\t{obs_n50}\t{len(exp_lengths)}\t{obs_num_ctg}\t{gci}\n')
            f.write('-' * 136 + '\n\n\n')
        print(f'Writing results to {directory}/{prefix}.gci done!!!\n\n')

    if len(regions_bed) > 0:
        print('Computing GCI scores for regions ...')
        windows = [(target, seg[0], seg[1]) for target, segments in regions_bed.items() for seg in segments]
        region_all_lengths = []
        for target, start, end in windows:
            if end - start > 0:
                region_all_lengths.append(end - start)
            else:
                print(f'Warning!!! The region "{target}:{start}-{end}" is not available', file=sys.stderr)
        per_type = []
        for depths in depths_list:
            depths = _adopt(depths, session)
            ses, ctx = depths.session, depths.session.ctx
            idx = ses.index
            if ses.plan is not None:
                # multi-GPU: every rank scans the regions lying on its contigs; All_regions needs all curated lengths
                mine = [w for w, (t, _, _) in enumerate(windows) if ses.selected[idx[t]]]
                local = {}
                if mine:
                    n_iv = ctx.scan_windows(depths.track, [idx[windows[w][0]] for w in mine], [windows[w][1] for w in mine],
                                            [windows[w][2] for w in mine], -1, threshold)
                    ses.__dict__.setdefault("_last_scan", {})[depths.track] = (None, n_iv)
                    a50, actg, lens, off = ctx.score_terms(depths.track, len(mine), n_iv, dist_percent, 0)
                    local = {w: (int(a50[k]), int(actg[k]), lens[off[k]:off[k + 1]].tolist()) for k, w in enumerate(mine)}
                full = _by_rank_merge(local, range(len(windows)))
                all_len = [x for w in range(len(windows)) for x in full[w][2]]
                per_type.append((np.array([full[w][0] for w in range(len(windows))] + [compute_n50(all_len)], np.int64),
                                 np.array([full[w][1] for w in range(len(windows))] +
                                          [sum(full[w][1] for w in range(len(windows)))], np.int64)))
                continue
            n_iv = ctx.scan_windows(depths.track, [idx[t] for t, _, _ in windows], [s for _, s, _ in windows],
                                    [e for _, _, e in windows], -1, threshold)
            ses.__dict__.setdefault("_last_scan", {})[depths.track] = (None, n_iv)
            per_type.append(ctx.score_terms(depths.track, len(windows), n_iv, dist_percent, 0)[:2])
        with _open_out(f'{directory}/{prefix}.regions.gci', 'a') as f:
            for w, (target, start, end) in enumerate(windows):
                gci = [_gci(int(n50[w]), end - start, int(nctg[w]), 1) for n50, nctg in per_type]
                f.write(f'{target}\t{start}\t{end}\t' + '\t'.join(map(str, gci)) + '\n')
            region_all_exp_n50 = compute_n50(region_all_lengths)
            region_all_gci = [_gci(int(n50[-1]), region_all_exp_n50, int(nctg[-1]), len(region_all_lengths))
                              for n50, nctg in per_type]
            f.write('-' * 136 + '\n\n\n')
            f.write(f'All_regions\t*\t*\t' + '\t'.join(map(str, region_all_gci)) + '\n')
        print('Computing GCI scores for regions done!!!\n\n')


# ------------------------------------------------------------------------------------------------
# L1: driver
# ------------------------------------------------------------------------------------------------
def get_Ns_ref(reference=None, prefix='GCI', directory='.', force=False, _parsed=None):
    """GCI.py:18-46."""
    ids, Ns_bed = _parsed if _parsed is not None else _read_reference(reference)
    if len(Ns_bed) > 0:
        if _exists(f'{directory}/{prefix}.gaps.bed') and force == False:
            sys.exit(f'ERROR!!! The file "{directory}/{prefix}.gaps.bed" exists\nPlease use "-f" or "--force" to rewrite')
        with _open_out(f'{directory}/{prefix}.gaps.bed', 'w') as f:
            for target, segments in Ns_bed.items():
                for segment in segments:
                    f.write(f'{target}\t{segment[0]}\t{segment[1]}\n')
        return Ns_bed, f'{directory}/{prefix}.gaps.bed'
    return None, None


def _read_reference(reference):
    if isinstance(reference, dict):          # in-memory: {"ids": [...], "gaps": {name: [(s,e)...]}}
        return list(reference["ids"]), {k: list(v) for k, v in reference["gaps"].items() if len(v)}
    return gio.read_fasta_gaps(reference)


def _is_bam(f):
    return isinstance(f, AlnTable) or (isinstance(f, str) and f.endswith('.bam'))


def _header(f):
    if isinstance(f, AlnTable):
        return dict(zip(f.contig_names, (int(x) for x in f.contig_lengths)))
    n, l = gio.read_bam_header(f)
    return dict(zip(n, l))


def _load_regions(regions):
    """GCI.py:902-912."""
    if regions is None:
        return {}
    if isinstance(regions, dict):
        return regions
    if not (os.path.exists(regions) and os.access(regions, os.R_OK)):
        sys.exit(f'ERROR!!! "{regions}" is not an available file')
    return gio.read_regions_bed(regions)


def _prepare_directory(directory, prefix):
    """GCI.py:914-925."""
    if directory.endswith('/'):
        directory = '/'.join(directory.split('/')[:-1])
    if not os.path.exists(directory):
        os.makedirs(directory, exist_ok=True)      # (several ranks of a multi-GPU run may get here together)
    else:
        for mode, what in ((os.R_OK, 'read'), (os.W_OK, 'write')):
            if not os.access(directory, mode):
                sys.exit(f'ERROR!!! The path "{directory}" is unable to {what}')
    if prefix.endswith('/'):
        sys.exit(f'ERROR!!! The prefix "{prefix}" is not allowed')
    return directory


def _check_names(ref_refs, chrs_list, regions_bed):
    """GCI.py:942-952."""
    for what, names in (('--chrs', chrs_list), ('--regions', list(regions_bed.keys()))):
        for i in names:
            if i not in ref_refs:
                sys.exit(f'ERROR!!! Chromosome "{i}" provided by `{what}` is not in the reference')
    if len(chrs_list) > 0 and len(regions_bed) > 0 and not all(i in chrs_list for i in regions_bed.keys()):
        sys.exit(f'ERROR!!! Chromosomes in the regions bed file are inconsistent with the provided list of chromosomes\nPlease read the help message use "-h" or "--help"')


class _ReadType:
    """The alignment files of one read type, split like GCI.py:959-980."""

    def __init__(self, files, label, ref_refs):
        self.given = files is not None
        self.bam, self.paf, self.refs_lengths = [], [], {}
        if not self.given:
            return
        for file in files:
            if _is_bam(file):
                self.bam.append(file)
                self.refs_lengths = _header(file)          # the last BAM's header, like the reference
            else:
                self.paf.append(file)
        if set(self.refs_lengths.keys()) != set(ref_refs):
            sys.exit(f'ERROR!!! The targets in {label} alignment files are inconsistent with the reference file\nPlease check both {label} alignment files and the reference')


def GCI(hifi=[], nano=[], directory='.', prefix='GCI', map_qual=30, mq_cutoff=50, iden_percent=0.9, ovlp_percent=0.9,
        clip_percent=0.1, flank_len=15, threshold=0, plot=False, depth_min=0.1, depth_max=4.0, window_size=50000,
        image_type='png', force=False, dist_percent=0.005, reference=None, regions=None, chrs=None, threads=1,
        session=None):
    """The reference's driver (GCI.py:897-1028) on the GPU path: same keyword arguments, output files, progress
    lines and exit texts; plotting (`-p`) is outside the hot path and not provided."""
    chrs_list = chrs.strip().split(',') if chrs != None else []
    regions_bed = _load_regions(regions)
    directory = _prepare_directory(directory, prefix)
    if plot == True:
        # plotting is outside the hot path: the run goes ahead and writes every other output
        print('WARNING!!! Plotting (-p) is not provided by the GPU build and is skipped; run the reference\'s utility/plot_depth.py on the .depth.gz outputs', file=sys.stderr)

    parsed_ref = _read_reference(reference)       # one pass over the FASTA: record ids + N-runs
    ref_refs = parsed_ref[0]
    _check_names(ref_refs, chrs_list, regions_bed)
    H = _ReadType(hifi, 'hifi', ref_refs)
    N = _ReadType(nano, 'ont', ref_refs)

    print('Finding gaps ...')
    Ns_bed, Ns_bed_file = get_Ns_ref(reference, prefix, directory, force, _parsed=parsed_ref)
    if Ns_bed_file != None:
        print(f'Finding gaps done!!! The gaps are in {Ns_bed_file}\n\n')
    else:
        print('Finding gaps done!!! Awesome! No gaps were found!\n\n')

    session = session or default_session()
    session.scan_hint = (-1, threshold)      # the depth kernels emit the issue flags in the same pass
    gates = dict(map_qual=map_qual, mq_cutoff=mq_cutoff, iden_percent=iden_percent, clip_percent=clip_percent,
                 ovlp_percent=ovlp_percent, flank_len=flank_len, directory=directory, force=force,
                 chrs_list=chrs_list, threads=threads, session=session)

    def one_type(rt, out_prefix, log_name):
        depths, targets_length = filter(rt.paf, rt.bam, out_prefix, log_reads_type=log_name, **gates)
        return merge_gaps_depths(depths, Ns_bed), targets_length

    def issues(depths, out_prefix, log_name):
        return merge_depth(depths, out_prefix, threshold, flank_len, directory, force, log_name)

    def index(targets_length, beds, labels, tracks):
        compute_index(targets_length, prefix, directory, force, beds, labels, flank_len, dist_percent, regions_bed,
                      tracks, threshold, chrs_list, session=session)

    try:
        if not (H.given and N.given):
            rt, log_name, label = (H, 'HiFi', 'HiFi') if H.given else (N, 'ONT', 'Nano')
            depths, targets_length = one_type(rt, prefix, log_name)
            index(targets_length, [issues(depths, prefix, log_name)], [label], [depths])
        else:
            if set(H.refs_lengths.keys()) != set(N.refs_lengths.keys()):
                sys.exit(f'ERROR!!! The targets in hifi and nano alignment files are inconsistent\nPlease check the reference used in mapping both hifi and ont reads')
            for target, length in H.refs_lengths.items():
                if length != N.refs_lengths[target]:
                    sys.exit(f'ERROR!!! The element "{target}:{length}" in hifi alignment files are inconsistent with that in ont alignment files which is "{target}:{N.refs_lengths[target]}"\nPlease check the reference used in mapping both hifi and ont reads')
            hifi_depths, targets_length = one_type(H, prefix + '_hifi', 'HiFi')
            nano_depths, targets_length = one_type(N, prefix + '_nano', 'ONT')
            both = merge_two_type_depth(hifi_depths, nano_depths, prefix + '_two_type', directory, force, threads)
            both = merge_gaps_depths(both, Ns_bed)
            tracks = [hifi_depths, nano_depths, both]
            beds = [issues(d, prefix + sfx, log_name) for d, sfx, log_name in
                    zip(tracks, ('_hifi', '_nano', '_two_type'), ('HiFi', 'ONT', 'two_types'))]
            index(targets_length, beds, ['HiFi', 'Nano', 'HiFi + Nano'], tracks)
    finally:
        session.scan_hint = None
    print('GCI finished!!!\nBye!!!')
