"""Host side of a read set sharded over several GPUs (libgci_cuda's shard.cu does the exchange).

Contigs have an owner rank (LPT over their lengths, `dist.assign_contigs`), reads a home rank (block-cyclic in the read id, `home_rank`).
The host only DEALS the decoded records:

    shard_bam(table, plan)   records lying on the contigs this rank owns          (global read / contig ids)
    shard_paf(table, plan)   lines of the reads this rank is home to, read ids turned into home-local ids (`home_local`)

Everything else — PAF election, gates, the two all-to-all dispatches over NVLink peer memory (winners to the read
homes, survivors to the contig owners), merge, join, depth, scan, score — runs inside `Context.pipeline`.
Nothing here touches alignment data with numpy beyond those two selections.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .records import AlnTable, PafTable

_PAF_COLS = ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")


HOME_BLOCK = 64      # csrc/common.cuh GCI_HOME_BLOCK: reads are dealt to their home ranks in blocks of this many ids


def home_rank(read_id, world):
    """home rank of every read id (csrc/common.cuh home_rank)"""
    return (np.asarray(read_id).astype(np.int64) // HOME_BLOCK) % int(world)


def home_local(read_id, world):
    """index of every read id among the reads of its home (csrc/common.cuh home_local)"""
    q = np.asarray(read_id).astype(np.int64)
    return q // (HOME_BLOCK * int(world)) * HOME_BLOCK + q % HOME_BLOCK


@dataclass
class ShardPlan:
    rank: int
    world: int
    owner: list                 # owner[contig] = rank
    selected: np.ndarray        # bool[n_contigs]: contigs of the run (--chrs), on any rank

    @property
    def owned(self):
        """contigs this rank stores depth for: selected AND owned"""
        return self.selected & (np.asarray(self.owner) == self.rank)


def make_plan(rank, world, lengths, selected=None):
    from .dist import assign_contigs
    n = len(lengths)
    sel = np.ones(n, bool) if selected is None else np.asarray(selected, bool)
    # unselected contigs hold no depth: they weigh nothing, the selected ones are balanced by length
    owner = assign_contigs([int(l) if s else 0 for l, s in zip(lengths, sel)], None, world)
    return ShardPlan(rank, world, owner, sel)


def shard_bam(tab: AlnTable, plan: ShardPlan) -> AlnTable:
    own = np.asarray(plan.owner) == plan.rank
    rid = tab.ref_id
    keep = np.flatnonzero((rid >= 0) & (rid < len(own)) & own[np.clip(rid, 0, len(own) - 1)])
    return tab.take(keep)


def shard_paf(tab: PafTable, plan: ShardPlan) -> PafTable:
    keep = np.flatnonzero(home_rank(tab.read_id, plan.world) == plan.rank)
    cols = {k: getattr(tab, k)[keep] for k in _PAF_COLS}
    cols["read_id"] = home_local(cols["read_id"], plan.world).astype(np.uint32)
    return PafTable(*[cols[k] for k in _PAF_COLS])


def shard_paf_source(tab: PafTable, plan: ShardPlan, bam: AlnTable) -> PafTable:
    """The PAF lines a rank 'decoded' in a run where every rank decodes a part of the PAF: here the lines of the reads
    whose first BAM record lies on a contig this rank owns (any partition of the lines by read works; tests and the
    bench use it so that `deal_paf_over_process_group` has real work to do)."""
    own = np.asarray(plan.owner) == plan.rank
    n = int(max(tab.read_id.max(initial=0), bam.read_id.max(initial=0))) + 1
    first_contig = np.full(n, -1, np.int64)
    rid = bam.read_id[::-1]
    first_contig[rid] = bam.ref_id[::-1]                 # the FIRST record of every read wins
    src = first_contig[tab.read_id]
    keep = np.flatnonzero(np.where(src >= 0, own[np.clip(src, 0, len(own) - 1)], plan.rank == 0))
    return PafTable(*[getattr(tab, k)[keep] for k in _PAF_COLS])


def configure(ctx, plan: ShardPlan, lengths, name_rank, max_reads, max_bam_files=2, ipc=True):
    """contig table, owners and this rank's exchange area -> its CUDA IPC handle (uint8[64]).  The caller gathers
    every rank's handle and calls ctx.shard_open (processes) or ctx.shard_attach (contexts of one process)."""
    ctx.set_contigs(lengths, plan.owned)
    ctx.set_name_rank(name_rank)
    ctx.shard_config(plan.rank, plan.world, plan.owner, plan.selected)
    return ctx.shard_alloc(max_reads, max_bam_files, ipc)


def open_over_process_group(ctx, plan: ShardPlan, handle):
    """all-gather the IPC handles over torch.distributed and map the peers' areas"""
    from . import dist as D
    handles = np.zeros((plan.world, 64), np.int64)
    handles[plan.rank] = handle
    ctx.shard_open(D.allreduce(handles, "sum").astype(np.uint8))


def deal_paf_over_process_group(tab: PafTable, plan: ShardPlan) -> PafTable:
    """Every rank holds the PAF lines it decoded / generated (any reads); returns the lines of the reads THIS rank
    is home to (home-local ids), moved with one all-to-all.  Lines of one read keep their relative order as long as
    they all come from one rank."""
    from . import dist as D
    if not D.is_dist() or plan.world == 1:
        return shard_paf(tab, plan)
    import torch
    import torch.distributed as dist
    world = plan.world
    mat = np.stack([getattr(tab, k).astype(np.int64) for k in _PAF_COLS], axis=1)
    dst = home_rank(tab.read_id, world)
    order = np.argsort(dst, kind="stable")
    counts = np.bincount(dst, minlength=world).astype(np.int64)
    if dist.get_backend() != "nccl":
        parts = D.allgather_varlen(mat.reshape(-1))
        allm = np.concatenate([p.reshape(-1, len(_PAF_COLS)) for p in parts])
        mine = allm[home_rank(allm[:, 0], world) == plan.rank]
    else:
        dev = torch.device("cuda", torch.cuda.current_device())
        send = torch.from_numpy(np.ascontiguousarray(mat[order])).to(dev)
        c_in = torch.from_numpy(counts).to(dev)
        c_out = torch.empty_like(c_in)
        dist.all_to_all_single(c_out, c_in)
        n_out = [int(x) for x in c_out.cpu().tolist()]
        recv = torch.empty((sum(n_out), len(_PAF_COLS)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=n_out, input_split_sizes=[int(x) for x in counts])
        mine = recv.cpu().numpy()
    cols = {k: mine[:, i] for i, k in enumerate(_PAF_COLS)}
    cols["read_id"] = home_local(cols["read_id"], world)
    return PafTable(*[cols[k] for k in _PAF_COLS])


def deal_bam_over_process_group(tab: AlnTable, plan: ShardPlan) -> AlnTable:
    """Every rank holds BAM records of ANY contig (e.g. what it decoded from its part of a file, or what a second
    aligner placed elsewhere); returns the records lying on the contigs THIS rank owns, coordinate sorted, moved with
    two all-to-alls (record columns, CIGAR ops).  Records on contigs without an owner here are dropped."""
    from . import dist as D
    if not D.is_dist() or plan.world == 1:
        return shard_bam(tab, plan)
    import torch
    import torch.distributed as dist
    world = plan.world
    owner = np.asarray(plan.owner, np.int64)
    rid = tab.ref_id.astype(np.int64)
    ok = (rid >= 0) & (rid < len(owner))
    dst = np.where(ok, owner[np.clip(rid, 0, len(owner) - 1)], -1)
    keep = np.flatnonzero(dst >= 0)
    order = keep[np.argsort(dst[keep], kind="stable")]
    t = tab.take(order)
    n_ops = np.diff(t.cigar_off.astype(np.int64))
    mat = np.stack([t.ref_id.astype(np.int64), t.ref_start.astype(np.int64), t.mapq.astype(np.int64),
                    t.flag.astype(np.int64), t.nm.astype(np.int64), t.qlen.astype(np.int64),
                    t.read_id.astype(np.int64), n_ops], axis=1)
    d_sorted = dst[order]
    rec_counts = np.bincount(d_sorted, minlength=world).astype(np.int64)
    op_counts = np.bincount(d_sorted, weights=n_ops, minlength=world).astype(np.int64)
    if dist.get_backend() != "nccl":
        # CPU process groups (tests): every rank gathers everything and keeps what lies on its contigs
        recs = D.allgather_varlen(mat.reshape(-1))
        opss = D.allgather_varlen(t.cigar.astype(np.int64))
        dsts = D.allgather_varlen(d_sorted)
        rec_parts, op_parts = [], []
        for r_flat, o_flat, d_of in zip(recs, opss, dsts):
            r_mat = r_flat.reshape(-1, mat.shape[1])
            mine = d_of == plan.rank
            o_off = np.concatenate([[0], np.cumsum(r_mat[:, 7])])
            rec_parts.append(r_mat[mine])
            op_parts.extend(o_flat[o_off[i]:o_off[i + 1]] for i in np.flatnonzero(mine))
        rec = np.concatenate(rec_parts) if rec_parts else np.zeros((0, mat.shape[1]), np.int64)
        ops = (np.concatenate(op_parts) if op_parts else np.zeros(0, np.int64)).astype(np.uint32)
        off = np.concatenate([[0], np.cumsum(rec[:, 7])]).astype(np.uint64)
        got = AlnTable(rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5], rec[:, 6], off, ops)
        return got.take(np.lexsort((got.ref_start, got.ref_id)))
    dev = torch.device("cuda", torch.cuda.current_device())

    def a2a(data, counts, width):
        c_in = torch.from_numpy(counts).to(dev)
        c_out = torch.empty_like(c_in)
        dist.all_to_all_single(c_out, c_in)
        n_out = [int(x) for x in c_out.cpu().tolist()]
        send = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
        shape = (sum(n_out), width) if width else (sum(n_out),)
        recv = torch.empty(shape, dtype=send.dtype, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=n_out, input_split_sizes=[int(x) for x in counts])
        return recv.cpu().numpy()

    rec = a2a(mat, rec_counts, mat.shape[1])
    ops = a2a(t.cigar.astype(np.int64), op_counts, 0).astype(np.uint32)
    off = np.concatenate([[0], np.cumsum(rec[:, 7])]).astype(np.uint64)
    got = AlnTable(rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3], rec[:, 4], rec[:, 5], rec[:, 6], off, ops)
    return got.take(np.lexsort((got.ref_start, got.ref_id)))
