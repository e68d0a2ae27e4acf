"""Host side of a read set sharded over several GPUs (libgci_cuda's shard.cu does the exchange).

Contigs have an owner rank (LPT over their lengths, `dist.assign_contigs`), reads a home rank (read id % world).
The host only DEALS the decoded records:

    shard_bam(table, plan)   records lying on the contigs this rank owns          (global read / contig ids)
    shard_paf(table, plan)   lines of the reads this rank is home to, read ids divided by world (home-local ids)

Everything else — PAF election, gates, the two all-to-all dispatches over NVLink peer memory (winners to the read
homes, survivors to the contig owners), merge, join, depth, scan, score — runs inside `Context.pipeline`.
Nothing here touches alignment data with numpy beyond those two selections.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .records import AlnTable, PafTable

_PAF_COLS = ("read_id", "qlen", "qstart", "qend", "ref_id", "tstart", "tend", "nmatch", "alnlen", "mapq")


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
    keep = np.flatnonzero(tab.read_id % np.uint32(plan.world) == plan.rank)
    cols = {k: getattr(tab, k)[keep] for k in _PAF_COLS}
    cols["read_id"] = cols["read_id"] // np.uint32(plan.world)
    return PafTable(*[cols[k] for k in _PAF_COLS])


def configure(ctx, plan: ShardPlan, lengths, name_rank, max_reads, max_bam_files=2):
    """contig table, owners and this rank's exchange area -> its CUDA IPC handle (uint8[64]).  The caller gathers
    every rank's handle and calls ctx.shard_open (processes) or ctx.shard_attach (contexts of one process)."""
    ctx.set_contigs(lengths, plan.owned)
    ctx.set_name_rank(name_rank)
    ctx.shard_config(plan.rank, plan.world, plan.owner, plan.selected)
    return ctx.shard_alloc(max_reads, max_bam_files)


def open_over_process_group(ctx, plan: ShardPlan, handle):
    """all-gather the IPC handles over torch.distributed and map the peers' areas"""
    from . import dist as D
    handles = np.zeros((plan.world, 64), np.int64)
    handles[plan.rank] = handle
    ctx.shard_open(D.allreduce(handles, "sum").astype(np.uint8))
