"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch; gloo on CPU
for tests).  Contigs shard across ranks; the only exchanges of the path are

  * the read-set exchange of a sharded run (winners to the read homes, survivors to the contig owners): inside
    libgci_cuda.so over NVLink peer memory (csrc/shard.cu, host side gci_b200/sharded.py) — not here;
  * genome_row(): all-reduce of (sum depth, sum length, curated-contig count) and all-gather of the
    curated lengths for the genome-level N50 / GCI row (GCI.py:572-587, :862-868).

"""
from __future__ import annotations

import os

import numpy as np


def is_dist():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def init(backend=None):
    """Initialise from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world, local)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def rank():
    import torch.distributed as dist
    return dist.get_rank() if is_dist() else 0


def world():
    import torch.distributed as dist
    return dist.get_world_size() if is_dist() else 1


def barrier():
    if is_dist():
        import torch.distributed as dist
        dist.barrier()


def gather_objects(obj):
    """every rank's (small, picklable) object, in rank order, on every rank"""
    if not is_dist():
        return [obj]
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _device():
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce(values, op="sum"):
    """numpy int64/float64 vector -> reduced over ranks (identity without a process group)."""
    values = np.asarray(values)
    if not is_dist():
        return values.copy()
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(values)).to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)
    return t.cpu().numpy()


def allgather_varlen(arr):
    """Each rank contributes a 1-D array of its own length; returns the list of all ranks' arrays."""
    arr = np.ascontiguousarray(arr)
    if not is_dist():
        return [arr.copy()]
    orig = arr.dtype
    if orig.kind == "u" or orig == np.bool_:      # unsigned types do not travel through every backend
        arr = arr.astype(np.int64)
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = _device()
    n = torch.tensor([arr.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(1, max(sizes))
    buf = torch.zeros(cap, dtype=torch.from_numpy(arr[:0]).dtype, device=dev)
    if arr.size:
        buf[:arr.size] = torch.from_numpy(arr).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o[:s].cpu().numpy().astype(orig) for o, s in zip(outs, sizes)]


def init_native_comm(ctx, p2p=True, cap=2048):
    """Create the library's own NCCL communicator (Context.genome_row): rank 0 makes the id, everybody gets it."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = ctx.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)
    uid = allreduce(uid.astype(np.int64), "sum").astype(np.uint8)
    ctx.comm_init(uid, rank, world)
    ctx.row_exchange = "ncclAllGather"
    import os
    if os.environ.get("GCI_P2P", "1").startswith("0"):
        p2p = False
    if world > 1 and p2p:
        # genome row over NVLink peer memory: all-gather the CUDA IPC handles of the receive areas, map the peers.
        # Every rank must end up on the same path, so the outcome is agreed on before anybody uses it.
        handles = np.zeros((world, 64), np.int64)
        ok = 1
        try:
            handles[rank] = ctx.comm_p2p_alloc(cap)
        except Exception:
            ok = 0
        handles = allreduce(handles, "sum").astype(np.uint8)
        if int(allreduce(np.array([1 - ok], np.int64), "max")[0]) == 0:
            try:
                ctx.comm_p2p_open(handles)
            except Exception:
                ok = 0
        if int(allreduce(np.array([1 - ok], np.int64), "max")[0]) != 0:
            ctx.comm_p2p_disable()
        else:
            ctx.row_exchange = "stores into NVLink peer memory (CUDA IPC), one kernel"


def assign_contigs(lengths, weights, world):
    """LPT bin packing of contigs onto ranks by (aligned bases, length).  Returns owner[contig]."""
    order = sorted(range(len(lengths)), key=lambda i: (-(weights[i] if weights is not None else 0), -lengths[i], i))
    load = [0.0] * world
    owner = [0] * len(lengths)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += (weights[i] if weights is not None else 0) + lengths[i]
    return owner


_ROW_CAP = 2048
_row_cache = {}


def genome_row(sum_depth, sum_len, n_ctg, lengths):
    """Whole-genome terms from per-rank partials: (mean depth, total curated contigs, all curated lengths).

    One fixed-size all-gather ([sum_depth, sum_len, n_ctg, n_lengths, lengths...] per rank) when the
    curated-length list fits _ROW_CAP entries — the exchange is latency-bound, so it is a single
    collective and a single device->host copy; longer lists take the variable-length path."""
    lengths = np.asarray(lengths, dtype=np.int64)
    if not is_dist():
        mean = float(sum_depth) / float(sum_len) if sum_len else float("nan")
        return mean, int(n_ctg), lengths.copy()
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = _device()
    fits = np.array([1 if len(lengths) <= _ROW_CAP else 0], dtype=np.int64)
    if len(lengths) > _ROW_CAP or not _row_cache.get("all_fit", True):
        if int(allreduce(fits, "sum")[0]) != world:        # somebody overflowed: variable-length path
            _row_cache["all_fit"] = False
            tot = allreduce(np.array([int(sum_depth), int(sum_len), int(n_ctg)], dtype=np.int64))
            all_len = np.concatenate(allgather_varlen(lengths))
            mean = float(tot[0]) / float(tot[1]) if tot[1] else float("nan")
            return mean, int(tot[2]), all_len
    key = (world, str(dev))
    if key not in _row_cache:
        pin = dev.type == "cuda"
        _row_cache[key] = (torch.zeros(4 + _ROW_CAP, dtype=torch.int64, pin_memory=pin),
                           torch.zeros(4 + _ROW_CAP, dtype=torch.int64, device=dev),
                           torch.zeros(world * (4 + _ROW_CAP), dtype=torch.int64, device=dev),
                           torch.zeros(world * (4 + _ROW_CAP), dtype=torch.int64, pin_memory=pin))
    h_in, d_in, d_out, h_out = _row_cache[key]
    v = h_in.numpy()
    v[0], v[1], v[2], v[3] = int(sum_depth), int(sum_len), int(n_ctg), len(lengths)
    v[4:4 + len(lengths)] = lengths
    d_in.copy_(h_in, non_blocking=True)
    dist.all_gather_into_tensor(d_out, d_in)
    h_out.copy_(d_out, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    o = h_out.numpy().reshape(world, 4 + _ROW_CAP)
    head = o[:, :4].sum(axis=0)
    tot_d, tot_l, tot_c = int(head[0]), int(head[1]), int(head[2])
    body = o[:, 4:]
    all_len = body[np.arange(_ROW_CAP)[None, :] < o[:, 3:4]]
    mean = float(tot_d) / float(tot_l) if tot_l else float("nan")
    return mean, tot_c, all_len
