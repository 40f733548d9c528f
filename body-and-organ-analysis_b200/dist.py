"""Multi-GPU sharding of one volume: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch) for the plumbing.

Work:      the slicer list (dim-0-major order, predict_from_raw_data.py:532-537) is cut into `world` contiguous,
           count-balanced runs; a rank accumulates its patches into a PRIVATE fp32 buffer that covers only the dim-0
           extent its run touches.
Ownership: the output volume is cut into `world` contiguous dim-0 slabs; rank r finalises slab r.
Exchange:  once per model, every rank sends each other owner the intersection of its touched extent with that
           owner's slab (grouped NCCL send / recv); owners add the pieces IN RANK ORDER (deterministic for a given
           world size) with boa_add_slab, run normalise + argmax on their slab, and the uint8 label slabs are
           all-gathered.  The weight sum `n` is input independent and never exchanged.
The reference has no multi-GPU inference path; nothing here replaces reference code.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .geometry import shard_patches


@dataclass
class DistContext:
    rank: int = 0
    world_size: int = 1
    group: object = None


@dataclass
class ShardPlan:
    begin: int            # my patches [begin, end) of the slicer list
    end: int
    zlo: int              # dim-0 extent my patches touch (zlo == zhi when I have no patch)
    zhi: int
    slabs: list           # [(lo, hi)] owned dim-0 slab of every rank
    touched: list         # [(zlo, zhi)] touched extent of every rank


def plan_shards(origins: np.ndarray, patch0: int, Z: int, world: int, rank: int) -> ShardPlan:
    slabs = [shard_patches(Z, world, r) for r in range(world)]
    touched = []
    for r in range(world):
        b, e = shard_patches(len(origins), world, r)
        if e > b:
            z = origins[b:e, 0]
            touched.append((int(z.min()), int(z.max()) + patch0))
        else:
            touched.append((0, 0))
    b, e = shard_patches(len(origins), world, rank)
    return ShardPlan(b, e, touched[rank][0], touched[rank][1], slabs, touched)


def _inter(a, b):
    lo, hi = max(a[0], b[0]), min(a[1], b[1])
    return (lo, hi) if hi > lo else None


def _add(dst: torch.Tensor, src: torch.Tensor) -> None:
    if dst.is_cuda:
        from . import _lib
        with torch.cuda.device(dst.device):
            _lib.check(_lib.lib().boa_add_slab(_lib.ptr(dst), _lib.ptr(src), dst.numel(), _lib.stream_ptr()))
    else:  # gloo / CPU tests of the exchange logic only
        dst.add_(src)


def exchange_slabs(acc_local: torch.Tensor, plan: ShardPlan, ctx: DistContext) -> torch.Tensor:
    """acc_local fp32 [C, zhi - zlo, Y, X] (my partial sums) -> fp32 [C, slab_len, Y, X]: the complete sums of my slab."""
    import torch.distributed as dist

    C, _, Y, X = acc_local.shape
    me, world = ctx.rank, ctx.world_size
    my_slab = plan.slabs[me]
    sends, recvs, ops = [], [], []
    for o in range(world):  # what I send
        if o == me:
            continue
        it = _inter(plan.touched[me], plan.slabs[o])
        if it is not None:
            buf = acc_local[:, it[0] - plan.zlo:it[1] - plan.zlo].contiguous()
            sends.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, o if ctx.group is None else dist.get_global_rank(ctx.group, o),
                                  ctx.group))
    for s in range(world):  # what I receive
        if s == me:
            continue
        it = _inter(plan.touched[s], my_slab)
        if it is not None:
            buf = torch.empty((C, it[1] - it[0], Y, X), dtype=torch.float32, device=acc_local.device)
            recvs.append((s, it, buf))
            ops.append(dist.P2POp(dist.irecv, buf, s if ctx.group is None else dist.get_global_rank(ctx.group, s),
                                  ctx.group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    slab = torch.zeros((C, my_slab[1] - my_slab[0], Y, X), dtype=torch.float32, device=acc_local.device)
    pieces = {s: (it, buf) for s, it, buf in recvs}
    own = _inter(plan.touched[me], my_slab)
    if own is not None:
        pieces[me] = (own, acc_local[:, own[0] - plan.zlo:own[1] - plan.zlo])
    for s in sorted(pieces):  # fixed rank order => deterministic fp32 sums
        it, buf = pieces[s]
        a, b = it[0] - my_slab[0], it[1] - my_slab[0]
        for c in range(C):  # [len, Y, X] blocks are contiguous inside both tensors
            _add(slab[c, a:b], buf[c].contiguous() if not buf[c].is_contiguous() else buf[c])
    return slab


def gather_label_slabs(lab_slab: torch.Tensor, plan: ShardPlan, ctx: DistContext) -> torch.Tensor:
    """uint8 [slab_len, Y, X] of every rank -> uint8 [Z, Y, X] on every rank."""
    import torch.distributed as dist

    max_len = max(hi - lo for lo, hi in plan.slabs)
    Y, X = lab_slab.shape[1:]
    padded = torch.zeros((max_len, Y, X), dtype=torch.uint8, device=lab_slab.device)
    padded[:lab_slab.shape[0]] = lab_slab
    flat = torch.empty((ctx.world_size * max_len, Y, X), dtype=torch.uint8, device=lab_slab.device)
    dist.all_gather_into_tensor(flat, padded, group=ctx.group)
    out = flat.view(ctx.world_size, max_len, Y, X)
    return torch.cat([out[r, :hi - lo] for r, (lo, hi) in enumerate(plan.slabs)], dim=0)
