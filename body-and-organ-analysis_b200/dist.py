"""Multi-GPU sharding of one volume: one process per GPU, `torch.distributed` (NCCL over
NVLink / NVSwitch) for the plumbing.

Work:      the slicer list (dim-0-major order, predict_from_raw_data.py:532-537) is cut into `world` contiguous,
           count-balanced runs; a rank accumulates its patches into a PRIVATE fp32 buffer that covers only the dim-0
           extent its run touches.
Ownership: the output volume is cut into `world` contiguous dim-0 slabs; rank r finalises slab r.
Exchange:  once per model.
           Peer-memory path (PeerExchange, the default on one box): the private buffers live in CUDA-IPC memory that
           every rank maps; after a stream-ordered barrier the owner of a slab runs ONE kernel that reads every rank's
           part of its slab over NVLink, adds the parts IN RANK ORDER, divides, takes the argmax and merges the part
           labels (boa_reduce_finalize_peers) - the summed logits are never stored, nothing is staged or copied; the
           uint8 label slabs are all-gathered once per task.
           NCCL path (exchange_slabs; gloo on the CPU for the tests, or no peer access): grouped send / recv of the
           pieces, the owner adds them in rank order (one strided launch per piece), finalises its slab and the
           label slabs are all-gathered per model.
           Both give the same bits (same additions in the same order).  The weight sum `n` is input independent and
           never exchanged.
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
    _peers: object = None   # PeerExchange, created on first use
    _peers_failed: bool = False

    def peer_exchange(self, device):
        """The PeerExchange of this context, or None when the peer-memory path is unavailable (no CUDA, no peer access
        between the GPUs, BOA_B200_EXCHANGE=nccl).  The decision is collective: every rank gets the same answer."""
        import os

        import torch.distributed as dist
        if self.world_size == 1 or self._peers_failed or os.environ.get("BOA_B200_EXCHANGE", "peer") == "nccl":
            return None
        if self._peers is None:
            ok = torch.ones(1, dtype=torch.int32, device=device)
            try:
                if self.world_size > MAX_PEER_RANKS:
                    raise RuntimeError(f"more than {MAX_PEER_RANKS} ranks")
                ex = PeerExchange(self, device)
            except Exception as e:  # noqa: BLE001 - any failure on any rank switches EVERY rank to the NCCL path
                ex, ok[0] = None, 0
                import logging
                logging.getLogger(__name__).warning("peer-memory exchange unavailable (%s): using NCCL send/recv", e)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if ex is not None:
                    ex.release(collective=False)
                self._peers_failed = True
                return None
            self._peers = ex
        return self._peers


MAX_PEER_RANKS = 8  # boa_reduce_finalize_peers (passes.cu MAX_PEERS)


class PeerExchange:
    """Two private accumulation buffers per rank (models alternate between them) in CUDA-IPC memory, and the mapped
    pointers of every peer's buffers.  Protocol per model, everything stream ordered on the caller's stream:
        zero my buffer b -> my patches accumulate into it -> barrier (every rank's patches are done)
        -> boa_reduce_finalize_peers on my slab (reads the peers' buffers b)
    Buffer b is cleared again two models later: that clear comes after the NEXT model's barrier in this rank's stream,
    which completes only when every rank has reached it, i.e. has finished its reduce of this model - so one barrier
    per model is enough."""

    GRANULE = 256 << 20

    def __init__(self, ctx: DistContext, device):
        from . import _lib
        self._lib = _lib
        self.ctx, self.device = ctx, torch.device(device)
        self.capacity = 0
        self.local = [None, None]
        self.peers = [[None] * ctx.world_size, [None] * ctx.world_size]
        self.turn = 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
        # probe: a small allocation must round-trip through IPC on every rank before the path is trusted
        self.ensure(1 << 20)

    def ensure(self, nbytes: int) -> None:
        """Collective (every rank passes the same nbytes): grow both buffers to at least nbytes."""
        import ctypes as C

        import torch.distributed as dist
        if nbytes <= self.capacity:
            return
        L, lib = self._lib, self._lib.lib()
        self.release(collective=True)
        cap = (nbytes + self.GRANULE - 1) // self.GRANULE * self.GRANULE
        handles = []
        with torch.cuda.device(self.device):
            for b in range(2):
                ptr = C.c_void_p()
                h = (C.c_ubyte * 64)()
                L.check(lib.boa_comm_alloc(cap, C.byref(ptr), h))
                self.local[b] = ptr.value
                handles.append(bytes(h))
            gathered = [None] * self.ctx.world_size
            dist.all_gather_object(gathered, handles, group=self.ctx.group)
            for r, hs in enumerate(gathered):
                for b in range(2):
                    if r == self.ctx.rank:
                        self.peers[b][r] = self.local[b]
                    else:
                        ptr = C.c_void_p()
                        h = (C.c_ubyte * 64).from_buffer_copy(hs[b])
                        L.check(lib.boa_comm_open(h, C.byref(ptr)))
                        self.peers[b][r] = ptr.value
        self.capacity = cap
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.ctx.group)

    def release(self, collective: bool = True) -> None:
        import torch.distributed as dist
        if self.capacity == 0 and self.local[0] is None:
            return
        L, lib = self._lib, self._lib.lib()
        torch.cuda.synchronize(self.device)
        if collective:
            dist.barrier(group=self.ctx.group)  # nobody still reads my buffers
        with torch.cuda.device(self.device):
            for b in range(2):
                for r, p in enumerate(self.peers[b]):
                    if p is not None and r != self.ctx.rank:
                        L.check(lib.boa_comm_close(p))
                    self.peers[b][r] = None
            if collective:
                dist.barrier(group=self.ctx.group)  # every mapping of my buffers is closed
            for b in range(2):
                if self.local[b] is not None:
                    L.check(lib.boa_comm_free(self.local[b]))
                    self.local[b] = None
        self.capacity = 0

    def next_buffer(self) -> int:
        self.turn ^= 1
        return self.turn

    def zero(self, b: int, nbytes: int) -> None:
        L = self._lib
        with torch.cuda.device(self.device):
            L.check(L.lib().boa_comm_zero(self.local[b], nbytes, L.stream_ptr()))

    def barrier(self) -> None:
        """Stream-ordered: the caller's stream continues once every rank's stream has reached this point."""
        import torch.distributed as dist
        dist.all_reduce(self._flag, group=self.ctx.group)

    def reduce_finalize(self, b: int, plan: "ShardPlan", Y: int, X: int, w_slab: torch.Tensor, C_: int, lut,
                        overwrite_nonzero_only: bool, lab_slab: torch.Tensor, bad: torch.Tensor) -> None:
        import ctypes as C
        L = self._lib
        W = self.ctx.world_size
        bases = (C.c_void_p * W)(*[self.peers[b][r] for r in range(W)])
        zlo = (C.c_int32 * W)(*[t[0] for t in plan.touched])
        zhi = (C.c_int32 * W)(*[t[1] for t in plan.touched])
        lo, hi = plan.slabs[self.ctx.rank]
        lut_arr = (C.c_uint8 * C_)(*(range(C_) if lut is None else [int(v) for v in lut]))
        with torch.cuda.device(self.device):
            L.check(L.lib().boa_reduce_finalize_peers(bases, zlo, zhi, W, lo, hi, Y, X, L.ptr(w_slab), C_, lut_arr,
                                                      int(overwrite_nonzero_only), L.ptr(lab_slab), L.ptr(bad),
                                                      L.stream_ptr()))


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


def _add_piece(slab: torch.Tensor, a: int, b: int, buf: torch.Tensor) -> None:
    """slab[:, a:b] += buf  ([C, b - a, Y, X] pieces; buf may itself be a dim-1 slice of a longer buffer)."""
    if slab.is_cuda:
        from . import _lib
        C_, _, Y, X = slab.shape
        n = (b - a) * Y * X
        with torch.cuda.device(slab.device):
            _lib.check(_lib.lib().boa_add_slab_strided(_lib.ptr(slab[:, a:b]), slab.stride(0), _lib.ptr(buf),
                                                       buf.stride(0), C_, n, _lib.stream_ptr()))
    else:  # gloo / CPU tests of the exchange logic only
        slab[:, a:b].add_(buf)


def exchange_slabs(acc_local: torch.Tensor, plan: ShardPlan, ctx: DistContext) -> torch.Tensor:
    """NCCL / gloo path.  acc_local fp32 [C, zhi - zlo, Y, X] (my partial sums) -> fp32 [C, slab_len, Y, X]: the
    complete sums of my slab.  Pieces are sent straight from the private buffer, one contiguous [len, Y, X] block per
    channel (no staging copy), all in one grouped call."""
    import torch.distributed as dist

    C, _, Y, X = acc_local.shape
    me, world = ctx.rank, ctx.world_size
    my_slab = plan.slabs[me]
    peer = lambda r: r if ctx.group is None else dist.get_global_rank(ctx.group, r)
    recvs, ops = [], []
    for o in range(world):  # what I send
        if o == me:
            continue
        it = _inter(plan.touched[me], plan.slabs[o])
        if it is not None:
            for c in range(C):
                ops.append(dist.P2POp(dist.isend, acc_local[c, it[0] - plan.zlo:it[1] - plan.zlo], peer(o), ctx.group))
    for s in range(world):  # what I receive
        if s == me:
            continue
        it = _inter(plan.touched[s], my_slab)
        if it is not None:
            buf = torch.empty((C, it[1] - it[0], Y, X), dtype=torch.float32, device=acc_local.device)
            recvs.append((s, it, buf))
            for c in range(C):
                ops.append(dist.P2POp(dist.irecv, buf[c], peer(s), ctx.group))
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
        _add_piece(slab, it[0] - my_slab[0], it[1] - my_slab[0], buf)
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
