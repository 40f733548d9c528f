"""CPU, world_size 2, 3, 4 and 8 over gloo: the patch sharding plan and the slab exchange reproduce the single-process
Gaussian-weighted accumulation (the arithmetic of the exchange; on GPUs the same code runs over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _patch_contribution(i, C, patch):
    rng = np.random.default_rng(1000 + i)
    return rng.integers(-8, 9, size=(C, *patch)).astype(np.float32)  # small integers: fp32 sums are order independent


def _worker(rank, world, port, shape, patch, step, C, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from boa_b200.dist import DistContext, exchange_slabs, gather_label_slabs, plan_shards
    from boa_b200.geometry import compute_gaussian, sliding_window_origins
    origins = sliding_window_origins(shape, patch, step)
    g = np.rint(compute_gaussian(patch).astype(np.float32) * 4)  # integer weights, exact sums
    plan = plan_shards(origins, patch[0], shape[0], world, rank)
    acc = np.zeros((C, plan.zhi - plan.zlo, shape[1], shape[2]), dtype=np.float32)
    for i in range(plan.begin, plan.end):
        o = origins[i]
        acc[:, o[0] - plan.zlo:o[0] - plan.zlo + patch[0], o[1]:o[1] + patch[1], o[2]:o[2] + patch[2]] += \
            _patch_contribution(i, C, patch) * g
    ctx = DistContext(rank, world, None)
    slab = exchange_slabs(torch.from_numpy(acc), plan, ctx)
    lab = slab.argmax(0).to(torch.uint8)
    full = gather_label_slabs(lab, plan, ctx)
    np.save(os.path.join(out_dir, f"slab_{rank}.npy"), slab.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "labels.npy"), full.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,patch,step", [(2, (40, 24, 20), (16, 16, 16), 0.5),
                                                    (3, (37, 16, 33), (16, 16, 16), 0.8),
                                                    (4, (16, 16, 40), (16, 16, 16), 0.5),
                                                    (8, (70, 20, 36), (16, 16, 16), 0.5),   # 64 patches on 8 ranks
                                                    (8, (16, 16, 24), (16, 16, 16), 0.5)])  # 2 patches: 6 idle ranks
def test_slab_exchange_matches_single_process(tmp_path, world, shape, patch, step):
    from boa_b200.geometry import compute_gaussian, shard_patches, sliding_window_origins
    C = 3
    mp.spawn(_worker, args=(world, _free_port(), shape, patch, step, C, str(tmp_path)), nprocs=world, join=True)
    origins = sliding_window_origins(shape, patch, step)
    g = np.rint(compute_gaussian(patch).astype(np.float32) * 4)
    ref = np.zeros((C, *shape), dtype=np.float32)
    for i, o in enumerate(origins):
        ref[:, o[0]:o[0] + patch[0], o[1]:o[1] + patch[1], o[2]:o[2] + patch[2]] += _patch_contribution(i, C, patch) * g
    for r in range(world):
        lo, hi = shard_patches(shape[0], world, r)
        assert np.array_equal(np.load(tmp_path / f"slab_{r}.npy"), ref[:, lo:hi])
    assert np.array_equal(np.load(tmp_path / "labels.npy"), ref.argmax(0).astype(np.uint8))


def test_shard_plan_is_balanced_and_complete():
    from boa_b200.dist import plan_shards
    from boa_b200.geometry import shard_patches, sliding_window_origins
    origins = sliding_window_origins((512, 512, 512), (128,) * 3, 0.8)
    for world in (1, 2, 4, 8):
        runs = [shard_patches(len(origins), world, r) for r in range(world)]
        assert runs[0][0] == 0 and runs[-1][1] == len(origins)
        assert all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
        sizes = [e - b for b, e in runs]
        assert max(sizes) - min(sizes) <= 1
        for r in range(world):
            p = plan_shards(origins, 128, 512, world, r)
            z = origins[p.begin:p.end, 0]
            assert p.zlo == z.min() and p.zhi == z.max() + 128
            if world >= 4:
                assert p.zhi - p.zlo <= 224  # at most two patch layers
