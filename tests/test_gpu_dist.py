"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): sharded prediction equals the single-GPU prediction."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from boa_b200 import zoo
    from boa_b200.dist import DistContext
    from boa_b200.labels import part_luts
    from boa_b200.pipeline import ModelZoo, segment_task
    specs = zoo.synthetic_specs((32, 32, 32), 32, 64, 3, bca_folds=1, seed=1, datasets=[291, 292])
    mz = ModelZoo.from_specs(specs, device=torch.device("cuda", rank), max_batch=2)
    ct = torch.from_numpy(zoo.synthetic_ct((72, 48, 40), seed=2)).cuda()
    lab = segment_task(ct, mz, [291, 292], [0], 0.8, part_luts()[:2], DistContext(rank, world, None))
    if rank == 0:
        single = segment_task(ct, mz, [291, 292], [0], 0.8, part_luts()[:2], None)
        np.save(os.path.join(out_dir, "sharded.npy"), lab.cpu().numpy())
        np.save(os.path.join(out_dir, "single.npy"), single.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_matches_single_gpu(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "sharded.npy"), np.load(tmp_path / "single.npy")
    # the owner adds the two partial sums in rank order, the single GPU adds patches in slicer order: fp32 sums may
    # differ in the last bit, argmax only flips on exact near-ties
    assert (a == b).mean() > 0.9999
