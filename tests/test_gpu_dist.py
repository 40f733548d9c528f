"""GPU, 2 ranks (skipped on a single-GPU box): the sharded prediction - over the peer-memory path (CUDA-IPC buffers,
fused reduce + finalize kernel reading the peers over NVLink) and over the NCCL send / recv path - equals the single-GPU
prediction, the two paths are bit-identical to each other, and the sharded post-processing / full pipeline agree."""
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
    from boa_b200.pipeline import ModelZoo, analyze_volume, segment_task
    specs = zoo.synthetic_specs((32, 32, 32), 32, 64, 3, bca_folds=1, seed=1, datasets=[291, 292, 293, 294, 295, 542, 543])
    mz = ModelZoo.from_specs(specs, device=torch.device("cuda", rank), max_batch=2)
    # odd in-plane size: slab offsets that are not 16-byte aligned take the scalar kernels
    for name, shape in (("even", (72, 48, 40)), ("odd", (70, 45, 39))):
        ct = torch.from_numpy(zoo.synthetic_ct(shape, seed=2)).cuda()
        ctx = DistContext(rank, world, None)
        peers = segment_task(ct, mz, [291, 292], [0], 0.8, part_luts()[:2], ctx)
        used_peers = ctx._peers is not None
        os.environ["BOA_B200_EXCHANGE"] = "nccl"
        nccl = segment_task(ct, mz, [291, 292], [0], 0.8, part_luts()[:2], DistContext(rank, world, None))
        del os.environ["BOA_B200_EXCHANGE"]
        if rank == 0:
            single = segment_task(ct, mz, [291, 292], [0], 0.8, part_luts()[:2], None)
            np.savez(os.path.join(out_dir, f"seg_{name}.npz"), peers=peers.cpu().numpy(), nccl=nccl.cpu().numpy(),
                     single=single.cpu().numpy(), used_peers=used_peers)
    # whole pipeline, sharded vs single: label maps, post-processing (labels dealt out to the ranks), measurements
    ct = torch.from_numpy(zoo.synthetic_ct((96, 64, 64), seed=5)).cuda()
    res = analyze_volume(ct, (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True,
                         dist_ctx=DistContext(rank, world, None))
    # the concurrent post-processing of the two body-composition maps (default from 4 ranks) on 2 ranks
    os.environ["BOA_B200_PAIR_MIN_RANKS"] = "2"
    res2 = analyze_volume(ct, (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True,
                          dist_ctx=DistContext(rank, world, None))
    del os.environ["BOA_B200_PAIR_MIN_RANKS"]
    same_pair = all(torch.equal(getattr(res, k), getattr(res2, k)) for k in ("body_parts", "body_regions", "tissues"))
    # host API on N ranks: one caller (rank 0) receives the label maps, every rank the measurement dicts
    from boa_b200.pipeline import analyze_from_host
    host = analyze_from_host(ct.cpu().pin_memory(), (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True,
                             dist_ctx=DistContext(rank, world, None))
    host_ok = host["total_measurements"] == res.total_measurements and (
        all(torch.equal(host[k], getattr(res, k).cpu()) for k in ("total", "body_parts", "body_regions", "tissues"))
        if rank == 0 else all(host[k] is None for k in ("total", "body_parts", "body_regions", "tissues")))
    flag = torch.tensor([int(host_ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ref = analyze_volume(ct, (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True)
        import json
        np.savez(os.path.join(out_dir, "pipeline.npz"),
                 **{f"{k}_{w}": getattr(r, k).cpu().numpy() for w, r in (("dist", res), ("single", ref))
                    for k in ("total", "body_parts", "body_regions", "tissues")})
        with open(os.path.join(out_dir, "meas.json"), "w") as f:
            json.dump({"dist": [res.total_measurements, res.bca_measurements],
                       "single": [ref.total_measurements, ref.bca_measurements], "same_pair": bool(same_pair),
                       "host_api_ok": bool(flag.item()),
                       "total_equal": bool(torch.equal(res.total, ref.total))}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_single_gpu(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json

    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for name in ("even", "odd"):
        z = np.load(tmp_path / f"seg_{name}.npz")
        assert bool(z["used_peers"]), "the peer-memory path was not taken on a 2-GPU box"
        # same additions in the same (rank) order on both paths
        assert np.array_equal(z["peers"], z["nccl"]), name
        # the owner adds the two partial sums in rank order, the single GPU adds patches in slicer order: fp32 sums may
        # differ in the last bit, argmax only flips on exact near-ties
        assert (z["peers"] == z["single"]).mean() > 0.9999, name
    p = np.load(tmp_path / "pipeline.npz")
    for k in ("total", "body_parts", "body_regions", "tissues"):
        agree = (p[f"{k}_dist"] == p[f"{k}_single"]).mean()
        print(f"pipeline {k}: sharded == single on {agree:.6f} of the voxels")
        assert agree > 0.999, k
    m = json.load(open(tmp_path / "meas.json"))
    # the measurement dicts are functions of the label maps: same keys, and equal wherever the maps are equal
    assert m["dist"][0]["segmentations"]["total"].keys() == m["single"][0]["segmentations"]["total"].keys()
    assert m["dist"][1].keys() == m["single"][1].keys()
    assert m["host_api_ok"], "analyze_from_host on 2 ranks: maps on rank 0 only, measurements everywhere"
    assert m["same_pair"], "concurrent post-processing of body_parts / body_regions differs from the sequential one"
    if m["total_equal"]:  # sharded histograms, all-reduced: exact - same label map => same statistics
        from test_oracle_golden import _close
        _close(m["single"][0], m["dist"][0])
