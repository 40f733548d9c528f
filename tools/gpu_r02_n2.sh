#!/bin/bash
# 2 GPUs: peer-memory exchange test + bench at N=2 (peer path and NCCL path)
set -u
TAG=${1:-r02n2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -s > $OUT/${TAG}_pytest_dist.log 2>&1
echo "dist test exit $?"; tail -12 $OUT/${TAG}_pytest_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --quick > $OUT/${TAG}_bench_peer.json 2> $OUT/${TAG}_bench_peer.err
echo "bench peer exit $?"; tail -c 2500 $OUT/${TAG}_bench_peer.json; tail -5 $OUT/${TAG}_bench_peer.err
BOA_B200_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 2 --quick > $OUT/${TAG}_bench_nccl.json 2> $OUT/${TAG}_bench_nccl.err
echo "bench nccl exit $?"; tail -c 1800 $OUT/${TAG}_bench_nccl.json; tail -3 $OUT/${TAG}_bench_nccl.err
