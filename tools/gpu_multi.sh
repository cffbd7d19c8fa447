#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded parity test, then the bench at N ranks. Usage: bash tools/gpu_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_shard_gpu.py -q -x --timeout 600 > gpurun_out/pytest_shard_gpu.log 2>&1; echo "pytest shard rc=$?"; tail -30 gpurun_out/pytest_shard_gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"; tail -c 5000 gpurun_out/bench_${N}gpu.json; tail -15 gpurun_out/bench_${N}gpu.err | cut -c1-300
