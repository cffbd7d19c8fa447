#!/bin/bash
# On an 8-GPU box: the bench at the given rank counts back to back. Usage: bash tools/gpu_scale.sh 4 8
set -u
mkdir -p gpurun_out
for N in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('n_gpus','value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms','shard_parity')}, d['e2e']['value'], d['clocks'])
    print({k:round(v['ms_avg'],4) for k,v in d['kernels'].items()})
except Exception as ex:
    print('parse failed', ex)
PY
  tail -5 gpurun_out/bench_${N}gpu.err | cut -c1-300
done
