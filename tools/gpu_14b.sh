#!/bin/bash
# 8-GPU box: BASELINE configs[4] — Wan2.1-T2V-14B 1280x720x81f sharded over 8 GPUs.
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --workload wan14b --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_wan14b_${N}gpu.json 2> gpurun_out/bench_wan14b_${N}gpu.err; echo "bench 14B N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_wan14b_${N}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('n_gpus','value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms','shard_parity')}, d['e2e']['value'], d['clocks'])
    print(d['roofline'])
    print({k:round(v['ms_avg'],4) for k,v in d['kernels'].items()})
except Exception as ex:
    print('parse failed', ex)
PY
tail -8 gpurun_out/bench_wan14b_${N}gpu.err | cut -c1-400
