#!/bin/bash
# 1/2/4/8-GPU strong-scaling sweep of bench.py on one node (run under `gpurun --gpus 8`).
STEPS=${1:-50}
GPUS=${2:-"1 2 4 8"}
mkdir -p gpurun_out
for N in $GPUS; do
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 2 --skip-cpu > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) \
      bench.py --gpus $N --steps $STEPS --warmup 2 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  echo "N=$N rc=$?"
  grep -h '^{' gpurun_out/scale_$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'steps/s', round(d['value'],3), 'sec/video', round(d['sec_per_video'],2), 'e2e', round(d['e2e']['value'],3), 'attn TF/s', round(d['roofline']['achieved']), d['clocks'], 'graphs', d['config'].get('cuda_graphs'), 'launches', d['gpu_launches'])" || tail -5 gpurun_out/scale_$N.err
done
