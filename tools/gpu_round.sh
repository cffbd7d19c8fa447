#!/bin/bash
# One GPU-box session: GPU test suite, head timing + ncu, bench line. Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/one_head.py 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']); print(d['hit_path'])
print({k:round(v['ms_avg'],4) for k,v in d['kernels'].items()})
PY
tail -5 gpurun_out/bench_1gpu.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_tc -s 2 -c 1 -o gpurun_out/r02_head_hit python tools/one_head.py > gpurun_out/ncu_head.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_head.log
