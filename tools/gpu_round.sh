#!/bin/bash
# One GPU-box session: GPU test suite, MMDiT full-size bench lines, Wan bench line. Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "ours vs|passed|failed|rel-L2" gpurun_out/pytest_gpu.log | tail -30 | cut -c1-250
timeout 600 python bench.py --workload flux --steps 28 --warmup 3 > gpurun_out/bench_flux.json 2> gpurun_out/bench_flux.err; echo "flux rc=$?"; tail -c 1800 gpurun_out/bench_flux.json; tail -3 gpurun_out/bench_flux.err
timeout 900 python bench.py --workload hunyuan720p --steps 12 --warmup 3 > gpurun_out/bench_hunyuan.json 2> gpurun_out/bench_hunyuan.err; echo "hunyuan rc=$?"; tail -c 1800 gpurun_out/bench_hunyuan.json; tail -3 gpurun_out/bench_hunyuan.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['frac'], d['hit_path'])
print({k:round(v,4) for k,v in d['attribution']['share_of_forward_time'].items()})
PY
tail -5 gpurun_out/bench_1gpu.err
