#!/bin/bash
# Final check of a build on one B200: the GPU test suite, smoke(), the driver's bench invocation.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['hit_path']['frac'], d['hit_path'].get('queued'), d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench_1gpu.err
