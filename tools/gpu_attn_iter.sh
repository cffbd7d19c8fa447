#!/bin/bash
# Attention iteration on one GPU: attention parity tests (optionally under an env override), isolated A/B timing, then the bench
# line. Usage: bash tools/gpu_attn_iter.sh [bench|nobench] [VAR=value ...]   Logs under gpurun_out/.
set -u
mode=${1:-bench}; shift || true
mkdir -p gpurun_out
env "$@" timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullshape_parity_gpu.py -m gpu -q -x --timeout 600 -k "attention or attn or full_shape" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_attn.log | cut -c1-250
timeout 300 python tools/attn_bench.py 32760 2 > gpurun_out/attn_bench.log 2>&1; echo "attn_bench rc=$?"; tail -22 gpurun_out/attn_bench.log
if [ "$mode" = "bench" ]; then
env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['kernels']['attn_self']['ms_avg'], d['hit_path']['frac'])
PY
tail -3 gpurun_out/bench_1gpu.err
fi
