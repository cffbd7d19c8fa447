#!/bin/bash
# One GPU-box session: GPU test suite, attention A/B, bench line (two emulation fractions). Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/attn_bench.py 32760 2 > gpurun_out/attn_bench.log 2>&1; echo "attn_bench rc=$?"; tail -12 gpurun_out/attn_bench.log
for EMU in 2 3; do
MC_ATTN_EMU=$EMU timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_1gpu_emu$EMU.json 2> gpurun_out/bench_1gpu.err; echo "bench emu=$EMU rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu_emu$EMU.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['kernels']['attn_self']['ms_avg'], d['hit_path']['frac'])
PY
done
tail -5 gpurun_out/bench_1gpu.err
MC_ATTN_EMU=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_long -s 1 -c 1 -o gpurun_out/r02_attn_long_spec python tools/one_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_attn.log
