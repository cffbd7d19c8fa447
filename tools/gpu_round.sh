#!/bin/bash
# One GPU-box session: GPU test suite, attention A/B timing, ncu capture of the attention kernel, bench line. Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 300 python tools/attn_bench.py 32760 2 > gpurun_out/attn_bench.log 2>&1; echo "attn_bench rc=$?"; tail -16 gpurun_out/attn_bench.log
timeout 300 python tools/attn_bench.py 4095 2 > gpurun_out/attn_bench_4095.log 2>&1; echo "attn_bench4095 rc=$?"; tail -24 gpurun_out/attn_bench_4095.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 7000 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
MC_ATTN_EMU=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_long -s 1 -c 1 -o gpurun_out/r02_attn_long python tools/one_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn.log
