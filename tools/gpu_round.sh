#!/bin/bash
# One GPU-box session: descriptor bring-up check, GPU test suite, attention A/B timing, bench line. Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
python tools/diag_vdesc.py > gpurun_out/diag_vdesc.log 2>&1; echo "diag_vdesc rc=$?"; tail -14 gpurun_out/diag_vdesc.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/attn_bench.py 32760 2 > gpurun_out/attn_bench.log 2>&1; echo "attn_bench rc=$?"; tail -24 gpurun_out/attn_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_1gpu.json; tail -5 gpurun_out/bench_1gpu.err
