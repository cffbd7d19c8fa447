#!/bin/bash
# GPU check of the paths added late in round 2: natively sequenced forward (mc_dit_forward), Wan2.2 TI2V-5B forward, TI2V-5B timing.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_native_forward_gpu.py "tests/test_wan_forward_gpu.py::test_ti2v_per_token_timesteps_vs_oracle_and_fp64" -q -x --timeout 500 -s > gpurun_out/pytest_new_paths.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_new_paths.log | cut -c1-300
timeout 400 python tools/bench_ti2v.py --steps 6 > gpurun_out/bench_ti2v.json 2> gpurun_out/bench_ti2v.err; echo "ti2v rc=$?"; cat gpurun_out/bench_ti2v.json | cut -c1-1500; tail -5 gpurun_out/bench_ti2v.err
