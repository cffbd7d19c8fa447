#!/bin/bash
# compute-sanitizer memcheck over a subset of the kernel tests (small shapes of every tcgen05 / TMA kernel). Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
SEL='test_ln_modulate or test_ln_affine or test_rmsnorm_rope or staged_form or test_head_unpatchify or head_step or (test_attention_long_kernel and 300-1000) or (test_attention_long_kernel and 513-1285) or test_gemm_epilogues or test_gemm_strided or test_cfg_step'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --timeout 1400 -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
