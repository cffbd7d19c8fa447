#!/bin/bash
# One GPU-box session of evidence: the GPU test suite, the driver's bench invocation, the ncu launch list of a short bench run and
# `ncu --set full` captures of the dominant kernels. Logs and reports under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','sec_per_video','speedup_vs_noncached','forward_ms')}, d['e2e']['value'], d['clocks'])
print(d['roofline']['frac'], d['roofline']['achieved'], d['kernels']['attn_self']['ms_avg'], d['hit_path']['frac'], d.get('cpu_baseline'))
print(d['attribution']['share_of_forward_time'])
PY
tail -3 gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; tail -c 600 gpurun_out/bench_reference.json
# our kernels only (torch's weight-initialisation launches would fill the capture window): three forwards' worth
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(attn_|gemm_bf16|ln_modulate|rmsnorm|head_|axpb|stats_|cfg_|patchify|time_sinusoid|linear_f32|cast_|residual|transpose|rel_l1|colmean|silu)' -c 1700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
[ "${NCU_FULL:-0}" = 1 ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_long -s 1 -c 1 -o gpurun_out/r02_attn_final python tools/one_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
[ "${NCU_FULL:-0}" = 1 ] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:head_tc -s 2 -c 1 -o gpurun_out/r02_head_final python tools/one_head.py > gpurun_out/ncu_head.log 2>&1; echo "ncu head rc=$?"
[ "${NCU_FULL:-0}" = 1 ] && timeout 400 ncu --set full --clock-control none -k regex:tma_kernel -s 8 -c 4 -o gpurun_out/r02_rowwise_final python tools/rowwise_bench.py > gpurun_out/ncu_rowwise.log 2>&1; echo "ncu rowwise rc=$?"
timeout 200 python tools/rowwise_bench.py 2>&1 | tail -6
