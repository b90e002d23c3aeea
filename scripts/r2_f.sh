#!/bin/bash
# Round 2, GPU session F (1 GPU): subdomain-resident ILU solves v2 (multi-level parts, bulk-copy pipeline):
# parity, kernel timings, per-launch times of one apply (ncu), bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -12 gpurun_out/r2f_pytest.log
if grep -q "pytest rc=0" gpurun_out/r2f_pytest.log; then
  echo "== 2M ordering=3"; timeout 200 python scripts/prof_kernels.py cyl3d-2M 3 10 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2f_prof_2M_o3.log
  echo "== 20M ordering=3"; NSB_VERBOSE=1 timeout 300 python scripts/prof_kernels.py cyl3d-20M 3 5 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2f_prof_20M_o3.log
  echo "== 20M ordering=3 one level"; NSB_SD_LEAF=3072 timeout 300 python scripts/prof_kernels.py cyl3d-20M 3 5 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2f_prof_20M_o3_1level.log
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sd_ -c 120 --csv --log-file gpurun_out/r2f_ncu_sd_20M.csv python scripts/prof_kernels.py cyl3d-20M 3 1 ilu_F,ilu_S > gpurun_out/r2f_ncu.log 2>&1
  NSB_BENCH_BUDGET_S=300 timeout 400 python bench.py --steps 3 --warmup 2 --ilu-ordering 3 --no-cpu-baseline > gpurun_out/r2f_bench_20M.json 2> gpurun_out/r2f_bench_20M.err
  echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2f_bench_20M.err | tail -20
fi
