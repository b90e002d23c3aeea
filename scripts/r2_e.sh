#!/bin/bash
# Round 2, GPU session E (1 GPU): subdomain-resident ILU solves (ilu_ordering = 3): parity, kernel timings, bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode or drag_lift" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -12 gpurun_out/r2e_pytest.log
if grep -q "pytest rc=0" gpurun_out/r2e_pytest.log; then
  echo "== 2M ordering=3"; NSB_VERBOSE=1 timeout 200 python scripts/prof_kernels.py cyl3d-2M 3 10 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/r2e_prof_2M_o3.log
  echo "== 20M ordering=3"; NSB_VERBOSE=1 timeout 300 python scripts/prof_kernels.py cyl3d-20M 3 5 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/r2e_prof_20M_o3.log
  echo "== 20M ordering=3 leaf 4096"; NSB_SD_LEAF=4096 timeout 300 python scripts/prof_kernels.py cyl3d-20M 3 5 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2e_prof_20M_o3_l4096.log
  NSB_BENCH_BUDGET_S=300 timeout 400 python bench.py --steps 3 --warmup 2 --ilu-ordering 3 --no-cpu-baseline > gpurun_out/r2e_bench_20M.json 2> gpurun_out/r2e_bench_20M.err
  echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2e_bench_20M.err | tail -20
fi
