#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_bsell.py cyl3d-270k 2 oracle 2>&1 | tee gpurun_out/r2b2_diag_270k.log
timeout 300 python scripts/diag_bsell.py cyl3d-2M 2 2>&1 | tee gpurun_out/r2b2_diag_2M.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -8 gpurun_out/r2b_pytest.log
if grep -q "pytest rc=0" gpurun_out/r2b_pytest.log; then
  for o in 2 1; do
    echo "== 20M ordering=$o"; timeout 300 python scripts/prof_kernels.py cyl3d-20M $o 5 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/r2b_prof_20M_o$o.log
  done
  echo "== 2M ordering=2"; timeout 200 python scripts/prof_kernels.py cyl3d-2M 2 10 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/r2b_prof_2M_o2.log
  NSB_BENCH_BUDGET_S=420 timeout 500 python bench.py --steps 3 --warmup 2 --ilu-ordering 2 --no-cpu-baseline > gpurun_out/r2b_bench_20M.json 2> gpurun_out/r2b_bench_20M.err
  echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2b_bench_20M.err | tail -20
fi
