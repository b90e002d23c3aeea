#!/bin/bash
# Round 2, GPU session I (1 GPU): block multicolour sweeps with the 4-pass ext phase and outside rows staged in shared memory:
# parity, kernel timings, per-launch times (ncu), bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode or batched_gram" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
if grep -q "pytest rc=0" gpurun_out/r2i_pytest.log; then
  echo "== 2M ordering=2"; timeout 200 python scripts/prof_kernels.py cyl3d-2M 2 10 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2i_prof_2M_o2.log
  echo "== 20M ordering=2"; NSB_VERBOSE=1 timeout 300 python scripts/prof_kernels.py cyl3d-20M 2 5 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2i_prof_20M_o2.log
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bsell -c 60 --csv --log-file gpurun_out/r2i_ncu_bsell_20M.csv python scripts/prof_kernels.py cyl3d-20M 2 1 ilu_F,ilu_S > gpurun_out/r2i_ncu.log 2>&1
fi
