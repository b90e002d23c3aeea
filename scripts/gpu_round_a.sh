#!/bin/bash
# GPU session A: parity tests, kernel microbenchmarks at the 20M workload (A/B of the SELL sort window), ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 > gpurun_out/a_prof_w4096.log 2>&1
cat gpurun_out/a_prof_w4096.log
NSB_SELL_WINDOW=1073741824 timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 ilu_F > gpurun_out/a_prof_winf.log 2>&1
cat gpurun_out/a_prof_winf.log
NSB_SELL_WINDOW=512 timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 ilu_F > gpurun_out/a_prof_w512.log 2>&1
cat gpurun_out/a_prof_w512.log
NSB_SELL_WINDOW=32768 timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 ilu_F > gpurun_out/a_prof_w32k.log 2>&1
cat gpurun_out/a_prof_w32k.log
# ncu: full sections for the SELL kernels (SpMV F and the ILU sweeps) and the reductions
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_sell3|k_dot|k_add_and_dot" --launch-skip 6 -c 48 \
  -o gpurun_out/a_ncu_sell -f python scripts/prof_kernels.py cyl3d-20M 1 1 spmv_F,ilu_F,dot,add_and_dot > gpurun_out/a_ncu_sell.log 2>&1
tail -5 gpurun_out/a_ncu_sell.log
