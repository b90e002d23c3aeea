#!/bin/bash
# GPU session D: tests; ILU apply at 20M with the 4-lane layout, with and without chunking
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
for cfg in "4 0" "4 1250000" "4 2500000" "4 600000" "1 0"; do
  set -- $cfg
  echo "== 20M lanes=$1 chunk=$2"; NSB_SELL_LANES=$1 NSB_ILU_CHUNK=$2 timeout 400 python scripts/prof_kernels.py cyl3d-20M 1 5 ilu_F,spmv_F 2>&1 | tee gpurun_out/d_prof_20M_l$1_c$2.log
done
