#!/bin/bash
# Round 2, GPU session G (1 GPU): ncu --set full of the block-multicolour sweep (k_bsell) and of the
# subdomain-resident part kernel (k_sd_trsv) at 19.9 M DoF: what limits them.
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bsell -s 3 -c 2 -o gpurun_out/r2g_bsell -f python scripts/prof_kernels.py cyl3d-20M 2 1 ilu_F > gpurun_out/r2g_ncu_bsell.log 2>&1
echo "bsell rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sd_trsv -c 2 -o gpurun_out/r2g_sd -f python scripts/prof_kernels.py cyl3d-20M 3 1 ilu_F > gpurun_out/r2g_ncu_sd.log 2>&1
echo "sd rc=$?"
ls -la gpurun_out/*.ncu-rep
