#!/bin/bash
# Round 2, GPU session P (1 GPU): ncu --set full of the block sweeps of F_s at 19.9 M DoF as they are now (PDL, software
# prefetch, 4-pass ext phase); the 2D family at 0.64 M DoF with the Schur factors in block multicolour order.
mkdir -p gpurun_out
timeout 150 python bench.py --workload cyl2d-640k --steps 3 --warmup 1 --no-cpu-baseline --ilu-ordering 1 --ilu-ordering-schur 2 > gpurun_out/r2p_bench_cyl2d_640k_s2.json 2> gpurun_out/r2p_bench_cyl2d_640k_s2.err
echo "cyl2d-640k (schur ordering 2) rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2p_bench_cyl2d_640k_s2.err | tail -6
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_bsell --launch-skip 37 --launch-count 4 -f -o gpurun_out/r2p_ncu_bsell_20M \
  python scripts/prof_kernels.py cyl3d-20M 2 1 ilu_F > gpurun_out/r2p_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2p_ncu.log; ls -la gpurun_out/r2p_ncu_bsell_20M.ncu-rep
