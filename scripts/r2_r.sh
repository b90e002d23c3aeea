#!/bin/bash
# Round 2, GPU session R (1 GPU): configs[3] -- the 2 M-DoF refined 2D cylinder, aSIMPLE, reference literals -- with the Schur
# factors in block multicolour order (pipelined block sweeps): one timed step after the two start-up steps.
mkdir -p gpurun_out
timeout 330 python bench.py --workload cyl2d-2M --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2r_bench_cyl2d_2M.json 2> gpurun_out/r2r_bench_cyl2d_2M.err
echo "cyl2d-2M rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2r_bench_cyl2d_2M.err | tail -8
