#!/bin/bash
# Round 2, GPU sessions S1 (1 GPU) / S2 (2 GPUs): the 2D family (aSIMPLE) at 0.64 M DoF -- the largest refinement on which
# the reference's inner GMRES on the Schur complement converges within its 10 000 iterations -- with the final defaults.
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 150 python bench.py --workload cyl2d-640k --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2s_bench_cyl2d_640k_n1.json 2> gpurun_out/r2s_bench_cyl2d_640k_n1.err
  echo "n1 rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2s_bench_cyl2d_640k_n1.err | tail -6
else
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --workload cyl2d-640k --steps 3 --warmup 1 > gpurun_out/r2s_bench_cyl2d_640k_n$N.json 2> gpurun_out/r2s_bench_cyl2d_640k_n$N.err
  echo "n$N rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2s_bench_cyl2d_640k_n$N.err | tail -6
fi
