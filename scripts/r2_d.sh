#!/bin/bash
# Round 2, GPU session D (2 GPUs): N-rank parity over both transports (incl. the cube p2p case that failed in
# round 1, and drag/lift across ranks), device drag/lift parity, 2-rank benches at 2 M DoF over both
# transports and at 19.9 M DoF.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2d_gpus.txt
timeout 700 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_drag_lift_coefficients" -q > gpurun_out/r2d_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_multi.log
tail -15 gpurun_out/r2d_pytest_multi.log
for p in 1 0; do
  NSB_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload cyl3d-2M --steps 5 --warmup 2 --ilu-ordering 1 > gpurun_out/r2d_bench_2M_n2_p2p$p.json 2> gpurun_out/r2d_bench_2M_n2_p2p$p.err
  echo "2M n2 p2p=$p rc=$?"; grep -E "^\[bench" gpurun_out/r2d_bench_2M_n2_p2p$p.err | tail -12
done
NSB_VERBOSE=1 NSB_BENCH_BUDGET_S=330 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus 2 --steps 6 --warmup 2 --ilu-ordering 1 > gpurun_out/r2d_bench_20M_n2.json 2> gpurun_out/r2d_bench_20M_n2.err
echo "20M n2 rc=$?"; grep -E "^\[bench|nsb setup rank 0" gpurun_out/r2d_bench_20M_n2.err | tail -30
