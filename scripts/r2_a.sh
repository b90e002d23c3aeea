#!/bin/bash
# Round 2, GPU session A (2 GPUs): multi-rank parity over both transports (incl. the cube case that failed in
# round 1), 2-rank bench at 2 M DoF over both transports, and the new bench flow at 19.9 M DoF on one GPU.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2a_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2a_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_multi.log
tail -15 gpurun_out/r2a_pytest_multi.log
for p in 1 0; do
  NSB_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload cyl3d-2M --steps 3 --warmup 2 --ilu-ordering 1 > gpurun_out/r2a_bench_2M_n2_p2p$p.json 2> gpurun_out/r2a_bench_2M_n2_p2p$p.err
  echo "2M n2 p2p=$p rc=$?"; grep -E "^\[bench" gpurun_out/r2a_bench_2M_n2_p2p$p.err | tail -12
done
NSB_BENCH_BUDGET_S=520 timeout 600 python bench.py --steps 3 --warmup 3 --ilu-ordering 1 --no-cpu-baseline > gpurun_out/r2a_bench_20M.json 2> gpurun_out/r2a_bench_20M.err
echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2a_bench_20M.err | tail -20
