#!/bin/bash
# Round 2, GPU session K (8 GPUs): the 19.9 M-DoF bench on 8 ranks (NCCL default, then peer memory), 4-rank parity tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r2k_gpus.txt
NSB_VERBOSE=1 NSB_BENCH_BUDGET_S=230 timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus 8 --steps 8 --warmup 2 > gpurun_out/r2k_bench_20M_n8.json 2> gpurun_out/r2k_bench_20M_n8.err
echo "20M n8 nccl rc=$?"; grep -E "^\[bench|nsb setup rank 0" gpurun_out/r2k_bench_20M_n8.err | tail -24
NSB_P2P=1 NSB_BENCH_BUDGET_S=170 timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 \
  bench.py --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2k_bench_20M_n8_p2p.json 2> gpurun_out/r2k_bench_20M_n8_p2p.err
echo "20M n8 p2p rc=$?"; grep -E "^\[bench" gpurun_out/r2k_bench_20M_n8_p2p.err | tail -12
timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "4]" > gpurun_out/r2k_pytest_multi4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_multi4.log
tail -6 gpurun_out/r2k_pytest_multi4.log
