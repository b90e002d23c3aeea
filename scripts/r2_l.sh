#!/bin/bash
# Round 2, GPU session L (2 GPUs): multi-rank parity after the transport changes (peer memory default, mailbox
# re-export), 2-rank bench at 2 M DoF with the default transport.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2l_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest_multi.log
tail -5 gpurun_out/r2l_pytest_multi.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --workload cyl3d-2M --steps 5 --warmup 2 > gpurun_out/r2l_bench_2M_n2.json 2> gpurun_out/r2l_bench_2M_n2.err
echo "2M n2 rc=$?"; grep -E "^\[bench" gpurun_out/r2l_bench_2M_n2.err | tail -6
