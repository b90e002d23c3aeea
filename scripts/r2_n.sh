#!/bin/bash
# Round 2, GPU session N (2 GPUs): the C++ drivers one process per GPU (rendezvous.hpp, nsh_local_*) against the Python
# multi-rank class; configs[3] (cyl2d-2M, aSIMPLE) on 2 ranks; 2-rank 2 M-DoF 3D bench with the session-M defaults.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_multi.py -q -k "cpp" > gpurun_out/r2n_pytest_cpp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest_cpp.log
tail -15 gpurun_out/r2n_pytest_cpp.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 2 --workload cyl2d-2M --steps 4 --warmup 1 > gpurun_out/r2n_bench_cyl2d_2M_n2.json 2> gpurun_out/r2n_bench_cyl2d_2M_n2.err
echo "cyl2d-2M n2 rc=$?"; grep -E "^\[bench" gpurun_out/r2n_bench_cyl2d_2M_n2.err | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 \
  bench.py --gpus 2 --workload cyl3d-2M --steps 5 --warmup 2 > gpurun_out/r2n_bench_2M_n2.json 2> gpurun_out/r2n_bench_2M_n2.err
echo "cyl3d-2M n2 rc=$?"; grep -E "^\[bench" gpurun_out/r2n_bench_2M_n2.err | tail -5
