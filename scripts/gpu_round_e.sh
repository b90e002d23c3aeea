#!/bin/bash
# GPU session E (2 GPUs): peer-memory transport vs NCCL -- parity tests and a short 2-rank bench at 2M
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_properties.py -x -q -m gpu > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -25 gpurun_out/e_pytest.log
for t in 1 0; do
  NSB_P2P=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload cyl3d-2M --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_2M_n2_p2p$t.json 2> gpurun_out/e_bench_2M_n2_p2p$t.err
  tail -3 gpurun_out/e_bench_2M_n2_p2p$t.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e_bench_2M_n2_p2p$t.json").read().strip().splitlines()[-1])
    print("2M N=2 p2p=$t", d["value"], d["ms_per_step"], d["detail"]["outer_iterations"], d["e2e"]["value"], d["config"].get("transport"))
except Exception as ex: print("no json", ex)
PY
done
