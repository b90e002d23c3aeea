#!/bin/bash
# GPU session H (8 GPUs): the 20M-DoF bench across 8 ranks over the peer-memory transport
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/h_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_20M_n8.json 2> gpurun_out/h_bench_20M_n8.err
echo "rc=$?"; tail -5 gpurun_out/h_bench_20M_n8.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/h_bench_20M_n8.json").read().strip().splitlines()[-1])
    print("20M N=8", d["value"], d["ms_per_step"], d["detail"]["outer_iterations"], d["detail"]["first_step_s"], d["detail"]["setup_s"], d["e2e"]["value"], d["config"].get("transport"))
    print({k:(v["ms"],round(v["gbs"])) for k,v in d["roofline"]["kernels"].items()})
except Exception as ex: print("no json", ex)
PY
