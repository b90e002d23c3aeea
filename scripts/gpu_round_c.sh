#!/bin/bash
# GPU session C: all GPU tests; SELL lanes A/B at 2M; ILU chunk A/B and new assembly kernel at 20M; bench at 2M
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
for lanes in 4 1; do
  echo "== 2M lanes=$lanes"; NSB_SELL_LANES=$lanes timeout 300 python scripts/prof_kernels.py cyl3d-2M 1 10 assemble_step,spmv_F,ilu_F,ilu_S,dot,add_and_dot 2>&1 | tee gpurun_out/c_prof_2M_lanes$lanes.log
done
for chunk in 1250000 0 800000 2500000; do
  echo "== 20M chunk=$chunk"; NSB_ILU_CHUNK=$chunk timeout 400 python scripts/prof_kernels.py cyl3d-20M 1 5 ilu_F,spmv_F,assemble_step 2>&1 | tee gpurun_out/c_prof_20M_chunk$chunk.log
done
timeout 600 python bench.py --workload cyl3d-2M --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_2M.json 2> gpurun_out/c_bench_2M.err
python - <<PY
import json
d=json.load(open("gpurun_out/c_bench_2M.json"))
print("2M", d["value"], d["ms_per_step"], d["detail"]["outer_iterations"], d["detail"]["first_step_iterations"], d["detail"]["last_step_counts"], d["e2e"]["value"])
print({k:(v["ms"],v["gbs"]) for k,v in d["roofline"]["kernels"].items()})
PY
