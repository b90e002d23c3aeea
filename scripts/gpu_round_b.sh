#!/bin/bash
# GPU session B: parity tests (incl. batched Gram-Schmidt mode), kernel A/B at 20M, bench A/B at 2M, bench at 20M
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 spmv_F,ilu_F > gpurun_out/b_prof_u4.log 2>&1; cat gpurun_out/b_prof_u4.log
NSB_SELL_U=8 timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 spmv_F,ilu_F > gpurun_out/b_prof_u8.log 2>&1; cat gpurun_out/b_prof_u8.log
NSB_SPMV_PAD=0 timeout 300 python scripts/prof_kernels.py cyl3d-20M 1 5 spmv_F > gpurun_out/b_prof_nopad.log 2>&1; cat gpurun_out/b_prof_nopad.log
for o in 0 1; do
  timeout 600 python bench.py --workload cyl3d-2M --steps 3 --warmup 3 --no-cpu-baseline --orthogonalisation $o > gpurun_out/b_bench_2M_orth$o.json 2> gpurun_out/b_bench_2M_orth$o.err
  python - <<PY
import json
d=json.load(open("gpurun_out/b_bench_2M_orth$o.json"))
print("2M orth=$o", d["value"], d["ms_per_step"], d["detail"]["outer_iterations"], d["detail"]["last_step_counts"], d["e2e"]["value"])
PY
done
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/b_bench_20M.json 2> gpurun_out/b_bench_20M.err
cat gpurun_out/b_bench_20M.json
