#!/bin/bash
# Round 2, GPU session J (1 GPU): k_bsell with capped staging (128 rows) + dependants mask: parity, kernel timings
# at 2 M and 19.9 M DoF for staging caps 128 / 64 / 0, and the 2 M-DoF bench for orderings 1 and 2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode or batched_gram" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
if grep -q "pytest rc=0" gpurun_out/r2j_pytest.log; then
  for cap in 128 64 0; do
    echo "== 2M ordering=2 cap=$cap"; NSB_BSELL_XCAP=$cap timeout 200 python scripts/prof_kernels.py cyl3d-2M 2 10 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2j_prof_2M_o2_cap$cap.log
  done
  for cap in 128 0; do
    echo "== 20M ordering=2 cap=$cap"; NSB_BSELL_XCAP=$cap timeout 300 python scripts/prof_kernels.py cyl3d-20M 2 5 ilu_F,ilu_S 2>&1 | tee gpurun_out/r2j_prof_20M_o2_cap$cap.log
  done
  for o in 2 1; do
    timeout 200 python bench.py --workload cyl3d-2M --steps 6 --warmup 2 --ilu-ordering $o --no-cpu-baseline > gpurun_out/r2j_bench_2M_o$o.json 2> gpurun_out/r2j_bench_2M_o$o.err
    echo "2M o$o rc=$?"; grep -E "^\[bench" gpurun_out/r2j_bench_2M_o$o.err | tail -4
  done
fi
