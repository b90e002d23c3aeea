#!/bin/bash
# Round 2, GPU session O (1 GPU): defaults after session M (PDL on, Schur factors in point multicolour order).
# 1. the new / changed tests (separate Schur ordering, 2D mid-size step against the oracle, 2D 2 M-DoF properties, C++ drivers)
# 2. the headline bench line (19.9 M DoF), short form of the driver's command
# 3. ncu launch list (gpu__time_duration) of a window of bench.py's time loop on the 0.27 M-DoF mesh
# 4. the 2D family (aSIMPLE) at 0.64 M DoF on one GPU with the reference's literals (at 2 M DoF the reference's inner
#    GMRES on the Schur complement hits its 10 000-iteration limit: sessions M and N)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_drivers.py -q -k "separate_ordering or 2d_ or driver" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -8 gpurun_out/r2o_pytest.log
# 1b. software prefetch of the block sweeps' gathers into L2 (NSB_BSELL_PREFETCH): parity, timings; the bench below uses it
#     when it wins by more than 3 % at 19.9 M DoF and the parity subset is green
NSB_BSELL_PREFETCH=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode or separate_ordering" > gpurun_out/r2o_pytest_prefetch.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_prefetch.log
tail -3 gpurun_out/r2o_pytest_prefetch.log
timeout 100 python scripts/prof_variants.py cyl3d-2M 2 10 "NSB_BSELL_PREFETCH=0;NSB_BSELL_PREFETCH=1" 2>&1 | tee gpurun_out/r2o_prof_2M_o2.log
timeout 200 python scripts/prof_variants.py cyl3d-20M 2 5 "NSB_BSELL_PREFETCH=0;NSB_BSELL_PREFETCH=1;NSB_BSELL_PREFETCH=0;NSB_BSELL_PREFETCH=1" 2>&1 | tee gpurun_out/r2o_prof_20M_o2.log
export NSB_BSELL_PREFETCH=$(python - <<'PY'
import re
t = {"0": [], "1": []}
try:
    for ln in open("gpurun_out/r2o_prof_20M_o2.log"):
        m = re.match(r"\[NSB_BSELL_PREFETCH=(\d)\s*\] ilu_F\s+([0-9.]+) ms", ln)
        if m:
            t[m.group(1)].append(float(m.group(2)))
    ok = "pytest rc=0" in open("gpurun_out/r2o_pytest_prefetch.log").read()
    print(1 if ok and t["0"] and t["1"] and min(t["1"]) < 0.97 * min(t["0"]) else 0)
except Exception:
    print(0)
PY
)
echo "NSB_BSELL_PREFETCH=$NSB_BSELL_PREFETCH for the bench"
NSB_BENCH_BUDGET_S=420 timeout 460 python bench.py --gpus 1 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2o_bench_20M.json 2> gpurun_out/r2o_bench_20M.err
echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2o_bench_20M.err | tail -12
timeout 120 python bench.py --workload cyl3d-270k --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2o_bench_270k.json 2> gpurun_out/r2o_bench_270k.err
echo "270k rc=$?"; grep -E "^\[bench" gpurun_out/r2o_bench_270k.err | tail -4
if [ -s gpurun_out/r2o_bench_270k.json ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 100000 --launch-count 8000 --csv \
    --log-file gpurun_out/r2o_ncu_launches_270k.csv python bench.py --workload cyl3d-270k --steps 1 --warmup 1 --no-cpu-baseline \
    > gpurun_out/r2o_ncu.log 2>&1
  echo "ncu rc=$?"; wc -l gpurun_out/r2o_ncu_launches_270k.csv
fi
timeout 150 python bench.py --workload cyl2d-640k --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r2o_bench_cyl2d_640k.json 2> gpurun_out/r2o_bench_cyl2d_640k.err
echo "cyl2d-640k rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2o_bench_cyl2d_640k.err | tail -6
