#!/bin/bash
# Round 2, GPU session O (1 GPU): defaults after session M (PDL on, Schur factors in point multicolour order).
# 1. the new / changed tests (separate Schur ordering, 2D mid-size step against the oracle, 2D 2 M-DoF properties, C++ drivers)
# 2. the headline bench line (19.9 M DoF), short form of the driver's command
# 3. ncu launch list (gpu__time_duration) of a window of bench.py's time loop on the 0.27 M-DoF mesh
# 4. configs[3] (cyl2d-2M, aSIMPLE) on one GPU with the reference's literals
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_drivers.py -q -k "separate_ordering or 2d_ or driver" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -8 gpurun_out/r2o_pytest.log
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
timeout 330 python bench.py --workload cyl2d-2M --steps 2 --warmup 0 --no-cpu-baseline > gpurun_out/r2o_bench_cyl2d_2M.json 2> gpurun_out/r2o_bench_cyl2d_2M.err
echo "cyl2d-2M rc=$?"; grep -E "^\[bench|NsbError" gpurun_out/r2o_bench_cyl2d_2M.err | tail -6
