#!/bin/bash
# Round 2, GPU session C (1 GPU): whole GPU suite, block-multicolour diagnostics, kernel timings for the two
# throughput orderings at 19.9 M DoF, then bench.py with the driver's exact arguments.
mkdir -p gpurun_out
{ echo "DEAL_II_DIR=$DEAL_II_DIR mkDealiiPrefix=$mkDealiiPrefix"; which mpirun mpicxx cmake 2>&1; ls baseline/_ref 2>&1 | head -3;
  find / -xdev \( -iname "*deal.II*" -o -iname "libepetra*" -o -iname "libifpack*" \) 2>/dev/null | head -5; echo "probe done"; nproc; } > gpurun_out/r2c_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
timeout 200 python scripts/diag_bsell.py cyl3d-270k 2 oracle 2>&1 | tee gpurun_out/r2c_diag_270k.log
for o in 2 1; do
  echo "== 20M ordering=$o"; timeout 300 python scripts/prof_kernels.py cyl3d-20M $o 5 ilu_F,ilu_S,spmv_F,spmv_S,assemble_step 2>&1 | tee gpurun_out/r2c_prof_20M_o$o.log
done
echo "== 2M ordering=2"; timeout 200 python scripts/prof_kernels.py cyl3d-2M 2 10 ilu_F,ilu_S,spmv_F,spmv_S 2>&1 | tee gpurun_out/r2c_prof_2M_o2.log
timeout 1000 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench_20M.json 2> gpurun_out/r2c_bench_20M.err
echo "20M rc=$?"; grep -E "^\[bench" gpurun_out/r2c_bench_20M.err | tail -30; head -c 600 gpurun_out/r2c_bench_20M.json
