#!/bin/bash
# Round 2, GPU session M (1 GPU): sweep kernels restructured for programmatic dependent launch (PDL) + L2 knobs.
# 1. whole GPU suite with NSB_PDL=1 (new launch path), the ILU-mode subset with PDL off (restructured kernels, old launches)
# 2. ILU applies at 2 M and 19.9 M DoF under the knob variants (one setup per ordering)
# 3. configs[3] (cyl2d-2M, aSIMPLE) bench on one GPU
mkdir -p gpurun_out
NSB_PDL=1 timeout 420 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest_pdl.log 2>&1; echo "pytest(pdl) rc=$?" >> gpurun_out/r2m_pytest_pdl.log
tail -4 gpurun_out/r2m_pytest_pdl.log
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "multicolour or batched" > gpurun_out/r2m_pytest_nopdl.log 2>&1; echo "pytest(nopdl) rc=$?" >> gpurun_out/r2m_pytest_nopdl.log
tail -3 gpurun_out/r2m_pytest_nopdl.log
V2="NSB_PDL=0;NSB_L2_FETCH=32;NSB_L2_FETCH=128;NSB_PDL=1;NSB_PDL=1,NSB_L2_FETCH=32"
timeout 150 python scripts/prof_variants.py cyl3d-2M 1 10 "$V2" 2>&1 | tee gpurun_out/r2m_prof_2M_o1.log
timeout 150 python scripts/prof_variants.py cyl3d-2M 2 10 "$V2;NSB_BSELL_OCC=10;NSB_PDL=1,NSB_BSELL_OCC=10" 2>&1 | tee gpurun_out/r2m_prof_2M_o2.log
V20="NSB_PDL=0;NSB_L2_PERSIST_MB=48;NSB_L2_PERSIST_MB=80;NSB_L2_PERSIST_MB=112;NSB_L2_FETCH=32;NSB_L2_FETCH=128;NSB_PDL=1;NSB_PDL=1,NSB_L2_PERSIST_MB=80"
NSB_VERBOSE=1 timeout 240 python scripts/prof_variants.py cyl3d-20M 2 5 "$V20;NSB_BSELL_OCC=10;NSB_PDL=1,NSB_BSELL_OCC=10;NSB_PDL=1,NSB_BSELL_OCC=10,NSB_L2_PERSIST_MB=80" 2>&1 | tee gpurun_out/r2m_prof_20M_o2.log
NSB_VERBOSE=1 timeout 240 python scripts/prof_variants.py cyl3d-20M 1 5 "$V20" 2>&1 | tee gpurun_out/r2m_prof_20M_o1.log
timeout 240 python bench.py --workload cyl2d-2M --steps 4 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_bench_cyl2d_2M.json 2> gpurun_out/r2m_bench_cyl2d_2M.err
echo "cyl2d-2M rc=$?"; grep -E "^\[bench" gpurun_out/r2m_bench_cyl2d_2M.err | tail -5
