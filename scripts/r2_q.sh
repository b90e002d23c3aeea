#!/bin/bash
# Round 2, GPU session Q (1 GPU): k_bsell with the software-pipelined walk over the four passes (NSB_BSELL_PIPE=1):
# parity of the block orderings, timings at 2 M and 19.9 M DoF with / without the L2 prefetch.
mkdir -p gpurun_out
NSB_BSELL_PIPE=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "multicolour_ilu_mode or separate_ordering or batched_gram" > gpurun_out/r2q_pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest_pipe.log
tail -3 gpurun_out/r2q_pytest_pipe.log
V="NSB_BSELL_PIPE=0;NSB_BSELL_PIPE=1;NSB_BSELL_PIPE=1,NSB_BSELL_PREFETCH=0;NSB_BSELL_PIPE=0;NSB_BSELL_PIPE=1"
timeout 100 python scripts/prof_variants.py cyl3d-2M 2 10 "$V" 2>&1 | tee gpurun_out/r2q_prof_2M_o2.log
timeout 200 python scripts/prof_variants.py cyl3d-20M 2 5 "$V" 2>&1 | tee gpurun_out/r2q_prof_20M_o2.log
