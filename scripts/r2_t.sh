#!/bin/bash
# Round 2, GPU session T (1 GPU): the pipelined walk for the 3-component block at 80 registers / 6 CTAs per SM.
mkdir -p gpurun_out
timeout 200 python scripts/prof_variants.py cyl3d-20M 2 5 "NSB_BSELL_PIPE=0;NSB_BSELL_PIPE=1;NSB_BSELL_PIPE=0;NSB_BSELL_PIPE=1" 2>&1 | tee gpurun_out/r2t_prof_20M_o2.log
