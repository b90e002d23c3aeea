#!/bin/bash
# Round 2, GPU session U (1 GPU): the whole GPU suite with the final defaults, as the driver runs it.
mkdir -p gpurun_out
timeout 225 python -m pytest tests -x -q -m gpu > gpurun_out/r2u_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest_gpu.log
tail -5 gpurun_out/r2u_pytest_gpu.log
