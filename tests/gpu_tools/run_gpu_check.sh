#!/bin/bash
# usage (on the GPU box): bash tests/gpu_tools/run_gpu_check.sh [all|simt|tc]
mkdir -p gpurun_out
timeout 600 python tests/gpu_tools/gpu_check.py ${1:-all} > gpurun_out/gpu_check.log 2>&1; echo "gpu_check exit $?"; tail -40 gpurun_out/gpu_check.log
