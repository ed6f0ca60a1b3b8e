#!/bin/bash
# usage (on the GPU box): bash tests/gpu_tools/run_gpu_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 120 xlstm_yolo_b200/lib/probe_umma > gpurun_out/probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/probe.log
timeout 600 python tests/gpu_tools/gpu_check.py > gpurun_out/gpu_check.log 2>&1; echo "gpu_check exit $?"; tail -30 gpurun_out/gpu_check.log
