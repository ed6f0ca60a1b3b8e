#!/bin/bash
# each case in its own process so one CUDA fault does not hide the others
mkdir -p gpurun_out; : > gpurun_out/cases.log
while read -r line; do
  [ -z "$line" ] && continue
  timeout 120 python tests/gpu_tools/gpu_check.py case $line 2>&1 | grep -E "^B[0-9]|Error|error" | head -3 >> gpurun_out/cases.log
done <<'CASES'
1 1 128 64 bfloat16 rand 0 0 0.125
1 1 128 128 bfloat16 rand 0 0 0.09
2 4 400 64 bfloat16 rand 0 0 0.125
2 4 400 64 bfloat16 refinit 1 0 0.125
2 4 400 128 bfloat16 forget 1 0 0.09
1 4 1600 128 bfloat16 rand 0 0 0.09
1 2 300 128 bfloat16 rand 0 1 0.09
1 2 300 64 bfloat16 rand 1 1 0.125
2 2 1 64 bfloat16 rand 0 0 0.125
CASES
cat gpurun_out/cases.log
