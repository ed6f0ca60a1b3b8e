#!/bin/bash
# Round-2 closing evidence run (one GPU), after the tensor-core gate forward and the shared state-walk / dq launch.
# Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_gputests.log
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err || tail -5 gpurun_out/r02_bench_default.err
timeout 300 python tests/gpu_tools/layer_bench.py > gpurun_out/r02_layer_bench.txt 2>&1; tail -3 gpurun_out/r02_layer_bench.txt
python tests/gpu_tools/layer_profile.py > gpurun_out/r02_layer_profile.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gates_fwd_tc' -s 4 -c 1 -o gpurun_out/r02_prof_gates_tc -f \
    python tests/gpu_tools/layer_profile.py > gpurun_out/ncu_full_gates_tc.log 2>&1
echo "full capture gates_tc exit $?"
head -12 gpurun_out/r02_layer_profile.txt
