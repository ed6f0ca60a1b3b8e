#!/bin/bash
# Round-2 evidence run (one GPU).  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards):
#   the GPU test-suite, the default bench line (+ `also` block), its ncu launch list, `ncu --set full` captures of the DH = 256
#   family, of the operand producer / gate / tail kernels of a ViLBlockPair step and of the cfg2 / cfg3 steps, the layer bench,
#   the reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r02_gputests.log
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err || tail -5 gpurun_out/r02_bench_default.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_default_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
for spec in "cfg3alt_B32_NH4_S1600_DH256 21 7" "cfg2_B32_NH4_S400_DH64 15 4" "cfg3_B32_NH4_S1600_DH128 15 4"; do
  set -- $spec
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --workload $1 > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'tc' -s $2 -c $3 -o gpurun_out/r02_prof_$1 -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --workload $1 > gpurun_out/ncu_full_$1.log 2>&1
  echo "full capture $1 exit $?"
done
python tests/gpu_tools/layer_profile.py > gpurun_out/r02_layer_profile.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'qkv_|gates_|glue_|colsum' -s 20 -c 10 -o gpurun_out/r02_prof_layer -f \
    python tests/gpu_tools/layer_profile.py > gpurun_out/ncu_full_layer.log 2>&1
echo "full capture layer exit $?"
timeout 300 python tests/gpu_tools/layer_bench.py > gpurun_out/r02_layer_bench.txt 2>&1; tail -3 gpurun_out/r02_layer_bench.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference.json
ls -la gpurun_out/*.ncu-rep
