#!/bin/bash
# Round-end evidence run (one GPU): GPU test-suite, default bench line + ncu launch list, one `ncu --set full` capture of the
# default workload, bench lines of the other BASELINE shapes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -3
bash tests/gpu_tools/prof_list.sh | tail -2
bash tests/gpu_tools/prof_full.sh cfg2_B32_NH4_S400_DH64 4 | head -1
bash tests/gpu_tools/prof_full.sh cfg3_B32_NH4_S1600_DH128 6 | head -1
timeout 200 python tests/gpu_tools/infer_bench.py 2>&1 | tail -8
for w in cfg2alt_B32_NH4_S400_DH128 cfg3_B32_NH4_S1600_DH128 cfg3_B32_NH4_S6400_DH128 ddp_B8_NH4_S1600_DH128; do
  timeout 200 python bench.py --no-cpu-baseline --workload $w > gpurun_out/bench_$w.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/bench_$w.json')); print(d['config']['workload'], round(d['value']/1e6,1), round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['roofline']['per_kernel_ms'].items()}, round(d['roofline']['step_hbm_frac'],3), d['roofline']['variants'], round(d['e2e']['value']/1e6,1))"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
