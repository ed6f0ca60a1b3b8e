#!/bin/bash
# item-order experiments for the chunk-parallel backward (MLSTM_BWD_ORDER = orders of kernels A, B1, B2)
for w in cfg3_B32_NH4_S1600_DH128 cfg3_B32_NH4_S6400_DH128; do
  for o in 000 212 112 221 211 121 012 021; do
    MLSTM_BWD_ORDER=$o python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('$w'[5:], '$o', round(d['value']/1e6,1), round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['per_kernel_ms'].items()})"
  done
done
