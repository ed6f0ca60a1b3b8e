# Developer A/B of the forward walk: warp-specialised (default) against the one-group kernel (MLSTM_FWD_WS=0)
for ws in 1 0; do
  for w in cfg2_B32_NH4_S400_DH64 cfg3_B32_NH4_S1600_DH128 cfg3_B32_NH4_S6400_DH128; do
    MLSTM_FWD_WS=$ws timeout 200 python bench.py --steps 50 --warmup 5 --no-also --no-cpu-baseline --workload $w 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ws=$ws $w', round(d['value']/1e6,1), round(d['roofline']['per_kernel_ms']['fwd'],4))"
  done
done
