#!/bin/bash
mkdir -p gpurun_out
W=${1:-cfg3_B32_NH4_S1600_DH128}; K=${2:-tc_}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s 10 -c 4 -o gpurun_out/prof2 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/ncu_full2.log 2>&1
echo "full capture exit $?"
