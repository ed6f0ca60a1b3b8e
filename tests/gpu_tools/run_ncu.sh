#!/bin/bash
# ncu launch list (default bench command) + one full capture of the top kernels (cfg3 shape)
mkdir -p gpurun_out
W=${1:-cfg3_B32_NH4_S1600_DH128}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 15 -c 5 -o gpurun_out/prof -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
