#!/bin/bash
# one `ncu --set full` capture of the library's kernels for a workload (after a clean plain run)
W=${1:-cfg2_B32_NH4_S400_DH64}
N=${2:-6}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'tc_|simt_' -s 15 -c $N -o gpurun_out/prof_$W -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $W > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -5
