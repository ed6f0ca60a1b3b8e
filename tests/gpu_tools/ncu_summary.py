"""Developer tool: condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel summary kept under profiles/.

    ncu -i gpurun_out/prof_W.ncu-rep --page raw --csv > /tmp/raw.csv
    python tests/gpu_tools/ncu_summary.py /tmp/raw.csv > profiles/rNN_ncu_full_W_summary.csv
"""
import csv
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        # the tensor-pipe counter that sees tcgen05.mma (UTCHMMA): agrees with issued MMAs x cycles per MMA; the TPC-triage
        # "realtime" variant below reads several times lower for the same launch (and differently for identical launches)
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.sum.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max"]

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
w = csv.writer(sys.stdout)
for r in rows:
    w.writerow([r[i] if i >= 0 else "" for i in idx])
