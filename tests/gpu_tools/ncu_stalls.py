"""Developer tool: warp-stall reasons of one kernel from an `ncu --set full --import-source on` report (source page).

    python tests/gpu_tools/ncu_stalls.py gpurun_out/X.ncu-rep <kernel-name-regex> [n-th match] > profiles/rNN_ncu_source_<kernel>_stalls.csv
Sums the per-instruction stall-sample columns of the first launch that matches and lists the hottest instructions."""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"      # which of the matching launches (e.g. template instances share a name regex)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
name = next((l for l in lines[:start] if l.startswith('"Kernel Name"')), "")
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr, body = rows[0], [r for r in rows[1:] if len(r) == len(rows[0]) and r[0] != "Address"]
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
tot = {hdr[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
n = sum(tot.values())
print(f"# warp-stall samples by reason, {name.split(',')[1][:90] if name else kre} ({rep.split('/')[-1]}, one launch, {n} samples)")
print("reason,samples,pct")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"{k},{v},{100.0 * v / max(n, 1):.1f}")
si = hdr.index("# Samples")
print("# hottest instructions (samples, SASS)")
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:15]:
    print(f"{r[si]},{r[hdr.index('Source')].strip()}")
