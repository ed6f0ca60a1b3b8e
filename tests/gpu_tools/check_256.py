"""Developer check of the head-dim-256 tcgen05 family against the fp64 oracle (prints per-tensor errors), then a timing of
the BASELINE-sized launch (B32 NH4 S1600 DH256) with an event pair around forward / backward parts.
    python tests/gpu_tools/check_256.py [n_cases] [--time]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_parity import make, oracle_on_kernel_side, rel, run_cuda  # noqa: E402
from xlstm_yolo_b200 import ops  # noqa: E402

cases = [(1, 1, 128, 256, "rand", False), (1, 2, 256, 256, "rand", False), (1, 2, 200, 256, "rand", False),
         (2, 2, 400, 256, "forget", True), (1, 2, 200, 256, "refinit", True), (3, 2, 129, 256, "rand", True),
         (1, 4, 1600, 256, "rand", False), (1, 2, 3200, 256, "rand", True)]
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if args:
    cases = cases[:int(args[0])]
bad = 0
for B, NH, S, DH, regime, rev in cases:
    inputs = make(B, NH, S, DH, torch.bfloat16, regime)
    try:
        t0 = time.time()
        fam = ops.kernel_family(inputs[0].cuda(), inputs[2].cuda())
        got = run_cuda(inputs, reverse=rev)
        ref = oracle_on_kernel_side(inputs, reverse=rev)
        errs = {n: rel(a, b) for n, a, b in zip(["h", "dq", "dk", "dv", "di", "df"], got, ref)}
        ok = errs["h"] < 1e-2 and all(v < 2e-2 for k, v in errs.items() if k != "h")
        bad += not ok
        print(f"B{B} NH{NH} S{S} DH{DH} {regime} rev={rev} [{fam}]: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()),
              "OK" if ok else "FAIL", f"({time.time() - t0:.1f}s)", flush=True)
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"B{B} NH{NH} S{S} DH{DH} {regime} rev={rev}: EXCEPTION {exc}", flush=True)
        break
print("failures:", bad)
if "--time" in sys.argv and not bad:
    for (B, NH, S, DH) in [(32, 4, 1600, 256), (8, 4, 1600, 256), (32, 4, 400, 256)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
        q, k = [(rn(B, S, NH, DH) * DH ** -0.5).bfloat16().transpose(1, 2) for _ in range(2)]
        v, dh = [rn(B, S, NH, DH).bfloat16().transpose(1, 2) for _ in range(2)]
        i = rn(B, S, NH).transpose(1, 2)
        f = (torch.linspace(3, 6, NH, device="cuda").view(1, 1, NH) + rn(B, S, NH)).transpose(1, 2)
        pl = ops.MLSTMPlan(q, k, v, i, f, dh, eps=1e-6, chunk_size=64)
        for _ in range(3):
            pl.forward(); pl.backward(0); pl.backward(1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        n = 10
        tot = [0.0, 0.0, 0.0]
        for _ in range(n):
            ev[0].record(); pl.forward(); ev[1].record(); pl.backward(0); ev[2].record(); pl.backward(1); ev[3].record()
            torch.cuda.synchronize()
            for j in range(3):
                tot[j] += ev[j].elapsed_time(ev[j + 1]) / n
        ms = sum(tot)
        flops = 3 * (4 * DH * DH + 2 * 65 * DH) * B * NH * S
        print(f"B{B} NH{NH} S{S} DH{DH} [{pl.family} {pl.variant_fwd}/{pl.variant_bwd}]: fwd {tot[0]*1e3:.0f} us, bwd A {tot[1]*1e3:.0f} us, "
              f"bwd rest {tot[2]*1e3:.0f} us -> {B*S/ms/1e3:.1f} M tok/s, {flops/ms/1e9:.0f} TFLOP/s algorithmic", flush=True)
sys.exit(1 if bad else 0)
