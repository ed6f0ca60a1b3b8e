"""Developer check of the DH = 128 fused-walk backward against the fp64 oracle (prints per-tensor errors).
    MLSTM_FORCE_VARIANT=13 python tests/gpu_tools/check_fused128.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import mlstm_oracle as O  # noqa: E402
from test_gpu_parity import make, oracle_on_kernel_side, rel, run_cuda  # noqa: E402
from xlstm_yolo_b200 import ops  # noqa: E402

cases = [(2, 2, 128, 128, "rand", False), (2, 2, 256, 128, "rand", False), (2, 4, 400, 128, "rand", False),
         (2, 4, 400, 128, "forget", True), (1, 4, 1600, 128, "refinit", True), (3, 2, 127, 128, "rand", True),
         (1, 2, 6400, 128, "rand", False)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
bad = 0
for B, NH, S, DH, regime, rev in cases:
    inputs = make(B, NH, S, DH, torch.bfloat16, regime)
    q = inputs[0].cuda()
    pl_variant = None
    try:
        t0 = time.time()
        got = run_cuda(inputs, reverse=rev)
        ref = oracle_on_kernel_side(inputs, reverse=rev)
        errs = {n: rel(a, b) for n, a, b in zip(["h", "dq", "dk", "dv", "di", "df"], got, ref)}
        ok = errs["h"] < 1e-2 and all(v < 2e-2 for k, v in errs.items() if k != "h")
        bad += not ok
        print(f"B{B} NH{NH} S{S} DH{DH} {regime} rev={rev}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()),
              "OK" if ok else "FAIL", f"({time.time() - t0:.1f}s)", flush=True)
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"B{B} NH{NH} S{S} DH{DH} {regime} rev={rev}: EXCEPTION {exc}", flush=True)
        break
print("variant:", os.environ.get("MLSTM_FORCE_VARIANT"), "failures:", bad)
sys.exit(1 if bad else 0)
