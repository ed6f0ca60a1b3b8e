"""Developer tool: per-phase clock64() timeline of CTA 0 of the fused backward kernel (needs the -DMLSTM_TIMELINE
build: lib/libmlstm_b200_tl.so)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xlstm_yolo_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libmlstm_b200.so", "libmlstm_b200_tl.so")
from xlstm_yolo_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_parity import make
B, NH, S, DH = 32, 4, 400, 64
q, k, v, i, f, dh = (x.cuda() for x in make(B, NH, S, DH, torch.bfloat16, "rand"))
pl = ops.MLSTMPlan(q, k, v, i, f, dh)
print("variants", pl.variant_fwd, pl.variant_bwd)
for _ in range(3):
    pl.forward(); pl.backward(1)
torch.cuda.synchronize()
tl = pl.ws.view(torch.uint8)[: 8 * 32 * 8].view(torch.int64).cpu().view(8, 32)
names = ["top", "P1 dn", "G1 wait", "sync0", "P2 tiles", "sync", "G2 issue", "G2 wait", "P3+P4 epi", "sync", "G3 wait", "P5 state", "end sync", "tail"]
t0 = tl[0, 0].item()
for who, off in (("compute thread 0", 0), ("issuer", 16)):
    print(who)
    for c in range(4):
        row = tl[c, off:off + 14] - t0
        d = [(row[j] - row[j - 1]).item() for j in range(1, 14)]
        print(f"step {c} start {row[0].item():7d}  " + " ".join(f"{n_[:8]}:{x:5d}" for n_, x in zip(names[1:], d)))
