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
raw = pl.ws.view(torch.uint8)[: 840 * 8].view(torch.int64).cpu()
tl = raw[:768].view(8, 4, 24)
hd = raw[800:800 + 32].view(4, 8)
names = ["top", "P1 dn", "G1 wait", "Vload", "P2 all", "sync", "G2 issue", "G2 wait", "P3+P4 epi", "sync", "G3 wait", "P5 state", "end sync", "tail"]
t0 = hd[0, 0].item()
whos = ["thread 0 (rg0,cq0 diag)", "thread 96 (rg3,cq0)", "thread 480 (rg3,cq3 diag)", "issuer"]
for w, who in enumerate(whos):
    h_ = (hd[w] - t0).tolist()
    print(f"{who}: entry {h_[0]} prologue-sync {h_[1]} gates+loads-sync {h_[2]} loop-start {h_[3]} loop-end {h_[4]} drained {h_[5]} exit {h_[6]}")
    for c in range(4):
        row = tl[c, w] - t0
        d = [(row[j] - row[j - 1]).item() for j in range(1, 14)]
        extra = ""
        if w < 3:
            extra = (f"  [P2: ldZ {(row[16] - row[3]).item()} dS {(row[17] - row[16]).item()} ldS {(row[18] - row[17]).item()} E {(row[14] - row[18]).item()} "
                     f"sync5+st {(row[4] - row[14]).item()} | colsum in G2 shadow {(row[15] - row[6]).item()}]")
        print(f" step {c} start {row[0].item():7d}  " + " ".join(f"{n_[:8]}:{x:5d}" for n_, x in zip(names[1:], d)) + extra)
