"""Developer tool: per-phase clock64() timeline of CTA 0 of the forward kernel (needs the
-DMLSTM_TIMELINE build: lib/libmlstm_b200_tl.so)."""
import os, sys, ctypes as C
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xlstm_yolo_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libmlstm_b200.so", "libmlstm_b200_tl.so")
from xlstm_yolo_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_parity import make
B, NH, S, DH = 32, 4, 1600, 128
q, k, v, i, f, dh = (x.cuda() for x in make(B, NH, S, DH, torch.bfloat16, "rand"))
pl = ops.MLSTMPlan(q, k, v, i, f, dh)
for _ in range(3):
    pl.forward()
torch.cuda.synchronize()
pl.ws.zero_()
pl.forward()
torch.cuda.synchronize()
tl = pl.ws[: 32 * 2 * 13].view(torch.int64).cpu().view(-1, 32)
names = ["top", "qk-landed", "qn", "MMA1-done", "P", "sync1", "npart+Kbar", "sync2",
         "stIss+MMA2w", "epi1", "state-done", "statepass", "end-sync"]
t0 = tl[0, 0].item()
for who, off in (("compute thread 0", 0), ("issuer", 16)):
    print(who)
    for c in range(13):
        row = tl[c, off:off + 13] - t0
        d = [(row[j] - row[j - 1]).item() for j in range(1, 13)]
        print(f"chunk {c:2d} start {row[0].item():7d}  " + " ".join(f"{n[:9]}:{x:5d}" for n, x in zip(names[1:], d)))
print("total cycles", (tl[12, 12] - t0).item())
