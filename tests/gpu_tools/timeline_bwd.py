"""Developer tool: per-phase clock64() timeline of CTA 0 of backward kernel A (needs the
-DMLSTM_TIMELINE build: lib/libmlstm_b200_tl.so)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xlstm_yolo_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libmlstm_b200.so", "libmlstm_b200_tl.so")
from xlstm_yolo_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_parity import make
B, NH, S, DH = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 4, 1600, 128)))
q, k, v, i, f, dh = (x.cuda() for x in make(B, NH, S, DH, torch.bfloat16, "rand"))
pl = ops.MLSTMPlan(q, k, v, i, f, dh)
for _ in range(2):
    pl.forward(); pl.backward(0)
torch.cuda.synchronize()
rows = B * NH * S
mode_b2 = os.environ.get("TL_B2") == "1"     # library built with -DMLSTM_TL_MODE=2: stamps of kernel B2 in the R partials
kpart_off = rows * 4 + (0 if mode_b2 else 4 * rows * 4)          # BwdLayout: dn | rpart(4) | kpart(4)
pl.forward(); pl.backward(0)
if mode_b2:
    pl.backward(1)
torch.cuda.synchronize()
tl = pl.ws.view(torch.uint8)[kpart_off:kpart_off + 8 * 32 * 8].view(torch.int64).cpu().view(8, 32)
names = ["top", "dn+sync3", "MMA1-wait", "xfree-sync", "tile", "sync", "MMA2iss", "MMA2-wait", "epilogue", "end-sync", "store+MMA1iss"]
t0 = tl[0, 0].item()
for who, off in (("compute thread 0", 0), ("issuer", 16)):
    print(who)
    for c in range(8):
        row = tl[c, off:off + 11] - t0
        d = [(row[j] - row[j - 1]).item() for j in range(1, 11)]
        print(f"item {c} start {row[0].item():7d}  " + " ".join(f"{n_[:9]}:{x:5d}" for n_, x in zip(names[1:], d)))
