"""Developer tool: clock64() timeline of CTA 0 of the single-pass forward kernel at the cfg2 shape, prologue included
(needs lib/libmlstm_b200_tl.so: tests/gpu_tools/build_tl.sh)."""
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
for _ in range(3):
    pl.forward()
torch.cuda.synchronize()
pl.ws.zero_()
pl.forward()
torch.cuda.synchronize()
raw = pl.ws.view(torch.uint8)[: 1100 * 8].view(torch.int64).cpu()
hd = raw[1024:1040].view(2, 8)
t0 = hd[0, 0].item()
names = ["top", "qk-landed", "qn", "MMA1-done", "P", "sync1", "npart+Kbar", "sync2", "stIss+MMA2w", "epi1", "state-done", "statepass", "end-sync"]
for w, (who, off) in enumerate((("compute thread 0", 0), ("issuer", 16))):
    h_ = (hd[w] - t0).tolist()
    print(f"{who}: entry {h_[0]} setup-done {h_[1]} gates+loads-sync {h_[2]} loop-start {h_[3]} loop-end {h_[4]} drained {h_[5]}")
    for c in range(4):
        row = raw[c * 32 + off: c * 32 + off + 13] - t0
        d = [(row[j] - row[j - 1]).item() for j in range(1, 13)]
        print(f" chunk {c} start {row[0].item():7d}  " + " ".join(f"{n[:9]}:{x:5d}" for n, x in zip(names[1:], d)))
