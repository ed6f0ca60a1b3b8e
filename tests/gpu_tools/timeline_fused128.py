"""Developer tool: per-phase clock64() timeline of CTA 0 of the DH = 128 fused backward (needs the -DMLSTM_TIMELINE
build: python -m xlstm_yolo_b200.build --timeline)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from xlstm_yolo_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libmlstm_b200.so", "libmlstm_b200_tl.so")
from xlstm_yolo_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_gpu_parity import make
B, NH, S, DH = 32, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 1600, 128
q, k, v, i, f, dh = (x.cuda() for x in make(B, NH, S, DH, torch.bfloat16, "rand"))
pl = ops.MLSTMPlan(q, k, v, i, f, dh)
print("variants", pl.variant_fwd, pl.variant_bwd)
for _ in range(3):
    pl.forward(); pl.backward(1)
torch.cuda.synchronize()
raw = pl.ws.view(torch.uint8)[: 6 * 64 * 8].view(torch.int64).cpu().view(6, 2, 32)
t0 = raw[0, 0, 0].item()
cn = {0: "start", 1: "dh landed", 2: "dn done", 3: "dS' stored", 4: "E^T in regs", 5: "B1", 6: "colsum", 7: "out(Q) done", 8: "B2", 9: "B3",
      10: "Vs copied", 11: "Z^T done", 12: "K tile", 13: "out(V) done", 14: "B4", 15: "dV ld", 16: "B5", 17: "out(K) done", 18: "dk math",
      19: "bar5", 20: "state pass", 21: "B6"}
io = {0: "start", 5: "B1", 22: "cs,k ok", 8: "B2", 9: "B3", 14: "B4", 23: "dC done", 16: "B5", 24: "out(K) done", 25: "dv read", 21: "B6",
      26: "dk read", 27: "dh,v next", 28: "q,k next"}
for c in range(2, min(5, (S + 127) // 128)):
    for w, (who, names) in enumerate((("compute t0", cn), ("issuer    ", io))):
        row = raw[c, w]
        ev = sorted(((row[k_].item() - t0, n_) for k_, n_ in names.items() if row[k_].item() > 0))
        print(f"step {c} {who}: " + "  ".join(f"{n_}@{t_}" for t_, n_ in ev))
