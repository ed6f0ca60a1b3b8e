"""One shape of the gate projection forward, a few calls (for an ncu capture of gates_fwd_tc_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from xlstm_yolo_b200 import ops

B, S, D, NH = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 6400, 512, 4
q, k, v = (torch.randn(B, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
w_i, w_f = (torch.randn(NH, 3 * D, device="cuda") * 0.05 for _ in range(2))
b_i, b_f = (torch.randn(NH, device="cuda") for _ in range(2))
for _ in range(3):
    ops.gate_proj_fwd_raw(q, k, v, w_i, b_i, w_f, b_f, NH)
torch.cuda.synchronize()
