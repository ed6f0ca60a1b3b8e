"""Gate projection: hand-written kernels vs the cat + nn.Linear formulation of the reference (cuBLAS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from xlstm_yolo_b200 import ops

def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for (B, S, D, NH) in [(32, 400, 256, 4), (32, 1600, 512, 4), (32, 6400, 512, 4)]:
    T = B * S
    q, k, v = (torch.randn(B, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    w_i, w_f = (torch.randn(NH, 3 * D, device="cuda") * 0.05 for _ in range(2))
    b_i, b_f = (torch.randn(NH, device="cuda") for _ in range(2))
    di, df = (torch.randn(B, S, NH, device="cuda") for _ in range(2))
    dq, dk, dv = (torch.randn(B, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    t_f = timeit(lambda: ops.gate_proj_fwd_raw(q, k, v, w_i, b_i, w_f, b_f, NH))
    t_b = timeit(lambda: ops.gate_proj_bwd_raw(q, k, v, w_i, w_f, NH, di, df, dq, dk, dv))
    wcat = torch.cat([w_i, w_f]).to(torch.bfloat16)
    bcat = torch.cat([b_i, b_f]).to(torch.bfloat16)
    def ref_fwd():
        return F.linear(torch.cat([q, k, v], dim=-1), wcat, bcat)
    dg = torch.cat([di, df], dim=-1).to(torch.bfloat16)
    def ref_bwd():
        x = torch.cat([q, k, v], dim=-1)
        dx = dg @ wcat
        dw = dg.flatten(0, 1).T @ x.flatten(0, 1)
        a, b_, c = dx.split(D, dim=-1)
        return dq + a, dk + b_, dv + c, dw
    r_f, r_b = timeit(ref_fwd), timeit(ref_bwd)
    bytes_f, bytes_b = T * 3 * D * 2, T * 9 * D * 2
    print(f"T={T} D={D}: fwd {t_f*1e3:.1f} us ({bytes_f/t_f/1e6:.0f} GB/s) vs cat+linear {r_f*1e3:.1f} us | "
          f"bwd {t_b*1e3:.1f} us ({bytes_b/t_b/1e6:.0f} GB/s) vs torch {r_b*1e3:.1f} us", flush=True)
