"""Fused layer tail vs the reference's separate ops (group_norm, skip add, SiLU gate) on the GPU."""
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

for (B, S, NH, DH) in [(32, 400, 4, 64), (32, 1600, 4, 128), (32, 6400, 4, 128)]:
    D, T = NH * DH, B * S
    h = torch.randn(B, S, NH, DH, device="cuda", dtype=torch.bfloat16).transpose(1, 2).requires_grad_(True)
    c = torch.randn(B, S, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    up = torch.randn(B, S, 2 * D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = torch.zeros(D, device="cuda", requires_grad=True); b = torch.zeros(D, device="cuda", requires_grad=True)
    skip = torch.ones(D, device="cuda", requires_grad=True)
    dy = torch.randn(B, S, D, device="cuda", dtype=torch.bfloat16)
    def fused():
        y = ops.layer_tail(h, c, up[..., D:], w, b, skip, 1e-3); y.backward(dy)
    def plain():
        x = h.transpose(1, 2).reshape(T, D)
        n = F.group_norm(x, NH, (1 + w).to(x.dtype), b.to(x.dtype), 1e-3).view(B, S, D)
        y = (n + skip.to(x.dtype) * c) * F.silu(up[..., D:]); y.backward(dy)
    tf, tp = timeit(fused), timeit(plain)
    alg = T * D * 2 * (4 + 7)      # fwd: read h,c,z write y ; bwd: read dy,h,c,z write dh,dc,dz
    print(f"T={T} D={D}: fused fwd+bwd {tf*1e3:.1f} us ({alg/tf/1e6:.0f} GB/s algorithmic) vs torch ops {tp*1e3:.1f} us", flush=True)
