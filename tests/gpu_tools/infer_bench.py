"""BASELINE configs[4] regime: forward-only (inference) cell throughput at batch 16, d=512 (4 heads of 128),
S = 25600 / 6400 / 1600 tokens (1280x1280 input: P3 / P4 / P5 maps), both scan directions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from xlstm_yolo_b200 import ops

def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

B, NH, DH = 16, 4, 128
for S in (1600, 6400, 25600):
    q, k, v, i, f, _ = bench.as_heads(bench.make_inputs(torch, B, NH, S, DH, 0, "cuda", torch.bfloat16))
    with torch.no_grad():
        for rev in (False, True):
            t = timeit(lambda: ops.mlstm(q, k, v, i, f, eps=1e-6, reverse=rev))
            alg = (8 * DH + 8) * B * NH * S          # read q,k,v + i,f ; write h
            print(f"B={B} S={S} reverse={int(rev)}: {t:.3f} ms  {B*S/t/1e3:.1f} M tok/s  {alg/t/1e6:.0f} GB/s algorithmic "
                  f"({alg/t/1e6/6537:.2f} of HBM peak)", flush=True)
