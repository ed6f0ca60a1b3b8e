"""Whole ViLBlockPair (TL then BR) forward + backward on the GPU: this repo's layer stack (CUDA cell, fused gate
projection, fused tail, flip-free reverse scan) vs the same modules with every fusion switched off and the literal
flip pair (what the reference's layer does around a cell).  bf16 autocast-free, synthetic input."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from xlstm_yolo_b200 import ViLBlockPair

def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

for (B, grid, dim, bs) in [(32, 20, 128, 64), (32, 40, 256, 128), (8, 80, 256, 128)]:
    S = grid * grid
    torch.manual_seed(0)
    pair = ViLBlockPair(dim=dim, chunk_size=64, qkv_block_size=bs).cuda().to(torch.bfloat16).train()
    plain = copy.deepcopy(pair)
    noprod = copy.deepcopy(pair)     # everything but the conv + SiLU + q / k / v producer kernel
    for blk in (noprod.rowwise_from_top_left, noprod.rowwise_from_bot_right):
        blk.layer.fused_producer = False
    for blk in (plain.rowwise_from_top_left, plain.rowwise_from_bot_right):
        blk.layer.fused_producer = False
        blk.layer.fused_tail = False
        blk.layer.flip_free = False
        blk.layer.mlstm_cell.fused_gates = False
    x = torch.randn(B, S, dim, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    dy = torch.randn(B, S, dim, device="cuda", dtype=torch.bfloat16)
    def run(m):
        def f():
            y = m(x); y.backward(dy)
        return f
    def fwd_only(m):
        def f():
            with torch.no_grad():
                m(x)
        return f
    tf, tn, tp = timeit(run(pair)), timeit(run(noprod)), timeit(run(plain))
    ff, fn_ = timeit(fwd_only(pair)), timeit(fwd_only(noprod))
    print(f"ViLBlockPair dim={dim} inner={2*dim} DH={bs} B={B} S={S}: fused {tf:.3f} ms ({B*S/tf/1e3:.1f} M tok/s)  "
          f"without the producer kernel {tn:.3f} ms  unfused+flips {tp:.3f} ms  speed-up {tp/tf:.2f}x | forward only: "
          f"{ff:.3f} ms vs {fn_:.3f} ms without the producer", flush=True)
