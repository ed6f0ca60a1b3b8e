"""torch.profiler table of one ViLBlockPair forward+backward (which ops dominate the layer around the cell)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
from xlstm_yolo_b200 import ViLBlockPair
B, grid, dim, bs = 32, 40, 256, 128
S = grid * grid
pair = ViLBlockPair(dim=dim, chunk_size=64, qkv_block_size=bs).cuda().to(torch.bfloat16).train()
x = torch.randn(B, S, dim, device="cuda", dtype=torch.bfloat16, requires_grad=True)
dy = torch.randn(B, S, dim, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    pair(x).backward(dy)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        pair(x).backward(dy)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3 / 1e3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"total device time per iteration {tot:.3f} ms")
for k, t, c in rows[:28]:
    print(f"{t:8.3f} ms  x{c:3d}  {k[:110]}")
