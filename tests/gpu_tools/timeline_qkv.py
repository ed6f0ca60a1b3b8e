"""Developer tool: phase timeline of the operand producer.  Build the stamped variant and select it:
    python -c "from xlstm_yolo_b200 import build; build.build_variant('qtl', ('QKV_TIMELINE',), ('mlstm_qkv.cu',))"
    MLSTM_B200_LIB=$PWD/xlstm_yolo_b200/lib/libmlstm_b200_qtl.so python tests/gpu_tools/timeline_qkv.py
CTA 0 dumps clock64() stamps of its first 8 tiles over the first bytes of v."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from xlstm_yolo_b200 import ops
B, gh, gw, NH, d = 32, 40, 40, 4, 128
D = NH * d
up = torch.randn(B, gh * gw, 2 * D, device="cuda").bfloat16()
w = lambda *s: torch.randn(*s, device="cuda") * 0.05
args = (up[..., :D], w(D, 1, 3, 3), w(D), w(NH, d, d), w(D), w(NH, d, d), w(D), w(NH, d, d), w(D), gh, gw, False)
for _ in range(3):
    c, q, k, v = ops.qkv_producer(*args)
torch.cuda.synchronize()
st = v.view(-1).view(torch.int64)[:128].cpu().tolist()
names = ["top", "x landed", "conv done", "sync", "mma issued", "mma done", "epilogue done"]
t00 = st[0]
for thr, off in (("compute0", 0), ("control ", 64)):
    for n in range(8):
        row = st[off + n * 8: off + n * 8 + 7]
        print(thr, "tile", n, f"start=+{(row[0] - t00) / 1.965e3:7.2f}us", " ".join(f"{names[k]}=+{(row[k] - row[0]) / 1.965e3:.2f}" for k in range(1, 7)))
