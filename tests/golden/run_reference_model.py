"""BASELINE configs[0] in this container: the reference's own n-scale xLSTM-YOLO detector (authored YAML of
SURVEY.md Appendix A, parsed by the reference's unmodified nn/tasks.py) forwards one 640x640 image on CPU

  (ref) with the reference's MatrixLSTMCell + its PyTorch chunkwise_simple behind a stand-in mlstm_kernels,
  (B2)  with this repo's mLSTMBackend objects plugged into the reference cell (operator seam),
  (B1)  with this repo's MatrixLSTMCell replacing the reference class instance by instance (module level),
  (B0)  with this repo's ViLBlockPair (head_compat: TL block only, as HEAD runs it) replacing the reference pairs,

and prints the max relative deviations from (ref), forward (eval) and input-gradient (train).  Needs /root/reference."""
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden_vil as G  # noqa: E402

V = G.import_reference()
import torch  # noqa: E402
import yaml  # noqa: E402
from ultralytics.nn.tasks import DetectionModel  # noqa: E402

import xlstm_yolo_b200 as X  # noqa: E402

YAML = open(os.path.join(os.path.dirname(os.path.dirname(HERE)), "xlstm_yolo_b200", "compat", "yamls", "xlstm-yolo.yaml")).read()


def run(model, x):
    model.eval()
    with torch.no_grad():
        y = model(x)[0]
    model.train()
    xi = x.clone().requires_grad_(True)
    out = model(xi)
    feats = out if isinstance(out, (list, tuple)) else [out]
    loss = sum((f.float() ** 2).mean() for f in feats)
    loss.backward()
    return y, xi.grad


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def swap(model, pred, make):
    for name, mod in list(model.named_modules()):
        for cname, child in list(mod.named_children()):
            if pred(child):
                setattr(mod, cname, make(child))


cfg = yaml.safe_load(YAML)
cfg["scale"] = "n"
torch.manual_seed(0)
ref = DetectionModel(cfg, ch=3, nc=80, verbose=False)
with torch.no_grad():   # the reference init zeroes the gate weights and the out-norm: make them matter
    g = torch.Generator().manual_seed(1)
    for n, p in ref.named_parameters():
        if "mlstm_cell" in n:
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.dim() > 1 else 1.0) * (0.2 if "outnorm" in n else 1.0))
x = torch.rand(1, 3, 640, 640, generator=torch.Generator().manual_seed(2))
y_ref, dx_ref = run(ref, x)
is_cell = lambda m: type(m).__name__ == "MatrixLSTMCell" and not isinstance(m, X.MatrixLSTMCell)
n_cells = sum(1 for m in ref.modules() if is_cell(m))

# (B2) operator seam: this repo's backend objects inside the reference cell
b2 = copy.deepcopy(ref)
for m in b2.modules():
    if is_cell(m):
        for attr in ("cpu_backend", "cpu_backend_infer", "gpu_backend", "gpu_backend_infer"):
            old = getattr(m, attr)
            c = old.config
            setattr(m, attr, X.mLSTMBackend(X.mLSTMBackendConfig(
                chunkwise_kernel="chunkwise--native_autograd", sequence_kernel=c.sequence_kernel, step_kernel=c.step_kernel,
                mode=c.mode, chunk_size=c.chunk_size, return_last_states=c.return_last_states,
                autocast_kernel_dtype=c.autocast_kernel_dtype, eps=c.eps)))
y_b2, dx_b2 = run(b2, x)


# (B1) module level: our cell in place of each reference cell instance
def our_cell(old):
    new = X.MatrixLSTMCell(dim=old.dim, num_heads=old.num_heads)
    new.load_state_dict(old.state_dict(), strict=True)
    return new


b1 = copy.deepcopy(ref)
swap(b1, is_cell, our_cell)
y_b1, dx_b1 = run(b1, x)


# (B0) layer stack: our ViLBlockPair in place of each reference pair (HEAD runs the TL block only)
def our_pair(old):
    tl = old.rowwise_from_top_left
    new = X.ViLBlockPair(dim=tl.dim, chunk_size=16, qkv_block_size=tl.layer.qkv_block_size)
    new.load_state_dict(old.state_dict(), strict=True)
    new.head_compat = True
    return new


b0 = copy.deepcopy(ref)
swap(b0, lambda m: type(m).__name__ == "ViLBlockPair" and not isinstance(m, X.ViLBlockPair), our_pair)
y_b0, dx_b0 = run(b0, x)

print(json.dumps({
    "cells": n_cells, "out_shape": list(y_ref.shape), "y_absmax": float(y_ref.abs().max()),
    "B2_y": rel(y_b2, y_ref), "B2_dx": rel(dx_b2, dx_ref),
    "B1_y": rel(y_b1, y_ref), "B1_dx": rel(dx_b1, dx_ref),
    "B0_y": rel(y_b0, y_ref), "B0_dx": rel(dx_b0, dx_ref),
}))
