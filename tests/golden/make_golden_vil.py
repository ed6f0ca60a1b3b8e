"""Generates tests/golden/vil_pair_*.npz by RUNNING THE REFERENCE's ViL classes (this container only).

    python tests/golden/make_golden_vil.py        # needs /root/reference; writes next to this file

How the reference is made importable (nothing is copied from it):
  * /tmp/refpkg/ultralytics -> /root/reference symlink, matplotlib mocked;
  * a stand-in ``mlstm_kernels`` package whose mLSTMBackend.forward calls the reference's own
    ``chunkwise_simple`` (xlstm/blocks/mlstm/backends.py:149) with the config's chunk size / eps;
  * HEAD's MatrixLSTMCell.forward returns the raw (B,NH,S,DH) backend output because its last two lines
    are commented out (vision_lstm2.py:950-952), which makes ViLLayer.forward fail at :498.  The wrapper
    below applies exactly those two lines (``outnorm`` + head merge) to the reference's own output.
Recorded per case: the full state_dict, the input, and the outputs/gradients of
  TL block, BR block (reference formulation with the literal flip pair), and BR(TL(x)).
"""
import os
import sys
import types
from dataclasses import dataclass
from typing import Literal
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    for m in ["matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.backends",
              "matplotlib.backends.backend_agg"]:
        sys.modules[m] = MagicMock()
    os.makedirs("/tmp/refpkg", exist_ok=True)
    if not os.path.exists("/tmp/refpkg/ultralytics"):
        os.symlink(REF, "/tmp/refpkg/ultralytics")
    sys.path.insert(0, "/tmp/refpkg")
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ref_backends", f"{REF}/nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py")
    rb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rb)

    @dataclass
    class mLSTMBackendConfig:
        chunkwise_kernel: str = "chunkwise--native_autograd"
        sequence_kernel: str = "native_sequence__native"
        step_kernel: str = "native"
        mode: str = "train"
        chunk_size: int = 64
        return_last_states: bool = False
        autocast_kernel_dtype: str = "bfloat16"
        eps: float = 1e-6
        inference_state_dtype: str = "float32"

    class mLSTMBackend(torch.nn.Module):
        def __init__(self, config):
            super().__init__()
            self.config = config

        def forward(self, q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None,
                    return_last_states=None, mode=None):
            S, L = q.shape[2], self.config.chunk_size
            while S % L:          # chunkwise_simple needs L | S; its output is chunk-size invariant
                L -= 1
            return rb.chunkwise_simple(q, k, v, i, f, chunk_size=L, eps=self.config.eps)

    names = ["mlstm_kernels", "mlstm_kernels.torch", "mlstm_kernels.torch.backend_module",
             "mlstm_kernels.torch.chunkwise", "mlstm_kernels.torch.chunkwise.triton_xl_chunk"]
    mods = {n: types.ModuleType(n) for n in names}
    for n in names[:2] + names[3:4]:
        mods[n].__path__ = []
    bm = mods["mlstm_kernels.torch.backend_module"]
    bm.mLSTMBackendConfig, bm.mLSTMBackend = mLSTMBackendConfig, mLSTMBackend
    for n in ["ChunkwiseKernelType", "SequenceKernelType", "StepKernelType", "DtypeType", "BackendModeType"]:
        setattr(bm, n, Literal["x"])
    mods["mlstm_kernels.torch.chunkwise.triton_xl_chunk"].mlstm_chunkwise__xl_chunk = None
    sys.modules.update(mods)
    from ultralytics.nn.modules.vision_lstm import vision_lstm2 as V
    head_forward = V.MatrixLSTMCell.forward

    def intended_forward(self, q, k, v):          # vision_lstm2.py:950-952, un-commented
        B, S, _ = q.shape
        h = head_forward(self, q, k, v)
        return self.outnorm(h).transpose(1, 2).reshape(B, S, -1)

    V.MatrixLSTMCell.forward = intended_forward
    return V


def randomise(module, gen):
    """Non-degenerate parameters: the reference init zeroes the gate weights and the outnorm."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("igate.weight") or name.endswith("fgate.weight"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
            elif name.endswith("igate.bias"):
                p.copy_(torch.randn(p.shape, generator=gen))
            elif name.endswith("outnorm.weight") or name.endswith("outnorm.bias") or name.endswith("proj_down.bias") \
                    or name.endswith("conv.bias") or name.endswith("q_proj.bias") or name.endswith("k_proj.bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
            elif name.endswith("learnable_skip") or name.endswith("layer.norm.weight"):
                p.copy_(1.0 + torch.randn(p.shape, generator=gen) * 0.2)


def make_case(V, tag, dim, qkv_block_size, grid, batch, chunk_size, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    pair = V.ViLBlockPair(dim=dim, chunk_size=chunk_size, qkv_block_size=qkv_block_size).double()
    randomise(pair, gen)
    pair.train()
    S = grid[0] * grid[1]
    x = torch.randn(batch, S, dim, generator=gen, dtype=torch.float64)
    dy = torch.randn(batch, S, dim, generator=gen, dtype=torch.float64)
    out = {"x": x.numpy(), "dy": dy.numpy(), "dim": dim, "qkv_block_size": qkv_block_size, "chunk_size": chunk_size}
    for key, fn in {
        "tl": lambda t: pair.rowwise_from_top_left(t),
        "br": lambda t: pair.rowwise_from_bot_right(t),
        "pair": lambda t: pair.rowwise_from_bot_right(pair.rowwise_from_top_left(t)),
    }.items():
        xi = x.clone().requires_grad_(True)
        pair.zero_grad()
        y = fn(xi)
        y.backward(dy)
        out[f"y_{key}"] = y.detach().numpy()
        out[f"dx_{key}"] = xi.grad.numpy()
        if key == "pair":
            for n, p in pair.named_parameters():
                if p.grad is not None and ("mlstm_cell" in n or n.endswith("conv.weight")):
                    out["grad__" + n] = p.grad.numpy()
    for n, p in pair.state_dict().items():
        out["param__" + n] = p.numpy()
    path = os.path.join(HERE, f"vil_pair_{tag}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB", {k: float(np.abs(out[k]).max()) for k in ("y_tl", "y_br", "y_pair")})


if __name__ == "__main__":
    V = import_reference()
    make_case(V, "dim32_dh16_8x8", dim=32, qkv_block_size=16, grid=(8, 8), batch=2, chunk_size=16, seed=11)
    make_case(V, "dim64_dh64_6x6", dim=64, qkv_block_size=64, grid=(6, 6), batch=1, chunk_size=12, seed=12)
