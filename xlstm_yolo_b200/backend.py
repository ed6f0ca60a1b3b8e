"""``mLSTMBackendConfig`` / ``mLSTMBackend``: the operator seam the reference calls.

Mirrors ``mlstm_kernels.torch.backend_module`` exactly as the reference uses it
(vision_lstm2.py:819-877,912-948; mlstm_large.py:20-28,139-149,295-330;
xlstm/xlstm_large/model.py:3-14,477-487): same dataclass fields, same ``forward`` keyword
arguments and return convention, same ``Literal`` aliases.  CUDA tensors go to the
hand-written sm_100a kernels behind the C ABI (``ops.mlstm``); CPU tensors take the
reference's own CPU route, a native-PyTorch chunkwise form (``native_cpu``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional, Union

import torch
from torch import nn

ChunkwiseKernelType = Literal[
    "chunkwise--native_autograd", "chunkwise--native_custbw", "chunkwise--triton_limit_chunk",
    "chunkwise--triton_xl_chunk", "chunkwise--triton_xl_chunk_siging", "chunkwise--b200_tcgen05",
    "parallel--native_autograd", "parallel--native_custbw", "parallel--native_stablef_autograd",
    "parallel--native_stablef_custbw", "parallel--triton_limit_headdim",
]
SequenceKernelType = Literal["native_sequence__native", "native_sequence__triton"]
StepKernelType = Literal["native", "triton"]
DtypeType = Literal["float32", "bfloat16", "float16"]
BackendModeType = Literal["train", "train_with_padding", "inference"]

_TORCH_DTYPE = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.bfloat16}
# float16 maps to the bf16 tensor-core kernels: same operand width, fp32 accumulation and
# fp32 state, without fp16's overflow hazard for the un-normalised C state.


@dataclass
class mLSTMBackendConfig:
    chunkwise_kernel: str = "chunkwise--b200_tcgen05"
    sequence_kernel: str = "native_sequence__native"
    step_kernel: str = "native"
    mode: str = "train"
    chunk_size: int = 64
    return_last_states: bool = False
    autocast_kernel_dtype: str = "bfloat16"
    eps: float = 1e-6
    inference_state_dtype: str = "float32"

    def __post_init__(self):
        if self.mode not in ("train", "train_with_padding", "inference"):
            raise ValueError(f"unknown mode {self.mode!r}")
        if self.autocast_kernel_dtype not in _TORCH_DTYPE:
            raise ValueError(f"unknown autocast_kernel_dtype {self.autocast_kernel_dtype!r}")

    @property
    def input_gate(self) -> str:
        """Gate arithmetic the kernel string selects, on every device: "sigmoid" for upstream's sigmoid-input-gate
        kernels ("...xl_chunk_siging", what HEAD configures for CUDA tensors, vision_lstm2.py:835,866), else "exp"
        (the exponential gate with max-stabiliser of the in-tree PyTorch mLSTM, backends.py:149-263, which is what
        HEAD's CPU strings "chunkwise--native_autograd" name, vision_lstm2.py:819-828,850-859)."""
        return "sigmoid" if "siging" in str(self.chunkwise_kernel) else "exp"


class mLSTMBackend(nn.Module):
    """Callable backend: ``h`` of shape (B,NH,S,DHv) [+ (C_last, n_last, m_last)]."""

    config_class = mLSTMBackendConfig

    def __init__(self, config: mLSTMBackendConfig):
        super().__init__()
        self.config = config

    def forward(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, i: torch.Tensor, f: torch.Tensor,
                c_initial: Optional[torch.Tensor] = None, n_initial: Optional[torch.Tensor] = None,
                m_initial: Optional[torch.Tensor] = None, return_last_states: Optional[bool] = None,
                mode: Optional[str] = None, reverse: bool = False):
        cfg = self.config
        if return_last_states is None:
            return_last_states = cfg.return_last_states
        mode = cfg.mode if mode is None else mode  # train / inference: same arithmetic here
        if mode not in ("train", "train_with_padding", "inference"):
            raise ValueError(f"unknown mode {mode!r}")
        if q.dim() != 4 or k.shape != q.shape or v.shape[:3] != q.shape[:3]:
            raise ValueError(f"q,k,v must be (B,NH,S,DH); got {tuple(q.shape)}, {tuple(k.shape)}, {tuple(v.shape)}")
        if i.shape != q.shape[:3] or f.shape != q.shape[:3]:
            raise ValueError(f"i,f must be (B,NH,S); got {tuple(i.shape)}, {tuple(f.shape)}")
        if q.is_cuda:
            from . import ops
            return ops.mlstm(q, k, v, i, f, c_initial, n_initial, m_initial, return_last_states,
                             eps=cfg.eps, chunk_size=cfg.chunk_size, reverse=reverse,
                             kernel_dtype=_TORCH_DTYPE[cfg.autocast_kernel_dtype], input_gate=self.input_gate)
        from .native_cpu import mlstm_chunkwise_cpu
        in_dtype = q.dtype
        cdt = torch.float32 if in_dtype in (torch.float16, torch.bfloat16) else in_dtype
        out = mlstm_chunkwise_cpu(q.to(cdt), k.to(cdt), v.to(cdt), i.to(cdt), f.to(cdt), c_initial, n_initial,
                                  m_initial, return_last_states, eps=cfg.eps, chunk_size=cfg.chunk_size,
                                  reverse=reverse, input_gate=self.input_gate)
        if return_last_states:
            return out[0].to(in_dtype), out[1]
        return out.to(in_dtype)

    def fused_cell(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, igate: nn.Linear, fgate: nn.Linear,
                   num_heads: int, reverse: bool = False) -> torch.Tensor:
        """CUDA only: gate projection (vision_lstm2.py:895-897) + cell in one autograd node.
        q,k,v (B,S,dim) -> h (B,NH,S,DH).  Not part of the reference seam; ``MatrixLSTMCell`` uses it."""
        cfg = self.config
        from . import ops
        return ops.fused_cell(q, k, v, igate.weight, igate.bias, fgate.weight, fgate.bias, num_heads, eps=cfg.eps,
                              chunk_size=cfg.chunk_size, reverse=reverse,
                              kernel_dtype=_TORCH_DTYPE[cfg.autocast_kernel_dtype], input_gate=self.input_gate)

    @property
    def input_gate(self) -> str:
        return self.config.input_gate

    def extra_repr(self) -> str:
        return f"{self.config}"
