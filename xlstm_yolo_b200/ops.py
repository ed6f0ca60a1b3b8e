"""The mLSTM cell operator on CUDA tensors: a ``torch.autograd.Function`` over the C ABI.

Functional mirror of the reference's backend call::

    h[, (C_last, n_last, m_last)] = mlstm(q, k, v, i, f, c_initial=None, n_initial=None,
                                          m_initial=None, return_last_states=False, ...)

with q,k: (B,NH,S,DHqk), v: (B,NH,S,DHv) — typically strided views of (B,S,NH,DH) storage
(vision_lstm2.py:900-902) — and i,f: (B,NH,S).  PyTorch is plumbing only (allocation,
streams, autograd bookkeeping); all arithmetic happens in ``lib/libmlstm_b200.so``.
CUDA tensors only: there is no CPU or PyTorch fallback here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import Act, Gate, Params

_DT = {torch.float32: _lib.MLSTM_F32, torch.bfloat16: _lib.MLSTM_BF16}


def _fail(rc: int, what: str):
    raise RuntimeError(f"mlstm_b200 {what} failed: {_lib.STATUS.get(rc, rc)}: {_lib.last_error()}")


def _act_ok(t: torch.Tensor) -> bool:
    if t.stride(-1) != 1 and t.shape[-1] != 1:
        return False
    if t.dtype == torch.bfloat16:  # TMA: 16-byte aligned base and strides
        if t.data_ptr() % 16 or any(s % 8 for s in t.stride()[:-1]):
            return False
    return True


def _prep_act(t: torch.Tensor) -> torch.Tensor:
    """Keep the caller's strided layout when the kernels can read it, else copy into
    (B,S,NH,DH) storage (the reference's native layout)."""
    if _act_ok(t):
        return t
    B, NH, S, D = t.shape
    out = torch.empty((B, S, NH, D), dtype=t.dtype, device=t.device).transpose(1, 2)
    out.copy_(t)
    return out


def _empty_act(B, NH, S, D, dtype, device) -> torch.Tensor:
    return torch.empty((B, S, NH, D), dtype=dtype, device=device).transpose(1, 2)


def _act(t: Optional[torch.Tensor]) -> Act:
    if t is None:
        return Act(None, 0, 0, 0)
    return Act(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def _gate(t: Optional[torch.Tensor]) -> Gate:
    if t is None:
        return Gate(None, 0, 0, 0)
    return Gate(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _base_params(q, k, v, i, f, eps, chunk_size, reverse, qk_scale) -> Params:
    B, NH, S, DK = q.shape
    p = Params()
    p.abi_version = _lib.ABI_VERSION
    p.B, p.NH, p.S, p.DHQK, p.DHV = B, NH, S, DK, v.shape[-1]
    p.dtype = _DT[q.dtype]
    p.reverse = 1 if reverse else 0
    p.chunk_size = int(chunk_size)
    p.eps = float(eps)
    p.qk_scale = float(qk_scale or 0.0)
    p.q, p.k, p.v = _act(q), _act(k), _act(v)
    p.i, p.f = _gate(i), _gate(f)
    return p


def kernel_family(q: torch.Tensor, v: torch.Tensor) -> str:
    """'tcgen05' or 'simt' — which kernel family the library would pick for these shapes."""
    p = Params()
    p.abi_version = _lib.ABI_VERSION
    p.B, p.NH, p.S, p.DHQK, p.DHV = q.shape[0], q.shape[1], q.shape[2], q.shape[3], v.shape[3]
    p.dtype = _DT[q.dtype]
    name = _lib.load().mlstm_b200_kernel_name(C.byref(p), 0)
    return name.decode() if name else "none"


def _alloc_states(lib, p, dev):
    """Per-chunk entry-state buffer of the tcgen05 family (None for SIMT); attaches it to p."""
    need = lib.mlstm_b200_state_bytes(C.byref(p))
    if not need:
        return None
    states = torch.empty(need, dtype=torch.uint8, device=dev)
    p.states, p.states_bytes = states.data_ptr(), need
    return states


def mlstm_fwd_raw(q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None, *, eps=1e-6, chunk_size=64,
                  reverse=False, save_rows=True, return_last_states=False, qk_scale=None):
    """One forward launch through the C ABI.  Inputs must already be CUDA, fp32 or bf16
    (q,k,v same dtype), gates fp32.  Returns (h, n_row, m_row, last_states_or_None, chunk_states)."""
    lib = _lib.load()
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dev = q.device
    h = _empty_act(B, NH, S, DV, q.dtype, dev)
    n_row = m_row = None
    if save_rows:
        n_row = torch.empty((B, NH, S), dtype=torch.float32, device=dev)
        m_row = torch.empty((B, NH, S), dtype=torch.float32, device=dev)
    last = None
    if return_last_states:
        last = (torch.empty((B, NH, DK, DV), dtype=torch.float32, device=dev),
                torch.empty((B, NH, DK), dtype=torch.float32, device=dev),
                torch.empty((B, NH, 1), dtype=torch.float32, device=dev))
    p = _base_params(q, k, v, i, f, eps, chunk_size, reverse, qk_scale)
    p.c_initial, p.n_initial, p.m_initial = _ptr(c_initial), _ptr(n_initial), _ptr(m_initial)
    p.h = _act(h)
    p.n_row, p.m_row = _ptr(n_row), _ptr(m_row)
    if last is not None:
        p.c_last, p.n_last, p.m_last = (_ptr(t) for t in last)
    states = _alloc_states(lib, p, dev)
    with torch.cuda.device(dev):
        rc = lib.mlstm_b200_fwd(C.byref(p), _stream())
    if rc:
        _fail(rc, "forward")
    return h, n_row, m_row, last, states


def mlstm_bwd_raw(q, k, v, i, f, h, n_row, m_row, dh, c_initial=None, n_initial=None, m_initial=None, *, eps=1e-6,
                  chunk_size=64, reverse=False, qk_scale=None, states=None):
    """One backward call through the C ABI.  Returns (dq, dk, dv, di, df) — dq,dk,dv in the
    layout/dtype of q,k,v; di,df fp32 (B,NH,S)."""
    lib = _lib.load()
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dev = q.device
    dq = _empty_act(B, NH, S, DK, q.dtype, dev)
    dk = _empty_act(B, NH, S, DK, q.dtype, dev)
    dv = _empty_act(B, NH, S, DV, q.dtype, dev)
    di = torch.empty((B, S, NH), dtype=torch.float32, device=dev).transpose(1, 2)
    df = torch.empty((B, S, NH), dtype=torch.float32, device=dev).transpose(1, 2)
    p = _base_params(q, k, v, i, f, eps, chunk_size, reverse, qk_scale)
    p.c_initial, p.n_initial, p.m_initial = _ptr(c_initial), _ptr(n_initial), _ptr(m_initial)
    p.h = _act(h)
    p.n_row, p.m_row = _ptr(n_row), _ptr(m_row)
    p.dh = _act(dh)
    p.dq, p.dk, p.dv = _act(dq), _act(dk), _act(dv)
    p.di, p.df = _gate(di), _gate(df)
    if states is not None:
        p.states, p.states_bytes = states.data_ptr(), states.numel()
    need = lib.mlstm_b200_workspace_bytes(C.byref(p), 1)
    ws = None
    if need:
        ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dev)
        p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    with torch.cuda.device(dev):
        rc = lib.mlstm_b200_bwd(C.byref(p), _stream())
    if rc:
        _fail(rc, "backward")
    return dq, dk, dv, di, df


class _MLSTMCellFn(torch.autograd.Function):
    """autograd wrapper: saves q,k,v,i,f,h and the per-row (n, m); gates/decay matrices are
    recomputed per chunk in the backward kernels."""

    @staticmethod
    def forward(ctx, q, k, v, i, f, c_initial, n_initial, m_initial, eps, chunk_size, reverse, return_last_states):
        need_grad = any(t.requires_grad for t in (q, k, v, i, f))
        h, n_row, m_row, last, states = mlstm_fwd_raw(
            q, k, v, i, f, c_initial, n_initial, m_initial, eps=eps, chunk_size=chunk_size, reverse=reverse,
            save_rows=need_grad, return_last_states=return_last_states)
        if need_grad:
            ctx.save_for_backward(q, k, v, i, f, h, n_row, m_row, c_initial, n_initial, m_initial, states)
        ctx.cfg = (eps, chunk_size, reverse)
        if return_last_states:
            ctx.mark_non_differentiable(*last)
            return (h,) + tuple(last)
        return h

    @staticmethod
    def backward(ctx, dh, *dstates):
        q, k, v, i, f, h, n_row, m_row, c0, n0, m0, states = ctx.saved_tensors
        eps, chunk_size, reverse = ctx.cfg
        if dh.dtype != q.dtype:
            dh = dh.to(q.dtype)  # e.g. loss-scaled fp16 -> bf16: no clamping, inf/NaN propagate
        dh = _prep_act(dh)
        dq, dk, dv, di, df = mlstm_bwd_raw(q, k, v, i, f, h, n_row, m_row, dh, c0, n0, m0, eps=eps,
                                           chunk_size=chunk_size, reverse=reverse, states=states)
        return dq, dk, dv, di, df, None, None, None, None, None, None, None


def mlstm(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, i: torch.Tensor, f: torch.Tensor,
          c_initial: Optional[torch.Tensor] = None, n_initial: Optional[torch.Tensor] = None,
          m_initial: Optional[torch.Tensor] = None, return_last_states: bool = False, *, eps: float = 1e-6,
          chunk_size: int = 64, reverse: bool = False, kernel_dtype: Optional[torch.dtype] = None):
    """mLSTM cell on CUDA tensors.

    ``kernel_dtype``: torch.bfloat16 -> tcgen05 kernels (bf16 operands, fp32 accumulation);
    torch.float32 -> fp32 SIMT kernels.  Default: bf16 for bf16/fp16 inputs, fp32 for fp32
    inputs.  The result has the dtype of ``q``.  Initial/last states are fp32 and carry no
    gradient (the reference never differentiates through them).
    """
    if not q.is_cuda:
        raise RuntimeError("xlstm_yolo_b200.ops.mlstm needs CUDA tensors (no CPU fallback)")
    in_dtype = q.dtype
    if kernel_dtype is None:
        kernel_dtype = torch.float32 if in_dtype == torch.float32 else torch.bfloat16
    if kernel_dtype not in _DT:
        raise ValueError(f"kernel_dtype must be float32 or bfloat16, got {kernel_dtype}")
    q, k, v = (_prep_act(t.to(kernel_dtype)) for t in (q, k, v))
    i, f = i.to(torch.float32), f.to(torch.float32)
    c0 = None if c_initial is None else c_initial.detach().to(torch.float32).contiguous()
    n0 = None if n_initial is None else n_initial.detach().to(torch.float32).contiguous()
    m0 = None if m_initial is None else m_initial.detach().to(torch.float32).reshape(q.shape[0], q.shape[1]).contiguous()
    out = _MLSTMCellFn.apply(q, k, v, i, f, c0, n0, m0, float(eps), int(chunk_size), bool(reverse),
                             bool(return_last_states))
    if return_last_states:
        h, C, n, m = out
        return h.to(in_dtype), (C, n, m)
    return out.to(in_dtype)


class MLSTMPlan:
    """Pre-bound forward+backward call on fixed device buffers (steady-state / benchmark use).

    All outputs and the workspace are allocated once; ``forward()`` / ``backward()`` are then a
    single C-ABI call each (plus the host-side TMA descriptor encode inside the library).
    """

    def __init__(self, q, k, v, i, f, dh, *, eps=1e-6, chunk_size=64, reverse=False):
        self.lib = _lib.load()
        B, NH, S, DK = q.shape
        DV = v.shape[-1]
        dev = q.device
        self.inputs = (q, k, v, i, f, dh)
        self.h = _empty_act(B, NH, S, DV, q.dtype, dev)
        self.n_row = torch.empty((B, NH, S), dtype=torch.float32, device=dev)
        self.m_row = torch.empty((B, NH, S), dtype=torch.float32, device=dev)
        self.dq = _empty_act(B, NH, S, DK, q.dtype, dev)
        self.dk = _empty_act(B, NH, S, DK, q.dtype, dev)
        self.dv = _empty_act(B, NH, S, DV, q.dtype, dev)
        self.di = torch.empty((B, S, NH), dtype=torch.float32, device=dev).transpose(1, 2)
        self.df = torch.empty((B, S, NH), dtype=torch.float32, device=dev).transpose(1, 2)
        p = _base_params(q, k, v, i, f, eps, chunk_size, reverse, None)
        p.h = _act(self.h)
        p.n_row, p.m_row = _ptr(self.n_row), _ptr(self.m_row)
        p.dh = _act(dh)
        p.dq, p.dk, p.dv = _act(self.dq), _act(self.dk), _act(self.dv)
        p.di, p.df = _gate(self.di), _gate(self.df)
        self.states = _alloc_states(self.lib, p, dev)
        need = self.lib.mlstm_b200_workspace_bytes(C.byref(p), 1)
        self.ws = torch.empty(max(1, (need + 3) // 4), dtype=torch.float32, device=dev)
        p.workspace, p.workspace_bytes = self.ws.data_ptr(), self.ws.numel() * 4
        self.p = p
        self.family = self.lib.mlstm_b200_kernel_name(C.byref(p), 0).decode()
        self.variant_fwd = self.lib.mlstm_b200_kernel_variant(C.byref(p), 0).decode()
        self.variant_bwd = self.lib.mlstm_b200_kernel_variant(C.byref(p), 1).decode()

    def forward(self):
        rc = self.lib.mlstm_b200_fwd(C.byref(self.p), _stream())
        if rc:
            _fail(rc, "forward")
        return self.h

    def backward(self, part: int = -1):
        rc = self.lib.mlstm_b200_bwd_part(C.byref(self.p), part, _stream())
        if rc:
            _fail(rc, "backward")
        return self.dq, self.dk, self.dv, self.di, self.df
