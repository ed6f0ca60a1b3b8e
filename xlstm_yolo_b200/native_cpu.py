"""Plain-PyTorch mLSTM for CPU tensors only.

The reference's cell keeps two backends and switches on the device
(``backend = self.gpu_backend if device.type == 'cuda' else self.cpu_backend``,
vision_lstm2.py:891-892; CPU = ``chunkwise--native_autograd``).  A CPU forward is needed at
model-construction time (the stride probe, nn/tasks.py:353-362).  This module is that CPU
branch.  It is NEVER used for CUDA tensors — ``ops.mlstm`` has no fallback — and is not the
oracle (tests compare the CUDA kernels with ``oracle/``, never with this file).
Algorithm: chunkwise form of backends.py:149-263 in the (u, M) variables of DESIGN.md,
any S (tail masked).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def mlstm_chunkwise_cpu(q, k, v, i, f, c_initial=None, n_initial=None, m_initial=None, return_last_states=False,
                        eps=1e-6, chunk_size=64, reverse=False, input_gate="exp"):
    if q.is_cuda:
        raise RuntimeError("native_cpu.mlstm_chunkwise_cpu is for CPU tensors only")
    if reverse:
        q, k, v, i, f = (t.flip(dims=[2]) for t in (q, k, v, i, f))
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    L = max(1, min(int(chunk_size), S))
    scale = 1.0 / math.sqrt(DK)
    C = q.new_zeros(B, NH, DK, DV) if c_initial is None else c_initial.to(q.dtype)
    n = q.new_zeros(B, NH, DK) if n_initial is None else n_initial.to(q.dtype)
    m = q.new_zeros(B, NH) if m_initial is None else m_initial.to(q.dtype).reshape(B, NH)
    logf = F.logsigmoid(f.to(q.dtype))
    i = i.to(q.dtype)
    sig = input_gate == "sigmoid"   # log-gate logsigmoid(i), no stabiliser: m == 0, normaliser max(|n|, 1)
    if sig:
        i = F.logsigmoid(i)
        m = torch.zeros_like(m)
    outs = []
    for a in range(0, S, L):
        e = min(S, a + L)
        qc, kc, vc = q[:, :, a:e] * scale, k[:, :, a:e], v[:, :, a:e]
        b = logf[:, :, a:e].cumsum(-1)
        u = i[:, :, a:e] - b
        M = -b if sig else torch.maximum(m[..., None], u.cummax(-1).values)
        causal = torch.ones(e - a, e - a, dtype=torch.bool, device=q.device).tril()
        D = torch.exp(u[..., None, :] - M[..., :, None]).masked_fill(~causal, 0.0)
        w = torch.exp(m[..., None] - M)
        E = (qc @ kc.transpose(-1, -2)) * D
        nrow = E.sum(-1) + w * (qc * n[..., None, :]).sum(-1)
        den = torch.maximum(nrow.abs(), torch.exp(-(b + M))) + eps
        outs.append((E @ vc + w[..., None] * (qc @ C)) / den[..., None])
        ML = M[..., -1]
        kbar = kc * torch.exp(u - ML[..., None])[..., None]
        decay = torch.exp(m - ML)
        C = decay[..., None, None] * C + kbar.transpose(-1, -2) @ vc
        n = decay[..., None] * n + kbar.sum(-2)
        m = b[..., -1] + ML
    h = torch.cat(outs, dim=2)
    if reverse:
        h = h.flip(dims=[2])
    if return_last_states:
        return h, (C, n, m.reshape(B, NH, 1))
    return h
