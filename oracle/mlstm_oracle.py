"""CPU oracle for the mLSTM cell hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU with plain torch ops, the algorithm of the reference's
own in-tree PyTorch mLSTM (DJT777/xlstm-yolo, paths relative to the reference root):

  * ``nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py:149-263``  chunkwise form
  * ``nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py:93-146``   recurrent step
  * ``nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py:9-90``     parallel form
  * ``nn/modules/vision_lstm/vision_lstm2.py:882-956`` (+ intended epilogue ``:950-952``,
    norm ``:1262-1325``)                                              cell wrapper

Nothing in the product path (``xlstm_yolo_b200``) imports this module: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do,
and there only as the checker / the timed CPU baseline.

Parity pin: the reference ships no tests and no golden vectors for this path
(SURVEY.md §8c).  The oracle is pinned instead against outputs of the reference's own
functions run in the build container (``tests/golden/make_golden.py`` imports
``backends.py`` by file path and commits the vectors); ``tests/test_oracle.py`` replays
them.

Notation (per batch element and head; ``t`` = query row, ``j`` = key row)::

    logf_s = logsigmoid(f_s)                      b_t = sum_{s<=t} logf_s   (cumsum)
    u_j    = i_j - b_j                            M_t = max(m_0, max_{j<=t} u_j)
    m_t    = b_t + M_t                            (== the recurrent stabiliser state)
    D_tj   = exp(u_j - M_t)  for j <= t           w_t = exp(m_0 - M_t)
    E_tj   = (q_t . k_j / sqrt(DH)) * D_tj
    n_t    = sum_j E_tj + w_t * (q_t . n_0) / sqrt(DH)
    h_t    = (sum_j E_tj v_j + w_t * (q_t C_0) / sqrt(DH)) / (max(|n_t|, exp(-m_t)) + eps)

which is algebraically what ``backends.py:220-263`` computes chunk by chunk
(``stab`` there is ``m_t``; ``backends.py:233``), with the state carry of
``backends.py:196-218``: ``m' = g + M_L``, ``C' = exp(m_0 - M_L) C_0 + sum_j exp(u_j - M_L) k_j v_j^T``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
States = Tuple[Tensor, Tensor, Tensor]


def _flip_seq(*ts):
    return tuple(None if t is None else t.flip(dims=[2]) for t in ts)


def mlstm_recurrent(
    q: Tensor, k: Tensor, v: Tensor, i: Tensor, f: Tensor,
    c_initial: Optional[Tensor] = None, n_initial: Optional[Tensor] = None,
    m_initial: Optional[Tensor] = None, eps: float = 1e-6, reverse: bool = False,
    return_last_states: bool = False,
):
    """Definitional step-by-step form (follows backends.py:93-146, one call per token).

    q,k: (B,NH,S,DHqk)  v: (B,NH,S,DHv)  i,f: (B,NH,S).  State convention follows the
    *chunkwise* reference (backends.py:168: the 1/sqrt(DH) scale sits on q, so C and n
    hold unscaled keys).
    """
    if reverse:
        q, k, v, i, f = _flip_seq(q, k, v, i, f)
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dt = q.dtype
    C = torch.zeros(B, NH, DK, DV, dtype=dt) if c_initial is None else c_initial.to(dt).clone()
    n = torch.zeros(B, NH, DK, dtype=dt) if n_initial is None else n_initial.to(dt).clone()
    m = torch.zeros(B, NH, dtype=dt) if m_initial is None else m_initial.to(dt).reshape(B, NH).clone()
    scale = 1.0 / math.sqrt(DK)
    logf = F.logsigmoid(f)
    hs = []
    for t in range(S):
        m_new = torch.maximum(logf[:, :, t] + m, i[:, :, t])           # backends.py:129
        fg = torch.exp(logf[:, :, t] + m - m_new)                       # :131
        ig = torch.exp(i[:, :, t] - m_new)                              # :132
        kt, vt, qt = k[:, :, t], v[:, :, t], q[:, :, t] * scale
        C = fg[..., None, None] * C + ig[..., None, None] * (kt[..., :, None] * vt[..., None, :])  # :136
        n = fg[..., None] * n + ig[..., None] * kt                      # :137
        num = torch.einsum("bhd,bhde->bhe", qt, C)                      # :139
        qn = (qt * n).sum(-1)                                           # :141
        den = torch.maximum(qn.abs(), torch.exp(-m_new)) + eps          # :142-143
        hs.append(num / den[..., None])
        m = m_new
    h = torch.stack(hs, dim=2)
    if reverse:
        h = h.flip(dims=[2])
    if return_last_states:
        return h, (C, n, m.reshape(B, NH, 1))
    return h


def mlstm_parallel(q: Tensor, k: Tensor, v: Tensor, i: Tensor, f: Tensor,
                   eps: float = 1e-6, reverse: bool = False) -> Tensor:
    """Quadratic stabilised form (follows backends.py:9-90): row-max stabiliser only,
    i.e. *without* the zero-initial-state ``m_0 = 0`` entering the max."""
    if reverse:
        q, k, v, i, f = _flip_seq(q, k, v, i, f)
    B, NH, S, DK = q.shape
    b = F.logsigmoid(f).cumsum(-1)                                     # backends.py:42-55
    u = i - b
    logD = b[..., :, None] + u[..., None, :]                            # :59-68
    causal = torch.ones(S, S, dtype=torch.bool).tril()
    logD = logD.masked_fill(~causal, float("-inf"))
    m = logD.max(dim=-1, keepdim=True).values                          # :71
    E = (q @ k.transpose(-1, -2)) / math.sqrt(DK) * torch.exp(logD - m)  # :75-82
    den = torch.maximum(E.sum(-1, keepdim=True).abs(), torch.exp(-m)) + eps  # :83-85
    h = (E / den) @ v                                                  # :88
    return h.flip(dims=[2]) if reverse else h


def mlstm_chunkwise(
    q: Tensor, k: Tensor, v: Tensor, i: Tensor, f: Tensor,
    c_initial: Optional[Tensor] = None, n_initial: Optional[Tensor] = None,
    m_initial: Optional[Tensor] = None, chunk_size: int = 64, eps: float = 1e-6,
    return_last_states: bool = False, reverse: bool = False,
    kink_side: Optional[Tensor] = None, return_rows: bool = False,
):
    """Chunkwise-parallel form with inter-chunk (C, n, m) carry (backends.py:149-263).

    ``return_rows`` also returns the per-token normaliser sum ``n_t`` and stabiliser ``m_t`` (B,NH,S, token order).
    ``kink_side`` (bool (B,NH,S), token order; True = the ``|n_t|`` side) replaces the ``max(|n_t|, exp(-m_t))`` of
    backends.py:249-252 by the named side of it.  With ``kink_side = (|n_t| >= exp(-m_t))`` nothing changes; the
    tests use it to evaluate the one-sided derivative on the side a bf16 kernel landed on for the few rows that sit
    within rounding distance of the kink, where ``max`` is not differentiable and the reference's gradient is
    whichever side its own rounding picked.

    Differences from the reference function, none of which change its results where it
    is defined: any ``S`` is accepted (the last chunk is masked; the reference needs
    ``S % chunk_size == 0``, backends.py:164-170), ``reverse=True`` runs the scan from the
    last token (equivalent to flip -> cell -> flip, vision_lstm2.py:479-480,505-506), and
    DHqk may differ from DHv.
    """
    if reverse:
        q, k, v, i, f = _flip_seq(q, k, v, i, f)
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dt = q.dtype
    L = int(chunk_size)
    NC = (S + L - 1) // L
    pad = NC * L - S
    if pad:
        q = F.pad(q, (0, 0, 0, pad))
        k = F.pad(k, (0, 0, 0, pad))
        v = F.pad(v, (0, 0, 0, pad))
        i = F.pad(i, (0, pad), value=float("-inf"))     # padded keys get zero weight
        f = F.pad(f, (0, pad), value=float("inf"))      # logsigmoid(+inf) = 0: no decay
    scale = 1.0 / math.sqrt(DK)
    qc = q.reshape(B, NH, NC, L, DK)
    kc = k.reshape(B, NH, NC, L, DK)
    vc = v.reshape(B, NH, NC, L, DV)
    b = F.logsigmoid(f).reshape(B, NH, NC, L).cumsum(-1)               # backends.py:173-174
    g = b[..., -1]                                                     # chunk decay (log)
    u = i.reshape(B, NH, NC, L) - b                                    # :177
    ucm = u.cummax(dim=-1).values                                      # running max of u

    # ---- sequential part: scalar stabiliser, then (C, n) (backends.py:196-218) --------
    m_list = [torch.zeros(B, NH, dtype=dt) if m_initial is None else m_initial.to(dt).reshape(B, NH)]
    for c in range(NC):
        m_list.append(g[:, :, c] + torch.maximum(m_list[-1], ucm[:, :, c, -1]))
    m_prev = torch.stack(m_list, dim=2)                                # (B,NH,NC+1)
    M_end = torch.maximum(m_prev[:, :, :-1], ucm[..., -1])             # (B,NH,NC)
    kw = torch.exp(u - M_end[..., None])                               # key weights :181
    decay = torch.exp(m_prev[:, :, :-1] - M_end)                       # :203-207
    kv = kc.transpose(-1, -2) @ (vc * kw[..., None])                   # :183
    ks = (kc * kw[..., None]).sum(-2)                                  # :184
    C_list = [torch.zeros(B, NH, DK, DV, dtype=dt) if c_initial is None else c_initial.to(dt)]
    n_list = [torch.zeros(B, NH, DK, dtype=dt) if n_initial is None else n_initial.to(dt)]
    for c in range(NC):
        C_list.append(decay[:, :, c, None, None] * C_list[-1] + kv[:, :, c])
        n_list.append(decay[:, :, c, None] * n_list[-1] + ks[:, :, c])
    Cs = torch.stack(C_list, dim=2)                                    # (B,NH,NC+1,DK,DV)
    ns = torch.stack(n_list, dim=2)

    # ---- parallel part (backends.py:220-263) ------------------------------------------
    M = torch.maximum(m_prev[:, :, :-1, None], ucm)                    # (B,NH,NC,L); m_t = b + M  (:233)
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    D = torch.exp(u[..., None, :] - M[..., :, None]).masked_fill(~causal, 0.0)   # :242-243
    w = torch.exp(m_prev[:, :, :-1, None] - M)                         # :235
    E = (qc @ kc.transpose(-1, -2)) * scale * D                        # :246-247
    n_row = E.sum(-1) + w * scale * (qc * ns[:, :, :-1, None, :]).sum(-1)        # :237-239,250
    floor_ = torch.exp(-(b + M))
    if kink_side is None:
        den = torch.maximum(n_row.abs(), floor_) + eps                 # :249-254
    else:
        side = kink_side.flip(dims=[2]) if reverse else kink_side
        side = F.pad(side, (0, pad), value=True).reshape(B, NH, NC, L)
        den = torch.where(side, n_row.abs(), floor_) + eps
    num = E @ vc + (w * scale)[..., None] * (qc @ Cs[:, :, :-1])       # :234-236,257
    h = (num / den[..., None]).reshape(B, NH, NC * L, DV)[:, :, :S]
    rows = None
    if return_rows:
        rows = (n_row.reshape(B, NH, NC * L)[:, :, :S].detach(), (b + M).reshape(B, NH, NC * L)[:, :, :S].detach())
        if reverse:
            rows = _flip_seq(*rows)
    if reverse:
        h = h.flip(dims=[2])
    out = (h,)
    if return_last_states:
        out = out + ((Cs[:, :, -1], ns[:, :, -1], m_prev[:, :, -1].reshape(B, NH, 1)),)
    if return_rows:
        out = out + (rows,)
    return out if len(out) > 1 else h


def multihead_layernorm(h: Tensor, weight: Optional[Tensor], bias: Optional[Tensor],
                        eps: float = 1e-3) -> Tensor:
    """Per-head layer norm over DH with ``1 + weight`` (vision_lstm2.py:1281-1325).
    h: (B,NH,S,DH) -> (B,NH,S,DH)."""
    B, NH, S, DH = h.shape
    mu = h.mean(-1, keepdim=True)
    var = h.var(-1, unbiased=False, keepdim=True)
    y = (h - mu) / torch.sqrt(var + eps)
    if weight is not None:
        y = y * (1.0 + weight).reshape(1, NH, 1, DH)
    if bias is not None:
        y = y + bias.reshape(1, NH, 1, DH)
    return y


def cell_forward(
    q: Tensor, k: Tensor, v: Tensor, num_heads: int,
    igate_w: Tensor, igate_b: Tensor, fgate_w: Tensor, fgate_b: Tensor,
    outnorm_w: Optional[Tensor], outnorm_b: Optional[Tensor],
    chunk_size: int = 64, eps: float = 5e-5, norm_eps: float = 1e-3, reverse: bool = False,
    form: str = "chunkwise",
) -> Tensor:
    """``MatrixLSTMCell.forward`` with the intended epilogue.

    q,k,v: (B,S,inner).  Gates from ``cat[q,k,v]`` (vision_lstm2.py:895-897), heads split
    (:900-902), chunkwise backend with eps=5e-5 (:827), then ``outnorm`` and head merge
    (:950-952, commented out at HEAD; cell.py:70-71 in the vendored package).
    """
    B, S, H = q.shape
    x = torch.cat([q, k, v], dim=-1)
    ig = (x @ igate_w.T + igate_b).transpose(-1, -2)                   # (B,NH,S)
    fg = (x @ fgate_w.T + fgate_b).transpose(-1, -2)
    qh = q.reshape(B, S, num_heads, -1).transpose(1, 2)
    kh = k.reshape(B, S, num_heads, -1).transpose(1, 2)
    vh = v.reshape(B, S, num_heads, -1).transpose(1, 2)
    if form == "parallel":      # what the vendored cell uses (cell.py:27)
        h = mlstm_parallel(qh, kh, vh, ig, fg, eps=eps, reverse=reverse)
    else:                       # what MatrixLSTMCell selects on CPU (vision_lstm2.py:851)
        h = mlstm_chunkwise(qh, kh, vh, ig, fg, chunk_size=chunk_size, eps=eps, reverse=reverse)
    h = multihead_layernorm(h, outnorm_w, outnorm_b, eps=norm_eps)
    return h.transpose(1, 2).reshape(B, S, H)


# ---------------------------------------------------------------------------------------------
# Sigmoid-input-gate variant ("siging").  PARITY UNPINNED: HEAD asks for it on CUDA through the kernel
# string "chunkwise--triton_xl_chunk_siging" (vision_lstm2.py:835,866), but the arithmetic lives in the
# external, un-vendored, un-versioned ``mlstm_kernels`` package and nothing in the
# reference tree restates or tests it.  What follows is its published algorithm (the package's native
# parallel form): log input gate logsigmoid(i), log forget gate logsigmoid(f), no max-stabiliser
# (every log-weight is <= 0), normaliser max(|n_t|, 1) + eps:
#     D_tj = exp(logsig(i_j) + sum_{j<s<=t} logsig(f_s))   (j <= t)
#     h_t  = sum_j D_tj (q_t.k_j / sqrt(DH)) v_j / (max(|sum_j D_tj q_t.k_j / sqrt(DH)|, 1) + eps)
# The recurrent form below is the same thing step by step and carries (C, n) states (m == 0).
# ---------------------------------------------------------------------------------------------
def mlstm_siging_parallel(q: Tensor, k: Tensor, v: Tensor, i: Tensor, f: Tensor, eps: float = 1e-6,
                          reverse: bool = False, kink_side: Optional[Tensor] = None, return_rows: bool = False):
    """``kink_side`` / ``return_rows``: as in ``mlstm_chunkwise`` (the floor of this normaliser is 1, m == 0)."""
    if reverse:
        q, k, v, i, f = _flip_seq(q, k, v, i, f)
    S, DH = q.shape[2], q.shape[3]
    b = F.logsigmoid(f).cumsum(-1)
    logD = b[..., :, None] - b[..., None, :] + F.logsigmoid(i)[..., None, :]
    causal = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()
    D = torch.exp(logD.masked_fill(~causal, -float("inf")))
    E = (q @ k.transpose(-1, -2)) * (DH ** -0.5) * D
    n_row = E.sum(-1, keepdim=True)
    one = torch.ones((), dtype=q.dtype, device=q.device)
    if kink_side is None:
        n = torch.maximum(n_row.abs(), one)
    else:
        side = kink_side.flip(dims=[2]) if reverse else kink_side
        n = torch.where(side[..., None], n_row.abs(), one)
    h = (E / (n + eps)) @ v
    h = h.flip(dims=[2]) if reverse else h
    if return_rows:
        nr = n_row.squeeze(-1).detach()
        nr = nr.flip(dims=[2]) if reverse else nr
        return h, (nr, torch.zeros_like(nr))
    return h


def mlstm_siging_recurrent(q: Tensor, k: Tensor, v: Tensor, i: Tensor, f: Tensor, c_initial: Optional[Tensor] = None,
                           n_initial: Optional[Tensor] = None, eps: float = 1e-6, return_last_states: bool = False,
                           reverse: bool = False):
    if reverse:
        q, k, v, i, f = _flip_seq(q, k, v, i, f)
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    C = q.new_zeros(B, NH, DK, DV) if c_initial is None else c_initial.to(q.dtype)
    n = q.new_zeros(B, NH, DK) if n_initial is None else n_initial.to(q.dtype)
    fg, ig = torch.sigmoid(f), torch.sigmoid(i)
    outs = []
    for t in range(S):
        kt, vt, qt = k[:, :, t], v[:, :, t], q[:, :, t] * (DK ** -0.5)
        C = fg[:, :, t, None, None] * C + ig[:, :, t, None, None] * kt[..., :, None] * vt[..., None, :]
        n = fg[:, :, t, None] * n + ig[:, :, t, None] * kt
        den = torch.maximum((qt * n).sum(-1).abs(), torch.ones((), dtype=q.dtype, device=q.device)) + eps
        outs.append((qt[..., None, :] @ C).squeeze(-2) / den[..., None])
    h = torch.stack(outs, dim=2)
    if reverse:
        h = h.flip(dims=[2])
    if return_last_states:
        return h, (C, n, q.new_zeros(B, NH, 1))
    return h


def mlstm_fwbw(q, k, v, i, f, dh, chunk_size=64, eps=1e-6, reverse=False, input_gate="exp", kink_side=None, **kw):
    """Forward + autograd backward through the chunkwise oracle.  Returns
    (h, dq, dk, dv, di, df).  Used as the gradient oracle and as the CPU baseline step.
    ``kink_side``: see ``mlstm_chunkwise`` (h is always the un-overridden forward)."""
    leaves = [t.detach().clone().requires_grad_(True) for t in (q, k, v, i, f)]

    def fwd(side):
        if input_gate == "sigmoid":
            if kw.get("c_initial") is not None or kw.get("n_initial") is not None:
                assert side is None, "kink_side is not available in the recurrent sigmoid-gate form"
                return mlstm_siging_recurrent(*leaves, c_initial=kw.get("c_initial"), n_initial=kw.get("n_initial"),
                                              eps=eps, reverse=reverse)
            return mlstm_siging_parallel(*leaves, eps=eps, reverse=reverse, kink_side=side)
        return mlstm_chunkwise(*leaves, chunk_size=chunk_size, eps=eps, reverse=reverse, kink_side=side, **kw)

    h = fwd(kink_side)
    h.backward(dh)
    h_out = h.detach()
    if kink_side is not None:
        with torch.no_grad():
            h_out = fwd(None)
    return (h_out,) + tuple(t.grad for t in leaves)


def kink_rows(n_row: Tensor, m_row: Tensor, width: float = 2e-2):
    """(side, near): side = |n_t| >= exp(-m_t) (the branch max() takes), near = |n_t| within ``width`` (relative) of
    the floor exp(-m_t), i.e. rows whose side a bf16-level perturbation of n_t can flip."""
    floor_ = torch.exp(-m_row)
    return n_row.abs() >= floor_, ((n_row.abs() - floor_).abs() / floor_) < width
