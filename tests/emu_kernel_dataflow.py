"""CPU emulation (torch, any dtype) of the chunk-level dataflow the CUDA kernels use.

Test infrastructure: it restates, chunk by chunk and in the kernels' own variables, what
``csrc/mlstm_tc_fwd.cu`` / ``mlstm_tc_bwd.cu`` (and the SIMT kernels) compute, so that the
hand-derived backward can be checked against autograd through the oracle without a GPU.

Forward (per chunk c, local cumsum b, u = i - b, M_t = max(m_prev, cummax u), m_t = b_t + M_t):
    D_tj = exp(u_j - M_t) [j<=t];  w_t = exp(m_prev - M_t);  E = s * (Q K^T) * D
    n_t = sum_j E_tj + w_t s (q_t . n_prev);  N_t = max(|n_t|, exp(-m_t)) + eps
    h_t = (E V + w_t s (Q C_prev))_t / N_t
    state: kw_j = exp(u_j - M_L); decay = exp(m_prev - M_L); C = decay C_prev + (kw K)^T V; m = g + M_L
Backward (stabiliser m treated as a constant, as the upstream Triton kernels do):
  kernel A, forward walk (recomputes C_prev like the forward):
    Z = dH V^T;  G = dH C_prev^T;  delta_t = sum_j E_tj Z_tj + w_t s (q_t . G_t)
    dn_t = -[|n_t| >= exp(-m_t)] sign(n_t) delta_t / N_t^2
    dS = s (Z / N_t + dn_t) * D;  dq_t = dS K + w_t s (G_t / N_t + dn_t n_prev);  R_t = q_t . dq_t
  kernel B, reverse walk carrying (dC, dnv):
    dv_j = sum_t E_tj dH_t / N_t + kw_j (k_j dC);  dk_j = sum_t dS_tj q_t + kw_j (dC v_j + dnv)
    dC <- decay dC + sum_t (w_t s / N_t) q_t (x) dH_t;  dnv <- decay dnv + sum_t w_t s dn_t q_t
    K_j = k_j . dk_j;  di_j = K_j;  dlogf_s = sum_{t>=s} (R_t - K_t);  df = dlogf * sigmoid(-f)
"""
import math

import torch
import torch.nn.functional as F


def _chunks(S, L):
    return [(c * L, min(S, (c + 1) * L)) for c in range((S + L - 1) // L)]


def emu_forward(q, k, v, i, f, L=128, eps=1e-6, c0=None, n0=None, m0=None, mma_dtype=None):
    """Returns h, n_row, m_row, (C, n, m) last, and per-chunk entry states (for kernel A)."""
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dt = q.dtype
    s = 1.0 / math.sqrt(DK)
    C = torch.zeros(B, NH, DK, DV, dtype=dt) if c0 is None else c0.clone()
    n = torch.zeros(B, NH, DK, dtype=dt) if n0 is None else n0.clone()
    m = torch.zeros(B, NH, dtype=dt) if m0 is None else m0.reshape(B, NH).clone()
    rnd = (lambda x: x.to(mma_dtype).to(dt)) if mma_dtype is not None else (lambda x: x)
    h = torch.empty(B, NH, S, DV, dtype=dt)
    n_row = torch.empty(B, NH, S, dtype=dt)
    m_row = torch.empty(B, NH, S, dtype=dt)
    entry = []
    logf = F.logsigmoid(f)
    for (a, e) in _chunks(S, L):
        entry.append((C.clone(), n.clone(), m.clone()))
        qc, kc, vc = q[:, :, a:e], k[:, :, a:e], v[:, :, a:e]
        b = logf[:, :, a:e].cumsum(-1)
        u = i[:, :, a:e] - b
        M = torch.maximum(m[..., None], u.cummax(-1).values)
        Lc = e - a
        tri = torch.ones(Lc, Lc, dtype=torch.bool).tril()
        D = torch.exp(u[..., None, :] - M[..., :, None]).masked_fill(~tri, 0.0)
        w = torch.exp(m[..., None] - M)
        E = s * (qc @ kc.transpose(-1, -2)) * D
        nr = E.sum(-1) + w * s * (qc * rnd(n)[..., None, :]).sum(-1)
        mr = b + M
        N = torch.maximum(nr.abs(), torch.exp(-mr)) + eps
        num = rnd(E) @ vc + (w * s)[..., None] * (qc @ rnd(C))
        h[:, :, a:e] = num / N[..., None]
        n_row[:, :, a:e] = nr
        m_row[:, :, a:e] = mr
        ML = M[..., -1]
        kw = torch.exp(u - ML[..., None])
        decay = torch.exp(m - ML)
        kbar = rnd(kc * kw[..., None])
        C = decay[..., None, None] * C + kbar.transpose(-1, -2) @ vc
        n = decay[..., None] * n + kbar.sum(-2)
        m = b[..., -1] + ML
    return h, n_row, m_row, (C, n, m), entry


def emu_backward(q, k, v, i, f, dh, n_row, m_row, L=128, eps=1e-6, c0=None, n0=None, m0=None, mma_dtype=None):
    B, NH, S, DK = q.shape
    DV = v.shape[-1]
    dt = q.dtype
    s = 1.0 / math.sqrt(DK)
    rnd = (lambda x: x.to(mma_dtype).to(dt)) if mma_dtype is not None else (lambda x: x)
    logf = F.logsigmoid(f)
    chunks = _chunks(S, L)
    m_init = torch.zeros(B, NH, dtype=dt) if m0 is None else m0.reshape(B, NH)

    def gate_terms(a, e):
        b = logf[:, :, a:e].cumsum(-1)
        u = i[:, :, a:e] - b
        M = m_row[:, :, a:e] - b                      # kernel B recovers M_t from the saved m_t
        m_prev = m_init if a == 0 else m_row[:, :, a - 1]
        tri = torch.ones(e - a, e - a, dtype=torch.bool).tril()
        D = torch.exp(u[..., None, :] - M[..., :, None]).masked_fill(~tri, 0.0)
        w = torch.exp(m_prev[..., None] - M)
        N = torch.maximum(n_row[:, :, a:e].abs(), torch.exp(-m_row[:, :, a:e])) + eps
        return b, u, M, m_prev, D, w, N

    # ---- kernel A: forward walk ------------------------------------------------------
    C = torch.zeros(B, NH, DK, DV, dtype=dt) if c0 is None else c0.clone()
    n = torch.zeros(B, NH, DK, dtype=dt) if n0 is None else n0.clone()
    dq = torch.empty_like(q)
    dn_row = torch.empty(B, NH, S, dtype=dt)
    R = torch.empty(B, NH, S, dtype=dt)
    for (a, e) in chunks:
        qc, kc, vc, dhc = q[:, :, a:e], k[:, :, a:e], v[:, :, a:e], dh[:, :, a:e]
        b, u, M, m_prev, D, w, N = gate_terms(a, e)
        E = s * (qc @ kc.transpose(-1, -2)) * D
        Z = dhc @ vc.transpose(-1, -2)
        G = dhc @ rnd(C).transpose(-1, -2)
        delta = (E * Z).sum(-1) + w * s * (qc * G).sum(-1)
        nr = n_row[:, :, a:e]
        active = (nr.abs() >= torch.exp(-m_row[:, :, a:e])).to(dt)
        dn = -active * torch.sign(nr) * delta / (N * N)
        dS = s * (Z / N[..., None] + dn[..., None]) * D
        dqc = rnd(dS) @ kc + (w * s)[..., None] * (G / N[..., None] + dn[..., None] * rnd(n)[..., None, :])
        dq[:, :, a:e] = dqc
        dn_row[:, :, a:e] = dn
        R[:, :, a:e] = (qc * dqc).sum(-1)
        ML = M[..., -1]
        kw = torch.exp(u - ML[..., None])
        decay = torch.exp(m_prev - ML)
        kbar = rnd(kc * kw[..., None])
        C = decay[..., None, None] * C + kbar.transpose(-1, -2) @ vc
        n = decay[..., None] * n + kbar.sum(-2)

    # ---- kernel B: reverse walk ------------------------------------------------------
    dC = torch.zeros(B, NH, DK, DV, dtype=dt)
    dnv = torch.zeros(B, NH, DK, dtype=dt)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    di = torch.empty_like(i)
    df = torch.empty_like(f)
    carry = torch.zeros(B, NH, dtype=dt)
    for (a, e) in reversed(chunks):
        qc, kc, vc, dhc = q[:, :, a:e], k[:, :, a:e], v[:, :, a:e], dh[:, :, a:e]
        b, u, M, m_prev, D, w, N = gate_terms(a, e)
        dn = dn_row[:, :, a:e]
        St = s * (kc @ qc.transpose(-1, -2))                       # S^T  [j, t]
        Zt = vc @ dhc.transpose(-1, -2)                            # Z^T  [j, t]
        Dt = D.transpose(-1, -2)
        Et = St * Dt / N[..., None, :]                             # E^T with 1/N_t folded in
        dSt = s * (Zt / N[..., None, :] + dn[..., None, :]) * Dt
        ML = M[..., -1]
        kw = torch.exp(u - ML[..., None])
        decay = torch.exp(m_prev - ML)
        kbar = rnd(kc * kw[..., None])
        vbar = rnd(vc * kw[..., None])
        dvc = rnd(Et) @ dhc + kbar @ rnd(dC)
        dkc = rnd(dSt) @ qc + vbar @ rnd(dC).transpose(-1, -2) + kw[..., None] * dnv[..., None, :]
        dk[:, :, a:e] = dkc
        dv[:, :, a:e] = dvc
        qbar = rnd(qc * (w * s / N)[..., None])
        dC = decay[..., None, None] * dC + qbar.transpose(-1, -2) @ dhc
        dnv = decay[..., None] * dnv + ((w * s * dn)[..., None] * qc).sum(-2)
        Kr = (kc * dkc).sum(-1)
        di[:, :, a:e] = Kr
        dB = R[:, :, a:e] - Kr
        rc = dB.flip(-1).cumsum(-1).flip(-1) + carry[..., None]   # reverse cumsum with carry
        df[:, :, a:e] = rc * torch.sigmoid(-f[:, :, a:e])
        carry = rc[..., 0]
    return dq, dk, dv, di, df
