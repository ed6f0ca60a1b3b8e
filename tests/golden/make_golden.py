"""Generate golden vectors by running the REFERENCE's own PyTorch mLSTM functions.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports, by file path, ``nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py``
(``parallel_stabilized_simple`` :9, ``recurrent_step_stabilized_simple`` :93,
``chunkwise_simple`` :149) and the vendored ``xlstm`` package's ``mLSTMCell``
(``blocks/mlstm/cell.py:20``) from the read-only reference tree, evaluates them in fp64 on
seeded inputs, and writes ``tests/golden/*.npz``.  Nothing from the reference is copied
into the repo; only inputs/outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("MLSTM_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _load_backends():
    path = os.path.join(REF, "nn/modules/vision_lstm/xlstm/blocks/mlstm/backends.py")
    spec = importlib.util.spec_from_file_location("ref_backends", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(seed, B, NH, S, DH, i_mean, i_std, f_lo=3.0, f_hi=6.0):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, NH, S, DH, generator=g, dtype=torch.float64)
    k = torch.randn(B, NH, S, DH, generator=g, dtype=torch.float64)
    v = torch.randn(B, NH, S, DH, generator=g, dtype=torch.float64)
    i = i_mean + i_std * torch.randn(B, NH, S, generator=g, dtype=torch.float64)
    f = torch.linspace(f_lo, f_hi, NH, dtype=torch.float64).view(1, NH, 1) + torch.randn(
        B, NH, S, generator=g, dtype=torch.float64)
    dh = torch.randn(B, NH, S, DH, generator=g, dtype=torch.float64)
    return q, k, v, i, f, dh


def _np(**kw):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()}


def case_sequence(rb, name, seed, B, NH, S, DH, L, i_mean, i_std, eps, f_lo=3.0, f_hi=6.0):
    q, k, v, i, f, dh = _inputs(seed, B, NH, S, DH, i_mean, i_std, f_lo, f_hi)
    leaves = [t.clone().requires_grad_(True) for t in (q, k, v, i, f)]
    h_chunk = rb.chunkwise_simple(*leaves, chunk_size=L, eps=eps)
    h_chunk.backward(dh)
    grads = [t.grad.clone() for t in leaves]
    h_par = rb.parallel_stabilized_simple(q, k, v, i.unsqueeze(-1), f.unsqueeze(-1), eps=eps)
    # recurrent: one reference step per token, zero initial state
    C = torch.zeros(B, NH, DH, DH, dtype=torch.float64)
    n = torch.zeros(B, NH, DH, 1, dtype=torch.float64)
    m = torch.zeros(B, NH, 1, 1, dtype=torch.float64)
    hs = []
    for t in range(S):
        ht, (C, n, m) = rb.recurrent_step_stabilized_simple(
            C, n, m, q[:, :, t:t + 1].clone(), k[:, :, t:t + 1].clone(), v[:, :, t:t + 1].clone(),
            i[:, :, t:t + 1, None], f[:, :, t:t + 1, None], eps=eps)
        hs.append(ht)
    h_rec = torch.cat(hs, dim=2)
    np.savez(os.path.join(HERE, name + ".npz"), **_np(
        q=q, k=k, v=v, i=i, f=f, dh=dh, chunk_size=L, eps=eps,
        h_chunkwise=h_chunk, h_parallel=h_par, h_recurrent=h_rec,
        dq=grads[0], dk=grads[1], dv=grads[2], di=grads[3], df=grads[4]))
    print(name, "chunk-vs-rec", (h_chunk - h_rec).abs().max().item(),
          "chunk-vs-par", (h_chunk - h_par).abs().max().item())


def case_states(rb, name, seed, B, NH, S, DH, L, eps):
    q, k, v, i, f, dh = _inputs(seed, B, NH, S, DH, 0.0, 1.0)
    g = torch.Generator().manual_seed(seed + 1000)
    C0 = torch.randn(B, NH, DH, DH, generator=g, dtype=torch.float64)
    n0 = torch.randn(B, NH, DH, generator=g, dtype=torch.float64)
    m0 = torch.randn(B, NH, 1, generator=g, dtype=torch.float64)
    # NB: the reference documents initial_m as (B,NH,1) but indexes it as (B,NH)
    # (backends.py:194 broadcasts `initial_m[:, :, None, None]` into a (B,NH,1,1) slot).
    h, (C1, n1, m1) = rb.chunkwise_simple(q, k, v, i, f, initial_C=C0, initial_n=n0,
                                          initial_m=m0.reshape(B, NH),
                                          chunk_size=L, return_last_state=True, eps=eps)
    np.savez(os.path.join(HERE, name + ".npz"), **_np(
        q=q, k=k, v=v, i=i, f=f, c_initial=C0, n_initial=n0, m_initial=m0, chunk_size=L, eps=eps,
        h=h, c_last=C1, n_last=n1, m_last=m1.reshape(B, NH, 1)))
    print(name, "ok")


def case_cell(name, seed, B, S, H, NH):
    sys.path.insert(0, os.path.join(REF, "nn/modules/vision_lstm"))
    from xlstm.blocks.mlstm.cell import mLSTMCell, mLSTMCellConfig  # vendored xlstm 2.0.3
    torch.manual_seed(seed)
    cell = mLSTMCell(mLSTMCellConfig(context_length=S, embedding_dim=H, num_heads=NH)).double()
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.05)
        cell.fgate.weight.normal_(0, 0.05)
        cell.outnorm.weight.normal_(0, 0.1)
    q, k, v = (torch.randn(B, S, H, dtype=torch.float64) for _ in range(3))
    y = cell(q, k, v)
    np.savez(os.path.join(HERE, name + ".npz"), **_np(
        q=q, k=k, v=v, num_heads=NH, igate_w=cell.igate.weight, igate_b=cell.igate.bias,
        fgate_w=cell.fgate.weight, fgate_b=cell.fgate.bias, outnorm_w=cell.outnorm.weight,
        eps=1e-6, norm_eps=cell.outnorm.eps, y=y))
    print(name, "ok")


def main():
    rb = _load_backends()
    case_sequence(rb, "seq_randgate_s64_dh16", 0, 1, 2, 64, 16, 16, 0.0, 1.0, 1e-6)
    case_sequence(rb, "seq_refinit_s128_dh32", 1, 1, 2, 128, 32, 32, -10.0, 0.1, 5e-5)
    case_sequence(rb, "seq_strongforget_s96_dh16", 2, 1, 2, 96, 16, 32, 0.0, 2.0, 1e-6, f_lo=-2.0, f_hi=1.0)
    case_states(rb, "states_s48_dh16", 3, 1, 2, 48, 16, 16, 1e-6)
    case_cell("cell_vendored_s32_h64", 4, 2, 32, 64, 4)


if __name__ == "__main__":
    main()
