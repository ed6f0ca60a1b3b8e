"""The oracle against the reference's own outputs (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference) and against itself."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mlstm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


SEQ = ["seq_randgate_s64_dh16", "seq_refinit_s128_dh32", "seq_strongforget_s96_dh16"]


@pytest.mark.parametrize("name", SEQ)
def test_chunkwise_matches_reference_outputs(name):
    g = _load(name)
    h = O.mlstm_chunkwise(g["q"], g["k"], g["v"], g["i"], g["f"], chunk_size=int(g["chunk_size"]), eps=g["eps"])
    assert _rel(h, g["h_chunkwise"]) < 1e-12
    assert _rel(h, g["h_recurrent"]) < 1e-10


@pytest.mark.parametrize("name", SEQ)
@pytest.mark.parametrize("L", [8, 24, 64, 100])
def test_chunk_size_invariance_and_tail(name, L):
    # The reference needs S % L == 0; the oracle masks the tail.  Output must not depend on L.
    g = _load(name)
    h = O.mlstm_chunkwise(g["q"], g["k"], g["v"], g["i"], g["f"], chunk_size=L, eps=g["eps"])
    assert _rel(h, g["h_chunkwise"]) < 1e-10


@pytest.mark.parametrize("name", SEQ)
def test_recurrent_and_parallel_match_reference(name):
    g = _load(name)
    hr = O.mlstm_recurrent(g["q"], g["k"], g["v"], g["i"], g["f"], eps=g["eps"])
    assert _rel(hr, g["h_recurrent"]) < 1e-10
    hp = O.mlstm_parallel(g["q"], g["k"], g["v"], g["i"], g["f"], eps=g["eps"])
    assert _rel(hp, g["h_parallel"]) < 1e-10


@pytest.mark.parametrize("name", SEQ)
def test_gradients_match_reference_autograd(name):
    g = _load(name)
    out = O.mlstm_fwbw(g["q"], g["k"], g["v"], g["i"], g["f"], g["dh"], chunk_size=int(g["chunk_size"]), eps=g["eps"])
    for got, key in zip(out[1:], ["dq", "dk", "dv", "di", "df"]):
        assert _rel(got, g[key]) < 1e-9, key
    # tail-masked chunking gives the same gradients
    out2 = O.mlstm_fwbw(g["q"], g["k"], g["v"], g["i"], g["f"], g["dh"], chunk_size=40, eps=g["eps"])
    for got, key in zip(out2[1:], ["dq", "dk", "dv", "di", "df"]):
        assert torch.isfinite(got).all()
        assert _rel(got, g[key]) < 1e-8, key


def test_states_match_reference():
    g = _load("states_s48_dh16")
    h, (C, n, m) = O.mlstm_chunkwise(g["q"], g["k"], g["v"], g["i"], g["f"], c_initial=g["c_initial"],
                                     n_initial=g["n_initial"], m_initial=g["m_initial"],
                                     chunk_size=int(g["chunk_size"]), eps=g["eps"], return_last_states=True)
    assert _rel(h, g["h"]) < 1e-12
    assert _rel(C, g["c_last"]) < 1e-12
    assert _rel(n, g["n_last"]) < 1e-12
    assert _rel(m, g["m_last"]) < 1e-12
    # splitting the sequence and carrying the state is the same as one pass
    s = 20
    h1, st = O.mlstm_chunkwise(g["q"][:, :, :s], g["k"][:, :, :s], g["v"][:, :, :s], g["i"][:, :, :s], g["f"][:, :, :s],
                               c_initial=g["c_initial"], n_initial=g["n_initial"], m_initial=g["m_initial"],
                               chunk_size=16, eps=g["eps"], return_last_states=True)
    h2 = O.mlstm_chunkwise(g["q"][:, :, s:], g["k"][:, :, s:], g["v"][:, :, s:], g["i"][:, :, s:], g["f"][:, :, s:],
                           c_initial=st[0], n_initial=st[1], m_initial=st[2], chunk_size=16, eps=g["eps"])
    assert _rel(torch.cat([h1, h2], 2), g["h"]) < 1e-10
    # recurrent form agrees on states too
    hr, (Cr, nr, mr) = O.mlstm_recurrent(g["q"], g["k"], g["v"], g["i"], g["f"], g["c_initial"], g["n_initial"],
                                         g["m_initial"], eps=g["eps"], return_last_states=True)
    assert _rel(hr, g["h"]) < 1e-10 and _rel(Cr, g["c_last"]) < 1e-10 and _rel(mr, g["m_last"]) < 1e-10


def test_reverse_is_flip_cell_flip():
    g = _load("seq_randgate_s64_dh16")
    args = [g[k] for k in "qkvif"]
    flipped = [a.flip(dims=[2]) for a in args]
    want = O.mlstm_chunkwise(*flipped, chunk_size=16, eps=1e-6).flip(dims=[2])
    got = O.mlstm_chunkwise(*args, chunk_size=24, eps=1e-6, reverse=True)
    assert _rel(got, want) < 1e-10
    assert _rel(O.mlstm_recurrent(*args, eps=1e-6, reverse=True), want) < 1e-10


def test_cell_matches_vendored_cell():
    g = _load("cell_vendored_s32_h64")
    y = O.cell_forward(g["q"], g["k"], g["v"], int(g["num_heads"]), g["igate_w"], g["igate_b"], g["fgate_w"],
                       g["fgate_b"], g["outnorm_w"], None, chunk_size=16, eps=g["eps"], norm_eps=g["norm_eps"],
                       form="parallel")
    assert _rel(y, g["y"]) < 1e-10
    # chunkwise backend (what MatrixLSTMCell selects) differs only through the m_0 = 0 floor
    y2 = O.cell_forward(g["q"], g["k"], g["v"], int(g["num_heads"]), g["igate_w"], g["igate_b"], g["fgate_w"],
                        g["fgate_b"], g["outnorm_w"], None, chunk_size=16, eps=g["eps"], norm_eps=g["norm_eps"])
    assert _rel(y2, g["y"]) < 1e-3


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "*.npz"))) >= 5


@pytest.mark.parametrize("DK,DV,DP,reverse", [(8, 12, 16, False), (16, 16, 64, True), (24, 8, 32, False)])
def test_zero_padded_head_dims_change_nothing(DK, DV, DP, reverse):
    """The identity the library's padded tensor-core path rests on (csrc/mlstm_api.cu), pinned on the oracle in
    fp64: q, k, v, dh padded with zero columns to a common DP — and the initial state with zero rows / columns — give the same
    h, dq, dk, dv in the leading columns (zeros behind them), the same di, df and the same leading block of the last state,
    provided the 1/sqrt(DHqk) scale of the ORIGINAL head dim is kept (here: folded into q, since the oracle, like the reference,
    takes the scale from the tensor's last dimension, backends.py:168)."""
    import torch.nn.functional as F
    B, NH, S = 1, 2, 70
    g = torch.Generator().manual_seed(3)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    q, k, v, dh = rn(B, NH, S, DK), rn(B, NH, S, DK), rn(B, NH, S, DV), rn(B, NH, S, DV)
    i, f = rn(B, NH, S), rn(B, NH, S) + 3.0
    st = dict(c_initial=rn(B, NH, DK, DV), n_initial=rn(B, NH, DK), m_initial=rn(B, NH, 1))
    small = O.mlstm_fwbw(q, k, v, i, f, dh, chunk_size=32, eps=1e-6, reverse=reverse, **st)
    r = (DP / DK) ** 0.5
    pad = lambda t, d: F.pad(t, (0, DP - d))
    stp = dict(c_initial=F.pad(st["c_initial"], (0, DP - DV, 0, DP - DK)), n_initial=pad(st["n_initial"], DK), m_initial=st["m_initial"])
    big = O.mlstm_fwbw(pad(q, DK) * r, pad(k, DK), pad(v, DV), i, f, pad(dh, DV), chunk_size=32, eps=1e-6, reverse=reverse, **stp)
    h, dq, dk, dv, di, df = small
    hp, dqp, dkp, dvp, dip, dfp = big
    tol = 1e-11
    assert _rel(hp[..., :DV], h) < tol and hp[..., DV:].abs().max() == 0
    assert _rel(dqp[..., :DK] * r, dq) < tol and _rel(dkp[..., :DK], dk) < tol and _rel(dvp[..., :DV], dv) < tol
    assert dkp[..., DK:].abs().max() == 0 and _rel(dip, di) < tol and _rel(dfp, df) < tol
    _, (C, n, m) = O.mlstm_chunkwise(q, k, v, i, f, **st, chunk_size=32, return_last_states=True, reverse=reverse)
    _, (Cp, np_, mp) = O.mlstm_chunkwise(pad(q, DK) * r, pad(k, DK), pad(v, DV), i, f, **stp, chunk_size=32, return_last_states=True,
                                         reverse=reverse)
    assert _rel(Cp[:, :, :DK, :DV], C) < tol and _rel(np_[..., :DK], n) < tol and _rel(mp, m) < tol
    assert Cp[:, :, DK:].abs().max() == 0 and Cp[..., DV:].abs().max() == 0
