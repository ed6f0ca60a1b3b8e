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
