"""ViL layer stack drop-ins vs golden vectors produced by the reference's own classes
(tests/golden/make_golden_vil.py; reference: vision_lstm2.py:386-530,685-735,1393-1441)."""
import glob
import os

import numpy as np
import pytest
import torch

from xlstm_yolo_b200.vil import (LinearHeadwiseExpand, SequenceConv2d, SequenceTraversal, ViLBlock, ViLBlockPair,
                                 ViLLayer)

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "vil_pair_*.npz")))


def load_pair(path, dtype=torch.float64, device="cpu"):
    z = np.load(path)
    pair = ViLBlockPair(dim=int(z["dim"]), chunk_size=int(z["chunk_size"]), qkv_block_size=int(z["qkv_block_size"]))
    sd = {k[len("param__"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param__")}
    assert set(sd) == set(pair.state_dict()), "state_dict keys differ from the reference's"
    for k, v in pair.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    pair.load_state_dict(sd, strict=True)
    return pair.to(device=device, dtype=dtype).train(), z


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_fixtures_present():
    assert len(GOLD) >= 2


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("flip_free", [True, False])
def test_pair_matches_reference_fp64(path, flip_free):
    pair, z = load_pair(path)
    for blk in (pair.rowwise_from_top_left, pair.rowwise_from_bot_right):
        blk.layer.flip_free = flip_free
    x, dy = torch.from_numpy(z["x"]), torch.from_numpy(z["dy"])
    fns = {"tl": pair.rowwise_from_top_left, "br": pair.rowwise_from_bot_right, "pair": pair}
    for key, fn in fns.items():
        xi = x.clone().requires_grad_(True)
        pair.zero_grad()
        y = fn(xi)
        y.backward(dy)
        assert rel(y.detach(), torch.from_numpy(z[f"y_{key}"])) < 1e-9, key
        assert rel(xi.grad, torch.from_numpy(z[f"dx_{key}"])) < 1e-8, key
    named = dict(pair.named_parameters())
    checked = 0
    for k in z.files:
        if k.startswith("grad__"):
            g = named[k[len("grad__"):]].grad
            assert rel(g, torch.from_numpy(z[k])) < 1e-7, k
            checked += 1
    assert checked >= 8


def test_head_compat_returns_top_left_only():
    pair, z = load_pair(GOLD[0])
    pair.head_compat = True
    y = pair(torch.from_numpy(z["x"]))
    assert rel(y.detach(), torch.from_numpy(z["y_tl"])) < 1e-9


def test_rotated_conv_equals_flip_conv_flip():
    torch.manual_seed(0)
    conv = SequenceConv2d(8, 8, kernel_size=3, padding=1, groups=8, bias=True, seqlens=(4, 6)).double()
    x = torch.randn(2, 24, 8, dtype=torch.float64)
    assert torch.allclose(conv(x, rotate=True), conv(x.flip(1)).flip(1), atol=1e-12)
    with pytest.raises(AssertionError):
        SequenceConv2d(8, 8, kernel_size=3, padding=1, groups=8)(torch.randn(1, 24, 8))   # 24 is not a square


def test_headwise_projection_is_block_diagonal():
    torch.manual_seed(0)
    proj = LinearHeadwiseExpand(dim=12, num_heads=3, bias=True).double()
    x = torch.randn(2, 5, 12, dtype=torch.float64)
    dense = torch.block_diag(*proj.weight.detach())
    assert torch.allclose(proj(x), x @ dense.T + proj.bias, atol=1e-12)


def test_constructor_contract_and_state_dict_names():
    blk = ViLBlock(dim=128, direction=SequenceTraversal.ROWWISE_FROM_TOP_LEFT, qkv_block_size=64, chunk_size=64)
    sd = blk.state_dict()
    expect = {  # SURVEY.md §8(a) state-dict table (dumped from the reference class)
        "norm.weight": (128,), "layer.norm.weight": (128,), "layer.proj_up.weight": (512, 128),
        "layer.conv.weight": (256, 1, 3, 3), "layer.q_proj.weight": (4, 64, 64), "layer.q_proj.bias": (256,),
        "layer.mlstm_cell.igate.weight": (4, 768), "layer.mlstm_cell.fgate.bias": (4,),
        "layer.mlstm_cell.outnorm.weight": (256,), "layer.mlstm_cell.outnorm.bias": (256,),
        "layer.learnable_skip": (256,), "layer.proj_down.weight": (128, 256), "layer.ffn_norm.weight": (128,),
        "layer.ffn.proj_up_gate_z.weight": (768, 128), "layer.ffn.proj_down.weight": (128, 384),
    }
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    assert torch.equal(sd["layer.mlstm_cell.igate.bias"], torch.full((4,), -10.0))
    assert torch.allclose(sd["layer.mlstm_cell.fgate.bias"], torch.tensor([3.0, 4.0, 5.0, 6.0]))
    layer = ViLLayer(dim=32, direction=SequenceTraversal.ROWWISE_FROM_BOT_RIGHT, qkv_block_size=16)
    assert layer.backwards and layer.mlstm_cell.num_heads == 4


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
@pytest.mark.parametrize("dtype,tol_y,tol_g", [(torch.float32, 2e-4, 2e-3), (torch.bfloat16, 3e-2, 6e-2)])
def test_pair_on_gpu_matches_reference(path, dtype, tol_y, tol_g):
    """Whole bidirectional pair through the CUDA cell (both scan directions, no flips).  The
    tolerances are layer-level (norms, GEMMs and the conv run in `dtype` as well), looser than
    the cell-level bounds in test_gpu_parity.py."""
    from xlstm_yolo_b200 import MatrixLSTMCell
    pair, z = load_pair(path, dtype=dtype, device="cuda")
    if dtype == torch.float32:
        # the cell re-casts to bf16 by default, as the reference does (vision_lstm2.py:839); for the
        # fp32 bound run the fp32 kernels and keep cuDNN off TF32
        for blk in (pair.rowwise_from_top_left, pair.rowwise_from_bot_right):
            old = blk.layer.mlstm_cell
            new = MatrixLSTMCell(dim=old.dim, num_heads=old.num_heads, use_autocast=False).to("cuda", dtype)
            new.load_state_dict(old.state_dict())
            blk.layer.mlstm_cell = new
    x = torch.from_numpy(z["x"]).to("cuda", dtype).requires_grad_(True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        y = pair(x)
        y.backward(torch.from_numpy(z["dy"]).to("cuda", dtype))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert rel(y.detach().double().cpu(), torch.from_numpy(z["y_pair"])) < tol_y
    assert rel(x.grad.double().cpu(), torch.from_numpy(z["dx_pair"])) < tol_g
