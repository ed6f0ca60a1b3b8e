"""Fused layer tail (csrc/mlstm_glue.cu): out-norm + learnable skip + SiLU(z) gate vs the reference's
separate ops (vision_lstm2.py:950 via :1309-1325, :498-499) evaluated in fp64."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).detach())


def reference(h, c, z, w, b, skip, eps):
    B, NH, S, DH = h.shape
    x = h.transpose(1, 2).reshape(B * S, NH * DH)
    n = F.group_norm(x, num_groups=NH, weight=1.0 + w, bias=b, eps=eps).view(B, S, NH * DH)
    return (n + skip * c) * F.silu(z)


SHAPES = [(2, 100, 4, 64), (1, 37, 4, 128), (2, 50, 32, 16), (1, 40, 4, 256), (1, 3, 2, 128)]


@pytest.mark.parametrize("B,S,NH,DH", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("bias", [True, False])
def test_layer_tail_forward_backward(B, S, NH, DH, dtype, bias):
    from xlstm_yolo_b200 import ops
    D = NH * DH
    g = torch.Generator().manual_seed(0)
    h = (torch.randn(B, S, NH, DH, generator=g) * 0.7 + 0.1).to(dtype).cuda().transpose(1, 2)
    c = torch.randn(B, S, D, generator=g).to(dtype).cuda()
    up = torch.randn(B, S, 2 * D, generator=g).to(dtype).cuda()
    z = up[..., D:]                                   # column slice, row stride 2*D, like chunk(2, -1)
    w = (torch.randn(D, generator=g) * 0.2).cuda()
    b = (torch.randn(D, generator=g) * 0.2).cuda() if bias else None
    skip = (1 + torch.randn(D, generator=g) * 0.2).cuda()
    dy = torch.randn(B, S, D, generator=g).to(dtype).cuda()
    assert ops.glue_supported(h, c, z)
    leaves = [t.detach().clone().requires_grad_(True) for t in (h, c, up, w, skip)] + ([b.clone().requires_grad_(True)] if bias else [])
    hh, cc, uu, ww, ss = leaves[:5]
    bb = leaves[5] if bias else None
    y = ops.layer_tail(hh, cc, uu[..., D:], ww, bb, ss, eps=1e-3)
    y.backward(dy)
    ref_leaves = [t.detach().double().requires_grad_(True) for t in (h, c, up, w, skip)] + ([b.double().requires_grad_(True)] if bias else [])
    rh, rc, ru, rw, rs = ref_leaves[:5]
    rb = ref_leaves[5] if bias else None
    yr = reference(rh, rc, ru[..., D:], rw, rb, rs, 1e-3)
    yr.backward(dy.double())
    ty, tg = (1e-2, 2e-2) if dtype == torch.bfloat16 else (1e-5, 1e-4)
    assert y.shape == (B, S, D) and y.dtype == dtype
    assert rel(y, yr) < ty
    names = ["dh", "dc", "dup", "dw", "dskip"] + (["db"] if bias else [])
    for n, a, r in zip(names, leaves, ref_leaves):
        assert a.grad is not None, n
        assert rel(a.grad, r.grad) < (tg if n in ("dh", "dc", "dup") else 1e-3 if dtype == torch.float32 else 2e-2), n
    # deterministic parameter gradients
    leaves2 = [t.detach().clone().requires_grad_(True) for t in (h, c, up, w, skip)]
    y2 = ops.layer_tail(leaves2[0], leaves2[1], leaves2[2][..., D:], leaves2[3], None if not bias else b, leaves2[4], eps=1e-3)
    y2.backward(dy)
    assert torch.equal(leaves2[3].grad, ww.grad) and torch.equal(leaves2[4].grad, ss.grad) and torch.equal(y2, y)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("direction", ["tl", "br"])
def test_vil_layer_fused_tail_matches_unfused(dtype, direction):
    """Whole ViLLayer (dim 128 -> inner 256, 4 heads of 64) with and without the fused tail."""
    from xlstm_yolo_b200 import MatrixLSTMCell, SequenceTraversal, ViLLayer
    torch.manual_seed(0)
    d = SequenceTraversal.ROWWISE_FROM_TOP_LEFT if direction == "tl" else SequenceTraversal.ROWWISE_FROM_BOT_RIGHT
    layer = ViLLayer(dim=128, direction=d, qkv_block_size=64, chunk_size=64)
    with torch.no_grad():
        layer.mlstm_cell.outnorm.weight.normal_(0, 0.2); layer.mlstm_cell.outnorm.bias.normal_(0, 0.2)
        layer.learnable_skip.normal_(1, 0.2); layer.mlstm_cell.igate.bias.normal_(0, 1)
        layer.mlstm_cell.igate.weight.normal_(0, 0.05); layer.mlstm_cell.fgate.weight.normal_(0, 0.05)
    if dtype == torch.float32:
        old = layer.mlstm_cell
        layer.mlstm_cell = MatrixLSTMCell(dim=old.dim, num_heads=old.num_heads, use_autocast=False)
        layer.mlstm_cell.load_state_dict(old.state_dict())
    fused = layer.cuda().to(dtype).train()
    plain = copy.deepcopy(fused)
    plain.fused_tail = False
    x = torch.randn(2, 144, 128, device="cuda").to(dtype)
    dy = torch.randn(2, 144, 128, device="cuda").to(dtype)
    res = []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for m in (fused, plain):
            xi = x.clone().requires_grad_(True)
            y = m(xi)
            y.backward(dy)
            res.append((y.detach(), xi.grad, m.learnable_skip.grad, m.mlstm_cell.outnorm.weight.grad, m.mlstm_cell.outnorm.bias.grad))
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    tol = 3e-2 if dtype == torch.bfloat16 else 1e-4
    for n, a, r in zip(["y", "dx", "dskip", "dnw", "dnb"], *res):
        assert a is not None and r is not None, n
        assert rel(a, r) < tol, (n, rel(a, r))
