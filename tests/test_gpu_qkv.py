"""GPU parity of the operand producer (csrc/mlstm_qkv.cu: depthwise 3x3 conv + SiLU + block-diagonal q / k / v in one
kernel; reference ops: vision_lstm_util.py:96-129, vision_lstm2.py:987-1022, 482-491) against the same four ops in fp64
PyTorch on the same (16-bit rounded) inputs, and of the ViL layer with and without it.

Tolerance: outputs are bf16, error metric max|a-b| / max|b| < 1e-2 (the north-star bf16 bound); gradients of the layer with
the producer against the layer without it < 2e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    b = b.double().cpu()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make(B, gh, gw, NH, d, dtype, seed=0, bias=True, ld_factor=2):
    g = torch.Generator().manual_seed(seed)
    D, S = NH * d, gh * gw
    up = torch.randn(B, S, ld_factor * D, generator=g).to(dtype)          # proj_up's output: x is its first D columns
    conv_w = torch.randn(D, 1, 3, 3, generator=g) * 0.3
    conv_b = torch.randn(D, generator=g) * 0.1 if bias else None
    ws = [torch.randn(NH, d, d, generator=g) * d ** -0.5 for _ in range(3)]
    bs = [torch.randn(D, generator=g) * 0.1 if bias else None for _ in range(3)]
    return up, conv_w, conv_b, ws, bs


CASES = [
    # B, gh, gw, NH, d, dtype, rotate, bias
    (2, 20, 20, 4, 64, torch.bfloat16, False, True),     # cfg2 layer (inner 256, 400 tokens: 3 tiles + 16 rows)
    (2, 20, 20, 4, 64, torch.float16, True, True),
    (2, 40, 40, 4, 128, torch.bfloat16, False, True),    # cfg3 layer (inner 512, 1600 tokens)
    (2, 40, 40, 4, 128, torch.float16, True, False),
    (1, 80, 80, 2, 128, torch.bfloat16, True, True),     # widest supported grid (halo 88 rows)
    (3, 7, 9, 8, 64, torch.bfloat16, False, True),       # less than one tile, odd grid
    (1, 16, 8, 2, 128, torch.float16, False, True),      # exactly one tile
    (40, 20, 20, 4, 128, torch.bfloat16, False, True),   # more tiles than CTAs per block
]


@pytest.mark.parametrize("B,gh,gw,NH,d,dtype,rotate,bias", CASES)
def test_producer_matches_fp64_ops(B, gh, gw, NH, d, dtype, rotate, bias):
    from xlstm_yolo_b200 import ops
    up, conv_w, conv_b, ws, bs = make(B, gh, gw, NH, d, dtype, bias=bias)
    D = NH * d
    x = up.cuda()[..., :D]
    assert ops.qkv_supported(x, D, NH, gh, gw)
    cu = lambda t: None if t is None else t.cuda()
    got = ops.qkv_producer(x, cu(conv_w), cu(conv_b), cu(ws[0]), cu(bs[0]), cu(ws[1]), cu(bs[1]), cu(ws[2]), cu(bs[2]), gh, gw, rotate)
    torch.cuda.synchronize()
    # the kernel's operands: q, k weights rounded to bf16, the v weight to the activation dtype
    r = lambda t, dt: t.to(dt).double()
    dbl = lambda t: None if t is None else t.double()
    ref = ops.qkv_reference(up[..., :D].double(), conv_w.double(), dbl(conv_b), r(ws[0], torch.bfloat16), dbl(bs[0]),
                            r(ws[1], torch.bfloat16), dbl(bs[1]), r(ws[2], dtype), dbl(bs[2]), gh, gw, rotate)
    for name, a, b in zip("cqkv", got, ref):
        assert a.dtype == torch.bfloat16 and torch.isfinite(a).all(), name
        assert rel(a, b) < 1e-2, f"{name}: {rel(a, b):.3e}"


def test_unsupported_shapes_are_refused():
    from xlstm_yolo_b200 import ops
    x = torch.zeros(1, 100 * 100, 256, dtype=torch.bfloat16, device="cuda")
    assert not ops.qkv_supported(x, 256, 4, 100, 100)                       # grid wider than the staged halo
    assert not ops.qkv_supported(x[:, :400, :64], 64, 4, 20, 20)            # block size 16
    assert not ops.qkv_supported(x[:, :400].float(), 256, 4, 20, 20)        # fp32 activations stay on cuDNN / cuBLAS


@pytest.mark.parametrize("direction,autocast", [("tl", None), ("br", None), ("br", torch.float16)])
def test_layer_with_producer_matches_layer_without(direction, autocast):
    from xlstm_yolo_b200.vil import SequenceTraversal, ViLLayer
    torch.manual_seed(0)
    d = SequenceTraversal.ROWWISE_FROM_TOP_LEFT if direction == "tl" else SequenceTraversal.ROWWISE_FROM_BOT_RIGHT
    layer = ViLLayer(dim=128, direction=d, qkv_block_size=64, chunk_size=64).cuda()
    for prm in (layer.conv.bias, layer.q_proj.bias, layer.k_proj.bias, layer.v_proj.bias):
        torch.nn.init.normal_(prm, std=0.1)
    if autocast is None:
        layer = layer.to(torch.bfloat16)
    x = torch.randn(2, 400, 128, device="cuda", dtype=torch.bfloat16 if autocast is None else torch.float32)
    outs = []
    for fused in (True, False):
        layer.fused_producer = fused
        layer.zero_grad()
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
            y = layer(xi)
        # fp16 autocast: loss scaled the way the trainer's GradScaler does, or the fp16 gradients underflow in both paths
        (y.float().square().mean() * (1.0 if autocast is None else 4096.0)).backward()
        outs.append([y.detach().float(), xi.grad.float()] + [prm.grad.float().clone() for prm in
                    (layer.conv.weight, layer.conv.bias, layer.q_proj.weight, layer.q_proj.bias, layer.k_proj.weight,
                     layer.v_proj.weight, layer.v_proj.bias, layer.proj_up.weight)])
    names = ["y", "dx", "dconv_w", "dconv_b", "dWq", "dbq", "dWk", "dWv", "dbv", "dW_up"]
    for n, a, b in zip(names, *outs):
        assert torch.isfinite(a).all(), n
        assert rel(a, b) < (1e-2 if n == "y" else 3e-2), f"{n}: {rel(a, b):.3e}"


@pytest.mark.parametrize("B,S,NH,d,xdtype,with_dc", [(2, 400, 4, 64, torch.bfloat16, True), (2, 1600, 4, 128, torch.bfloat16, True),
                                                     (3, 129, 2, 128, torch.float16, True), (1, 70, 8, 64, torch.float16, False),
                                                     (40, 400, 4, 128, torch.bfloat16, False)])
def test_projection_backward_kernel_matches_fp64(B, S, NH, d, xdtype, with_dc):
    """qkv_bwd_kernel (five GEMMs + bias sums in one kernel, weight gradients accumulated in TMEM over the tiles of a CTA,
    fixed-order reduction) against the same contractions in fp64 on the same 16-bit inputs; bit-identical when repeated."""
    from xlstm_yolo_b200 import ops
    g = torch.Generator().manual_seed(3)
    D, T = NH * d, B * S
    up = torch.randn(B, S, 2 * D, generator=g).to(xdtype)
    c, dq, dk, dv, dc = (torch.randn(B, S, D, generator=g).bfloat16() for _ in range(5))
    ws = [(torch.randn(NH, d, d, generator=g) * d ** -0.5).bfloat16() for _ in range(3)]
    x = up.cuda()[..., :D]
    cu = lambda t: t.cuda()
    run = lambda: ops.qkv_proj_backward(x, cu(c), cu(ws[0]), cu(ws[1]), cu(ws[2]), cu(dc) if with_dc else None, cu(dq), cu(dk), cu(dv))
    got = run()
    torch.cuda.synchronize()
    hd = lambda t: t.double().reshape(T, NH, d)
    dxc = torch.einsum("tho,hoi->thi", hd(dq), ws[0].double()) + torch.einsum("tho,hoi->thi", hd(dk), ws[1].double())
    if with_dc:
        dxc = dxc + hd(dc)
    dxv = torch.einsum("tho,hoi->thi", hd(dv), ws[2].double())
    xd = up[..., :D].double().reshape(T, NH, d)
    dws = [torch.einsum("tho,thi->hoi", hd(dq), hd(c)), torch.einsum("tho,thi->hoi", hd(dk), hd(c)), torch.einsum("tho,thi->hoi", hd(dv), xd)]
    db = torch.stack([t.double().reshape(T, D).sum(0) for t in (dq, dk, dv)])
    ref = [dxc.reshape(B, S, D), dxv.reshape(B, S, D)] + dws + [db]
    for name, a, b_ in zip(["dxc", "dxv", "dWq", "dWk", "dWv", "db"], got, ref):
        assert torch.isfinite(a).all(), name
        assert rel(a, b_) < (1e-2 if name.startswith("dx") else 2e-3), f"{name}: {rel(a, b_):.3e}"
    again = run()
    for a, b_ in zip(got, again):
        assert torch.equal(a, b_)


def test_half_precision_inference_pair_runs_with_and_without_the_producer():
    """engine/validator.py:117-119 runs the detector as model.half() without autocast: fp16 parameters and activations around
    bf16 cell kernels.  The pair with the producer kernel agrees with the pair without it (regression: the tail used to assume
    conv_act and z share a dtype)."""
    from xlstm_yolo_b200 import ViLBlockPair
    torch.manual_seed(0)
    pair = ViLBlockPair(dim=128, chunk_size=64, qkv_block_size=64).cuda().half().eval()
    x = torch.randn(2, 400, 128, device="cuda", dtype=torch.float16)
    outs = []
    for fused in (True, False):
        for blk in (pair.rowwise_from_top_left, pair.rowwise_from_bot_right):
            blk.layer.fused_producer = fused
        with torch.no_grad():
            outs.append(pair(x).float())
    assert outs[0].dtype == torch.float32 and torch.isfinite(outs[0]).all()
    assert rel(outs[0], outs[1]) < 1e-2


@pytest.mark.parametrize("B,gh,gw,D,xdtype,rotate,bias", [(2, 20, 20, 256, torch.bfloat16, False, True), (2, 40, 40, 512, torch.bfloat16, True, True),
                                                          (1, 80, 80, 128, torch.float16, False, False), (3, 7, 9, 64, torch.float16, True, True),
                                                          (40, 20, 20, 256, torch.bfloat16, False, True)])
def test_conv_silu_backward_kernel_matches_fp64(B, gh, gw, D, xdtype, rotate, bias):
    """conv_bwd_kernel (du = dxc * silu'(u) in place over the staged rows, transposed depthwise conv, weight / bias gradients through
    fixed-order partials) against autograd of silu(conv(x)) in fp64 with the same 16-bit inputs; bit-identical when repeated."""
    from xlstm_yolo_b200 import ops
    g = torch.Generator().manual_seed(5)
    S = gh * gw
    up = torch.randn(B, S, 2 * D, generator=g).to(xdtype)
    conv_w = torch.randn(D, 1, 3, 3, generator=g) * 0.3
    conv_b = torch.randn(D, generator=g) * 0.1 if bias else None
    dxc, dxv = (torch.randn(B, S, D, generator=g).bfloat16() for _ in range(2))
    xd = up[..., :D].double().requires_grad_(True)
    wd = conv_w.double().requires_grad_(True)
    bd = conv_b.double().requires_grad_(True) if bias else None
    w_used = wd.flip(-1, -2) if rotate else wd
    u = torch.nn.functional.conv2d(xd.reshape(B, gh, gw, D).permute(0, 3, 1, 2), w_used, bd, padding=1, groups=D)
    # the kernel's operands: sp rounded to bf16, du = dxc * sp rounded to bf16
    sg = torch.sigmoid(u.detach())
    sp = (sg * (1 + u.detach() * (1 - sg))).permute(0, 2, 3, 1).reshape(B, S, D).bfloat16()
    du = (dxc.double() * sp.double()).bfloat16().double()
    grads = torch.autograd.grad(u, [xd, wd] + ([bd] if bias else []), du.reshape(B, gh, gw, D).permute(0, 3, 1, 2))
    want = [grads[0] + dxv.double(), grads[1]] + ([grads[2]] if bias else [])
    x = up.cuda()[..., :D]
    run = lambda: ops.conv_silu_backward(x, sp.cuda(), dxc.cuda(), dxv.cuda(), conv_w.cuda(), bias, gh, gw, rotate)
    got = run()
    torch.cuda.synchronize()
    assert got[0].dtype == xdtype and (got[2] is None) == (not bias)
    for name, a, b_ in zip(["dx", "dwc", "dbc"], got, want):
        assert torch.isfinite(a).all(), name
        assert rel(a, b_) < (1e-2 if name == "dx" else 2e-3), f"{name}: {rel(a, b_):.3e}"
    again = run()
    for a, b_ in zip(got, again):
        assert a is None or torch.equal(a, b_)
