"""Gate-projection kernels (csrc/mlstm_gates.cu, the tensor-core forward csrc/mlstm_gates_tc.cu) and the fused cell node vs PyTorch on the same inputs.

Reference arithmetic: vision_lstm2.py:895-897 — ``cat[q,k,v]`` then two ``nn.Linear(3*dim, NH)``."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make(B, S, D, NH, dtype, seed=0, padded=False):
    g = torch.Generator().manual_seed(seed)
    ld = D + 8 if padded else D
    q, k, v = ((torch.randn(B, S, ld, generator=g) * 0.5).to(dtype).cuda()[..., :D] for _ in range(3))
    w_i, w_f = (torch.randn(NH, 3 * D, generator=g).mul(0.05).cuda() for _ in range(2))
    b_i, b_f = (torch.randn(NH, generator=g).cuda() for _ in range(2))
    return q, k, v, w_i, b_i, w_f, b_f


def ref_gates(q, k, v, w_i, b_i, w_f, b_f):
    x = torch.cat([q, k, v], dim=-1).double()
    return F.linear(x, w_i.double(), b_i.double()), F.linear(x, w_f.double(), b_f.double())


SHAPES = [  # B, S, D, NH
    (2, 400, 256, 4),      # cfg2 cell
    (1, 1600, 512, 4),     # cfg3 cell
    (1, 203, 1024, 4),     # two column slabs, ragged token tiles
    (2, 100, 512, 32),     # reference default qkv_block_size=16: eight output groups
    (1, 3, 64, 2),         # fewer tokens than a tile
    (2, 300, 256, 8),      # 16 gate outputs: the N = 48 operand of the tensor-core forward (mlstm_gates_tc.cu)
    (1, 129, 128, 6),      # 12 outputs padded to 16, one token past a tile
    (1, 256, 64, 8),       # whole tiles, a single 64-column slice per tensor
]


@pytest.mark.parametrize("B,S,D,NH", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gate_projection_forward(B, S, D, NH, dtype):
    from xlstm_yolo_b200 import ops
    q, k, v, w_i, b_i, w_f, b_f = make(B, S, D, NH, dtype, padded=(D == 256))
    i, f = ops.gate_proj_fwd_raw(q, k, v, w_i, b_i, w_f, b_f, NH)
    ri, rf = ref_gates(q, k, v, w_i, b_i, w_f, b_f)
    assert i.shape == (B, S, NH) and i.dtype == torch.float32
    assert rel(i, ri) < 2e-6 and rel(f, rf) < 2e-6       # inputs are exact in both; fp32 accumulation
    i2, f2 = ops.gate_proj_fwd_raw(q, k, v, w_i, None, w_f, None, NH)
    assert rel(i2, ri - b_i.double()) < 1e-5


@pytest.mark.parametrize("B,S,D,NH", SHAPES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gate_projection_backward_in_place(B, S, D, NH, dtype):
    from xlstm_yolo_b200 import ops
    # q,k,v as column slices of a wider buffer (row stride D + 8) while the gradients are dense (row stride D):
    # the two strides travel separately through the ABI (ld / ld_d)
    q, k, v, w_i, b_i, w_f, b_f = make(B, S, D, NH, dtype, seed=1, padded=(D in (256, 1024)))
    g = torch.Generator().manual_seed(5)
    di, df = (torch.randn(B, S, NH, generator=g).cuda() for _ in range(2))
    dq0, dk0, dv0 = (torch.randn(B, S, D, generator=g).to(dtype).cuda() for _ in range(3))
    dq, dk, dv = dq0.clone(), dk0.clone(), dv0.clone()
    dw_i, db_i, dw_f, db_f = ops.gate_proj_bwd_raw(q, k, v, w_i, w_f, NH, di, df, dq, dk, dv)
    x = torch.cat([q, k, v], dim=-1).double()
    dx = di.double() @ w_i.double() + df.double() @ w_f.double()
    # bf16: the updated gradient is rounded once per group of 8 gate outputs (NH=32 -> 8 passes)
    tol = (1e-2 if NH <= 4 else 2e-2) if dtype == torch.bfloat16 else 2e-6
    for got, base, sl in ((dq, dq0, slice(0, D)), (dk, dk0, slice(D, 2 * D)), (dv, dv0, slice(2 * D, 3 * D))):
        assert rel(got, base.double() + dx[..., sl]) < tol
    assert rel(dw_i, di.double().flatten(0, 1).T @ x.flatten(0, 1)) < 1e-5
    assert rel(dw_f, df.double().flatten(0, 1).T @ x.flatten(0, 1)) < 1e-5
    assert rel(db_i, di.double().sum((0, 1))) < 1e-5 and rel(db_f, df.double().sum((0, 1))) < 1e-5
    # deterministic: bit-identical on a second run
    dq2, dk2, dv2 = dq0.clone(), dk0.clone(), dv0.clone()
    again = ops.gate_proj_bwd_raw(q, k, v, w_i, w_f, NH, di, df, dq2, dk2, dv2)
    assert torch.equal(again[0], dw_i) and torch.equal(again[2], dw_f) and torch.equal(dq2, dq)


@pytest.mark.parametrize("dim,NH,S,dtype,reverse", [(256, 4, 400, torch.bfloat16, False), (512, 4, 300, torch.bfloat16, True),
                                                    (64, 4, 100, torch.float32, False)])
def test_fused_cell_matches_fp64_module(dim, NH, S, dtype, reverse):
    """MatrixLSTMCell through the fused CUDA node vs the same module (same rounded parameters and
    inputs) evaluated in fp64 on the CPU path (F.linear gates + native chunkwise form)."""
    import copy
    from xlstm_yolo_b200 import MatrixLSTMCell
    torch.manual_seed(0)
    cell = MatrixLSTMCell(dim=dim, num_heads=NH, chunk_size=64, reverse=reverse, use_autocast=(dtype != torch.float32))
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.05); cell.fgate.weight.normal_(0, 0.05); cell.igate.bias.normal_(0, 1.0)
        cell.outnorm.weight.normal_(0, 0.2)
    fused = cell.cuda().to(dtype)
    ref = copy.deepcopy(fused).cpu().double()
    g = torch.Generator().manual_seed(3)
    q, k, v = ((torch.randn(2, S, dim, generator=g) * s).to(dtype) for s in (dim ** -0.25, dim ** -0.25, 1.0))
    dy = torch.randn(2, S, dim, generator=g).to(dtype)
    outs, gate_out = [], {}
    plain_gates = ref._gates

    def keep_gate_grads(q_, k_, v_):   # the reference's gate gradients set the scale of the bias-grad noise
        i_, f_ = plain_gates(q_, k_, v_)
        i_.retain_grad(); f_.retain_grad()
        gate_out.update(igate=i_, fgate=f_)
        return i_, f_

    ref._gates = keep_gate_grads
    for m, dev, dt in ((fused, "cuda", dtype), (ref, "cpu", torch.float64)):
        leaves = [t.to(dev, dt).requires_grad_(True) for t in (q, k, v)]
        y = m(*leaves)
        y.backward(dy.to(dev, dt))
        outs.append((y.detach(), [t.grad for t in leaves], [m.igate.weight.grad, m.igate.bias.grad,
                                                            m.fgate.weight.grad, m.fgate.bias.grad]))
    (y1, gx1, gw1), (y2, gx2, gw2) = outs
    ty, tg = (1e-2, 2e-2) if dtype == torch.bfloat16 else (1e-4, 1e-3)
    assert all(t is not None for t in gw1)
    errs = {"y": rel(y1, y2), **{f"d{n}": rel(a, b) for n, a, b in zip("qkv", gx1, gx2)},
            **{n: rel(a, b) for n, a, b in zip(["dWi", "dbi", "dWf", "dbf"], gw1, gw2)}}
    # a bias gradient is a sum over all tokens of per-token gate gradients of both signs; its error is
    # bounded against sum_t |dgate_t| (the quantity the per-token tolerance applies to), not against itself
    for n, gname, got, want in (("dbi", "igate", gw1[1], gw2[1]), ("dbf", "fgate", gw1[3], gw2[3])):
        scale = gate_out[gname].grad.abs().sum((0, 2)).max()   # (B,NH,S) -> per head
        errs[n] = float((got.double().cpu() - want).abs().max() / scale)
    assert errs["y"] < ty, errs
    assert all(e < tg for n, e in errs.items() if n != "y"), errs


def test_fused_cell_under_fp16_autocast_scaler():
    from xlstm_yolo_b200 import MatrixLSTMCell
    torch.manual_seed(0)
    cell = MatrixLSTMCell(dim=256, num_heads=4, chunk_size=64).cuda()
    scaler = torch.amp.GradScaler("cuda")
    x = torch.randn(2, 400, 256, device="cuda") * 0.1
    with torch.autocast("cuda", dtype=torch.float16):
        y = cell(x, x, x)
        loss = y.float().pow(2).mean()
    scaler.scale(loss).backward()
    assert y.dtype in (torch.float16, torch.float32)
    for p in cell.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


@pytest.mark.parametrize("mode", ["chunk3", "padded"])
def test_fused_cell_backward_with_strided_qkv(mode):
    """q,k,v handed over as row-strided views (bf16 ``qkv.chunk(3, -1)`` or a padded ``[..., :D]`` slice) stay strided in
    the fused node; the backward must write the dense dq,dk,dv with their own row stride (single `ld`)."""
    from xlstm_yolo_b200 import MatrixLSTMCell
    torch.manual_seed(0)
    dim, NH, B, S = 256, 4, 2, 200
    cell = MatrixLSTMCell(dim=dim, num_heads=NH, chunk_size=64)
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.05); cell.fgate.weight.normal_(0, 0.05); cell.igate.bias.normal_(0, 1.0)
    cell = cell.cuda().to(torch.bfloat16)
    width = 3 * dim if mode == "chunk3" else dim + 8
    base = [(torch.randn(B, S, width, device="cuda") * 0.2).bfloat16().requires_grad_(True) for _ in range(1 if mode == "chunk3" else 3)]
    if mode == "chunk3":
        q, k, v = base[0].chunk(3, dim=-1)
    else:
        q, k, v = (t[..., :dim] for t in base)
    assert q.stride(1) != dim
    dy = torch.randn(B, S, dim, device="cuda").bfloat16()
    y = cell(q, k, v)
    y.backward(dy)
    got = [t.grad.clone() for t in base]
    dw = cell.igate.weight.grad.clone()
    for t in base:
        t.grad = None
    cell.zero_grad()
    # same inputs, contiguous
    qc, kc, vc = (t.detach().contiguous().requires_grad_(True) for t in (q, k, v))
    y2 = cell(qc, kc, vc)
    y2.backward(dy)
    assert torch.equal(y, y2)
    if mode == "chunk3":
        want = [torch.cat([qc.grad, kc.grad, vc.grad], dim=-1)]
    else:
        want = [torch.nn.functional.pad(t.grad, (0, 8)) for t in (qc, kc, vc)]
    for a, b in zip(got, want):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    assert torch.equal(dw, cell.igate.weight.grad)


def test_unsupported_dims_fall_back_instead_of_raising():
    """inner dim 768 (D / 256 = 3: no fused tail) and 2560 (gate weight tile > shared memory) run through the
    PyTorch tail / cuBLAS gates instead of raising UNSUPPORTED."""
    from xlstm_yolo_b200 import ops
    from xlstm_yolo_b200.vil import ViLLayer, SequenceTraversal
    assert not ops.gates_supported(2560) and ops.gates_supported(2048) and ops.gates_supported(768)
    torch.manual_seed(0)
    layer = ViLLayer(dim=384, direction=SequenceTraversal.ROWWISE_FROM_TOP_LEFT, qkv_block_size=64, chunk_size=64).cuda()
    x = torch.randn(1, 64, 384, device="cuda", requires_grad=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = layer(x)
    y.float().square().mean().backward()
    assert y.shape == x.shape and torch.isfinite(y).all() and torch.isfinite(x.grad).all()
    from xlstm_yolo_b200 import MatrixLSTMCell
    cell = MatrixLSTMCell(dim=2560, num_heads=20, chunk_size=64).cuda().to(torch.bfloat16)
    q = torch.randn(1, 64, 2560, device="cuda").bfloat16()
    assert torch.isfinite(cell(q, q, q)).all()
