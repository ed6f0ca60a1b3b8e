"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden vectors of the reference, and — at BASELINE sizes —
through size-independent properties.

Tolerances (BASELINE.json north star): outputs within max relative error 1e-2 (bf16 tcgen05
path) / 1e-4 (fp32 mode); gradients within 2e-2 (bf16) / 2e-4 (fp32).  "Relative" is
max|a-b| / max|b| (in the reference-init regime |h| ~ 1e-4, so element-wise
relative error is meaningless).
"""
import os

import numpy as np
import pytest
import torch

from oracle import mlstm_oracle as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {torch.float32: (1e-4, 2e-4), torch.bfloat16: (1e-2, 2e-2)}


def rel(a, b):
    b = b.double()
    return ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make(B, NH, S, DH, dtype, regime, qk_std=None, seed=0):
    """Inputs in the reference's memory layout: (B,S,NH,DH) storage viewed (B,NH,S,DH)."""
    g = torch.Generator().manual_seed(seed)
    qk_std = DH ** -0.5 if qk_std is None else qk_std
    q = (torch.randn(B, S, NH, DH, generator=g) * qk_std).to(dtype)
    k = (torch.randn(B, S, NH, DH, generator=g) * qk_std).to(dtype)
    v = torch.randn(B, S, NH, DH, generator=g).to(dtype)
    dh = torch.randn(B, S, NH, DH, generator=g).to(dtype)
    i = {"rand": torch.randn(B, S, NH, generator=g), "refinit": -10 + 0.1 * torch.randn(B, S, NH, generator=g),
         "forget": 2 * torch.randn(B, S, NH, generator=g)}[regime]
    f = torch.linspace(3, 6, NH).view(1, 1, NH) + torch.randn(B, S, NH, generator=g)
    if regime == "forget":
        f = f - 4
    return [x.transpose(1, 2) for x in (q, k, v, i, f, dh)]


def run_cuda(inputs, eps=1e-6, reverse=False, states=None, **kw):
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, dh = (x.cuda() for x in inputs)
    leaves = [x.detach().requires_grad_(True) for x in (q, k, v, i, f)]
    st = {} if states is None else {n: s.float().cuda() for n, s in states.items()}
    h = ops.mlstm(*leaves, eps=eps, reverse=reverse, **st, **kw)
    h.backward(dh)
    torch.cuda.synchronize()
    return [h.detach()] + [l.grad for l in leaves]


def near_tie_rows(inputs, reverse=False, states=None, eps=1e-6, width=2e-2):
    """Rows where |n_t| is within `width` of exp(-m_t).  The reference normaliser
    max(|n|, exp(-m)) (backends.py:249-252) makes dh/dn discontinuous there: any bf16-level
    perturbation of n flips the branch and changes that row's dq by O(1).  Such rows are a
    property of the reference's math, not of an implementation, and are excluded from the dq
    comparison (they are counted and must stay rare)."""
    from emu_kernel_dataflow import emu_forward
    a = [x.double() for x in inputs[:5]]
    if reverse:
        a = [x.flip(dims=[2]) for x in a]
    st = {} if states is None else dict(c0=states["c_initial"].double(), n0=states["n_initial"].double(),
                                        m0=states["m_initial"].double())
    _, nr, mr, _, _ = emu_forward(*a, L=64, eps=eps, **st)
    floor_ = torch.exp(-mr)
    tie = ((nr.abs() - floor_).abs() / floor_) < width
    return tie.flip(dims=[2]) if reverse else tie


def check(got, ref, dtype, what="", tie=None):
    th, tg = TOL[dtype]
    got, ref = list(got), list(ref)
    if tie is not None and tie.any():
        assert tie.float().mean().item() < 0.02, "too many near-tie rows for a meaningful comparison"
        keep = (~tie)[..., None]
        got[1], ref[1] = got[1].cpu() * keep, ref[1] * keep
        # a flipped row also leaks into dk and, through the suffix sum, into df of every earlier
        # position (the CPU emulation of the kernel dataflow shows the same 2.2e-2): relax those.
        tg = 2.5 * tg
    for name, a, b in zip(["h", "dq", "dk", "dv", "di", "df"], got, ref):
        assert torch.isfinite(a).all(), f"{what} {name} not finite"
        tol = th if name == "h" else tg
        err = rel(a, b)
        # gate gradients can be ~0 everywhere (e.g. S=1): accept a tiny absolute error too
        small = (a.double().cpu() - b.double()).abs().max().item() < 1e-5
        assert err < tol or (name in ("di", "df") and small), f"{what} {name}: rel err {err:.3e} > {tol}"


CASES = [
    # B, NH, S, DH, dtype, regime, reverse, eps
    (2, 2, 100, 16, torch.float32, "rand", False, 1e-6),
    (2, 2, 100, 16, torch.float32, "refinit", False, 5e-5),
    (2, 2, 100, 16, torch.float32, "forget", True, 1e-6),
    (2, 4, 400, 64, torch.float32, "rand", True, 1e-6),
    (1, 2, 300, 128, torch.float32, "refinit", False, 5e-5),
    (2, 8, 256, 16, torch.bfloat16, "rand", False, 1e-6),       # reference default qkv_block_size=16 -> SIMT
    (2, 4, 400, 64, torch.bfloat16, "rand", False, 1e-6),       # cfg2 shape (16-token tail)
    (2, 4, 400, 64, torch.bfloat16, "refinit", True, 5e-5),
    (2, 4, 400, 128, torch.bfloat16, "forget", True, 1e-6),
    (2, 4, 400, 128, torch.bfloat16, "rand", False, 1e-6),
    (1, 4, 1600, 128, torch.bfloat16, "rand", False, 1e-6),     # cfg3 shape
    (1, 4, 1600, 128, torch.bfloat16, "refinit", True, 5e-5),
    (1, 2, 6400, 128, torch.bfloat16, "rand", True, 1e-6),      # 80x80 map: df stays anchored over 50 chunks
    (1, 2, 3200, 64, torch.bfloat16, "forget", False, 1e-6),
    (1, 2, 6400, 64, torch.bfloat16, "rand", False, 1e-6),      # 50 chunks at DH 64 (fused-walk pin: df carried over the whole walk)
    (3, 2, 129, 64, torch.bfloat16, "rand", False, 1e-6),
    (3, 2, 127, 128, torch.bfloat16, "rand", True, 1e-6),
    (2, 2, 1, 64, torch.bfloat16, "rand", False, 1e-6),
    (2, 2, 1, 16, torch.float32, "rand", False, 1e-6),
    (1, 2, 200, 256, torch.bfloat16, "rand", False, 1e-6),      # secondary-reading head dim: value-sliced SIMT
    (1, 2, 200, 256, torch.bfloat16, "refinit", True, 5e-5),
    (1, 2, 70, 256, torch.float32, "rand", True, 1e-6),
    (1, 2, 96, 192, torch.float32, "forget", False, 1e-6),      # three slices
]


@pytest.mark.parametrize("B,NH,S,DH,dtype,regime,reverse,eps", CASES)
def test_cuda_matches_oracle(B, NH, S, DH, dtype, regime, reverse, eps):
    inputs = make(B, NH, S, DH, dtype, regime)
    ref = O.mlstm_fwbw(*(x.double() for x in inputs), chunk_size=64, eps=eps, reverse=reverse)
    got = run_cuda(inputs, eps=eps, reverse=reverse)
    tie = near_tie_rows(inputs, reverse, eps=eps) if dtype == torch.bfloat16 else None
    check(got, ref, dtype, f"B{B} NH{NH} S{S} DH{DH} {regime} rev={reverse}", tie)


def siging_tie_rows(inputs, reverse=False, width=2e-2):
    """Rows where |n_t| is within `width` of the sigmoid-gate normaliser's floor 1 (same discontinuity)."""
    q, k, v, i, f = (x.double() for x in inputs[:5])
    if reverse:
        q, k, v, i, f = (x.flip(dims=[2]) for x in (q, k, v, i, f))
    S, DH = q.shape[2], q.shape[3]
    b = torch.nn.functional.logsigmoid(f).cumsum(-1)
    logD = b[..., :, None] - b[..., None, :] + torch.nn.functional.logsigmoid(i)[..., None, :]
    D = torch.exp(logD.masked_fill(~torch.ones(S, S, dtype=torch.bool).tril(), -float("inf")))
    n = ((q @ k.transpose(-1, -2)) * DH ** -0.5 * D).sum(-1)
    tie = (n.abs() - 1).abs() < width
    return tie.flip(dims=[2]) if reverse else tie


SIG_CASES = [
    # B, NH, S, DH, dtype, regime, reverse   (PARITY UNPINNED: restated algorithm, see oracle/mlstm_oracle.py)
    (2, 4, 400, 64, torch.bfloat16, "rand", False),
    (2, 4, 400, 128, torch.bfloat16, "rand", True),
    (1, 4, 1600, 128, torch.bfloat16, "rand", False),
    (1, 2, 700, 64, torch.bfloat16, "forget", True),
    (3, 2, 129, 64, torch.bfloat16, "refinit", False),
    (2, 2, 100, 16, torch.float32, "rand", True),
    (1, 2, 300, 128, torch.float32, "rand", False),
    (1, 2, 200, 256, torch.bfloat16, "rand", False),
]


@pytest.mark.parametrize("B,NH,S,DH,dtype,regime,reverse", SIG_CASES)
def test_sigmoid_input_gate_matches_oracle(B, NH, S, DH, dtype, regime, reverse):
    inputs = make(B, NH, S, DH, dtype, regime)
    ref = O.mlstm_fwbw(*(x.double() for x in inputs), eps=1e-6, reverse=reverse, input_gate="sigmoid")
    got = run_cuda(inputs, reverse=reverse, input_gate="sigmoid")
    tie = siging_tie_rows(inputs, reverse) if dtype == torch.bfloat16 else None
    check(got, ref, dtype, f"siging B{B} NH{NH} S{S} DH{DH} {regime} rev={reverse}", tie)


@pytest.mark.parametrize("dtype,DH", [(torch.float32, 32), (torch.bfloat16, 64), (torch.bfloat16, 128)])
def test_sigmoid_input_gate_states(dtype, DH):
    from xlstm_yolo_b200 import ops
    B, NH, S = 2, 2, 300
    inputs = make(B, NH, S, DH, dtype, "rand")
    g = torch.Generator().manual_seed(7)
    C0, n0 = torch.randn(B, NH, DH, DH, generator=g), torch.randn(B, NH, DH, generator=g)
    q, k, v, i, f, _ = (x.cuda() for x in inputs)
    h, (C, n, m) = ops.mlstm(q, k, v, i, f, C0.cuda(), n0.cuda(), torch.randn(B, NH, 1).cuda(), return_last_states=True,
                             input_gate="sigmoid")   # m_initial is ignored by the sigmoid gate
    hr, (Cr, nr, _) = O.mlstm_siging_recurrent(*(x.double() for x in inputs[:5]), C0.double(), n0.double(),
                                               return_last_states=True)
    th = TOL[dtype][0]
    assert rel(h, hr) < th and rel(C, Cr) < th and rel(n, nr) < th and float(m.abs().max()) == 0.0


@pytest.mark.parametrize("DH", [64, 128])
def test_long_sequence_inference_forward(DH):
    """BASELINE configs[4] regime: 1280x1280 input -> 160x160 P3 map = 25600 tokens, forward only (no saved rows)."""
    from xlstm_yolo_b200 import ops
    B, NH, S = 1, 2, 25600
    q, k, v, i, f, _ = make(B, NH, S, DH, torch.bfloat16, "rand")
    with torch.no_grad():
        h = ops.mlstm(q.cuda(), k.cuda(), v.cuda(), i.cuda(), f.cuda(), eps=1e-6)
        hr = ops.mlstm(q.cuda(), k.cuda(), v.cuda(), i.cuda(), f.cuda(), eps=1e-6, reverse=True)
    ref = O.mlstm_chunkwise(*(x.double() for x in (q, k, v, i, f)), chunk_size=64, eps=1e-6)
    refr = O.mlstm_chunkwise(*(x.double() for x in (q, k, v, i, f)), chunk_size=64, eps=1e-6, reverse=True)
    assert torch.isfinite(h).all() and rel(h, ref) < 1e-2 and rel(hr, refr) < 1e-2


def test_kernel_family_dispatch():
    from xlstm_yolo_b200 import ops
    bf = lambda d: torch.empty(1, 1, 8, d, dtype=torch.bfloat16, device="cuda")
    assert ops.kernel_family(bf(64), bf(64)) == "tcgen05"
    assert ops.kernel_family(bf(128), bf(128)) == "tcgen05"
    assert ops.kernel_family(bf(16), bf(16)) == "simt"
    assert ops.kernel_family(bf(128).float(), bf(128).float()) == "simt"
    assert ops.kernel_family(bf(256), bf(256)) == "simt"


@pytest.mark.parametrize("dtype,DH", [(torch.float32, 32), (torch.bfloat16, 64), (torch.bfloat16, 128),
                                      (torch.float32, 256)])
@pytest.mark.parametrize("reverse", [False, True])
def test_initial_and_last_states(dtype, DH, reverse):
    B, NH, S = 2, 2, 300
    inputs = make(B, NH, S, DH, dtype, "rand")
    g = torch.Generator().manual_seed(7)
    st = dict(c_initial=torch.randn(B, NH, DH, DH, generator=g), n_initial=torch.randn(B, NH, DH, generator=g),
              m_initial=torch.randn(B, NH, 1, generator=g))
    ref = O.mlstm_fwbw(*(x.double() for x in inputs), chunk_size=64, eps=1e-6, reverse=reverse,
                       **{n: s.double() for n, s in st.items()})
    got = run_cuda(inputs, reverse=reverse, states=st)
    check(got, ref, dtype, "states", near_tie_rows(inputs, reverse, st) if dtype == torch.bfloat16 else None)
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, _ = (x.cuda() for x in inputs)
    h, (C, n, m) = ops.mlstm(q, k, v, i, f, *(s.cuda() for s in st.values()), return_last_states=True, reverse=reverse)
    _, (Cr, nr, mr) = O.mlstm_chunkwise(*(x.double() for x in inputs[:5]), *(s.double() for s in st.values()),
                                        chunk_size=64, return_last_states=True, reverse=reverse)
    th = TOL[dtype][0]
    assert rel(C, Cr) < th and rel(n, nr) < th and rel(m, mr) < 1e-4


SEQ = ["seq_randgate_s64_dh16", "seq_refinit_s128_dh32", "seq_strongforget_s96_dh16"]


@pytest.mark.parametrize("name", SEQ)
def test_golden_reference_vectors_fp32(name):
    """CUDA fp32 mode against outputs the REFERENCE's chunkwise_simple produced (tests/golden)."""
    z = np.load(os.path.join(GOLD, name + ".npz"))
    t = lambda k: torch.from_numpy(z[k]).float()
    inputs = [t(k) for k in ("q", "k", "v", "i", "f", "dh")]
    got = run_cuda(inputs, eps=float(z["eps"]))
    ref = [torch.from_numpy(z[k]) for k in ("h_chunkwise", "dq", "dk", "dv", "di", "df")]
    check(got, ref, torch.float32, name)
    assert rel(got[0], torch.from_numpy(z["h_recurrent"])) < 1e-4


def test_golden_states_fp32():
    z = np.load(os.path.join(GOLD, "states_s48_dh16.npz"))
    from xlstm_yolo_b200 import ops
    t = lambda k: torch.from_numpy(z[k]).float().cuda()
    h, (C, n, m) = ops.mlstm(t("q"), t("k"), t("v"), t("i"), t("f"), t("c_initial"), t("n_initial"), t("m_initial"),
                             return_last_states=True, eps=float(z["eps"]))
    for a, k in ((h, "h"), (C, "c_last"), (n, "n_last"), (m, "m_last")):
        assert rel(a, torch.from_numpy(z[k])) < 1e-4, k


def test_module_matches_oracle_cell_and_vendored_golden():
    from xlstm_yolo_b200 import MatrixLSTMCell
    torch.manual_seed(0)
    cell = MatrixLSTMCell(dim=256, num_heads=4, chunk_size=64)
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.02)
        cell.fgate.weight.normal_(0, 0.02)
        cell.igate.bias.fill_(-2.0)
        cell.outnorm.weight.normal_(0, 0.1)
    q, k, v = (torch.randn(4, 400, 256) * s for s in (0.125, 0.125, 1.0))
    want = O.cell_forward(q.double(), k.double(), v.double(), 4, *(p.double() for p in (
        cell.igate.weight, cell.igate.bias, cell.fgate.weight, cell.fgate.bias, cell.outnorm.weight, cell.outnorm.bias)),
        chunk_size=64, eps=5e-5)
    cell = cell.cuda()
    import copy
    for dt, tol in ((torch.float32, 2e-2), (torch.bfloat16, 3e-2)):   # kernels run in bf16 (autocast_kernel_dtype)
        mod = cell if dt == torch.float32 else copy.deepcopy(cell).to(dt)   # e.g. the .half() val path
        y = mod(q.cuda().to(dt), k.cuda().to(dt), v.cuda().to(dt))
        assert y.shape == (4, 400, 256) and y.dtype == dt
        assert rel(y.float(), want) < tol
    # fp32-mode cell against the vendored mLSTMCell golden (parallel form; differs only by the m_0 floor)
    z = np.load(os.path.join(GOLD, "cell_vendored_s32_h64.npz"))
    c32 = MatrixLSTMCell(dim=64, num_heads=int(z["num_heads"]), norm_bias=False, use_autocast=False)
    c32.outnorm.eps = float(z["norm_eps"])
    c32.gpu_backend.config.eps = c32.gpu_backend_infer.config.eps = float(z["eps"])
    with torch.no_grad():
        for n_, key in (("igate.weight", "igate_w"), ("igate.bias", "igate_b"), ("fgate.weight", "fgate_w"),
                        ("fgate.bias", "fgate_b"), ("outnorm.weight", "outnorm_w")):
            dict(c32.named_parameters())[n_].copy_(torch.from_numpy(z[key]).float())
    c32 = c32.cuda()
    y = c32(*(torch.from_numpy(z[k_]).float().cuda() for k_ in ("q", "k", "v")))
    assert rel(y, torch.from_numpy(z["y"])) < 1e-3


def test_module_trains_under_fp16_autocast():
    from xlstm_yolo_b200 import MatrixLSTMCell
    torch.manual_seed(0)
    cell = MatrixLSTMCell(dim=128, num_heads=2, chunk_size=64).cuda()
    q, k, v = (torch.randn(2, 200, 128, device="cuda", requires_grad=True) for _ in range(3))
    scaler = torch.amp.GradScaler("cuda")
    with torch.autocast("cuda", dtype=torch.float16):
        y = cell(q * 0.1, k * 0.1, v)
        loss = y.float().square().mean()
    scaler.scale(loss).backward()
    for t in (q, k, v, cell.igate.bias, cell.fgate.bias, cell.outnorm.weight):
        assert t.grad is not None and torch.isfinite(t.grad).all()
    assert cell.fgate.bias.grad.abs().sum() > 0


# ---- size-independent properties at BASELINE.json sizes ---------------------------------------

FULL = [(32, 4, 400, 64), (32, 4, 1600, 128)]


@pytest.mark.parametrize("B,NH,S,DH", FULL)
def test_full_size_state_carry_and_reverse_properties(B, NH, S, DH):
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, _ = (x.cuda() for x in make(B, NH, S, DH, torch.bfloat16, "rand", seed=3))
    h = ops.mlstm(q, k, v, i, f)
    assert torch.isfinite(h).all()
    # (1) determinism: bit-identical on a second run
    assert torch.equal(h, ops.mlstm(q, k, v, i, f))
    # (2) splitting the sequence and carrying (C, n, m) equals one pass
    s0 = (S // 2 // 8) * 8 + 8
    h1, st = ops.mlstm(q[:, :, :s0], k[:, :, :s0], v[:, :, :s0], i[:, :, :s0], f[:, :, :s0], return_last_states=True)
    h2 = ops.mlstm(q[:, :, s0:], k[:, :, s0:], v[:, :, s0:], i[:, :, s0:], f[:, :, s0:], *st)
    assert rel(torch.cat([h1, h2], 2).float(), h.float().cpu()) < 1e-2
    # (3) reverse scan == flip -> cell -> flip (the reference's ViLLayer, vision_lstm2.py:479-480,505-506)
    fl = lambda x: x.flip(dims=[2])
    hr = ops.mlstm(q, k, v, i, f, reverse=True)
    hf = fl(ops.mlstm(fl(q), fl(k), fl(v), fl(i), fl(f)))
    assert rel(hr.float(), hf.float().cpu()) < 1e-2
    # (4) linearity in v (exact for a power-of-two factor)
    assert torch.equal(ops.mlstm(q, k, v * 2, i, f), h * 2)


@pytest.mark.parametrize("B,NH,S,DH", FULL)
def test_full_size_gradient_identities(B, NH, S, DH):
    """di = k.dk and sum_t(q.dq - k.dk) = 0 hold for the true gradient (zero initial state)."""
    inputs = make(B, NH, S, DH, torch.bfloat16, "rand", seed=5)
    h, dq, dk, dv, di, df = run_cuda(inputs)
    q, k = inputs[0].cuda().float(), inputs[1].cuda().float()
    kdk = (k * dk.float()).sum(-1)
    assert rel(di, kdk.cpu()) < 2e-2
    tot = ((q * dq.float()).sum(-1) - kdk).sum(-1)
    scale = (q * dq.float()).sum(-1).abs().sum(-1)
    assert (tot.abs() / scale.clamp_min(1e-6)).max().item() < 2e-2
    for t in (h, dq, dk, dv, di, df):
        assert torch.isfinite(t).all()


def test_empty_and_unaligned_inputs():
    from xlstm_yolo_b200 import ops
    z = torch.empty(0, 2, 16, 64, dtype=torch.bfloat16, device="cuda")
    g = torch.empty(0, 2, 16, device="cuda")
    assert ops.mlstm(z, z, z, g, g).shape == (0, 2, 16, 64)
    # contiguous (B,NH,S,DH) and odd views are accepted (copied into an aligned layout if needed)
    inputs = make(1, 2, 200, 64, torch.bfloat16, "rand")
    ref = O.mlstm_chunkwise(*(x.double() for x in inputs[:5]), chunk_size=64)
    q, k, v, i, f = (x.cuda() for x in inputs[:5])
    h1 = ops.mlstm(q.contiguous(), k.contiguous(), v.contiguous(), i.contiguous(), f.contiguous())
    big = torch.zeros(1, 2, 200, 72, dtype=torch.bfloat16, device="cuda")
    big[..., 4:68] = q
    h2 = ops.mlstm(big[..., 4:68], k, v, i, f)          # misaligned base pointer -> internal copy
    assert rel(h1.float(), ref) < 1e-2 and rel(h2.float(), ref) < 1e-2


def test_unsupported_head_dim_fails_loudly():
    from xlstm_yolo_b200 import ops
    x = torch.zeros(1, 1, 8, 512, dtype=torch.bfloat16, device="cuda")
    g = torch.zeros(1, 1, 8, device="cuda")
    with pytest.raises(RuntimeError, match="UNSUPPORTED"):
        ops.mlstm(x, x, x, g, g)


def test_forward_backward_capture_into_cuda_graph():
    """The C-ABI calls are plain launches on the caller's stream (no allocation, no synchronisation): a forward + backward
    captured into a CUDA graph and replayed gives the eager results bit for bit."""
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, dh = (x.cuda() for x in make(20, 4, 300, 64, torch.bfloat16, "rand", seed=11))
    pl = ops.MLSTMPlan(q, k, v, i, f, dh)
    pl.forward(); pl.backward()
    torch.cuda.synchronize()
    want = [t.clone() for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df)]
    for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df):
        t.zero_()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            pl.forward(); pl.backward()
    torch.cuda.current_stream().wait_stream(cap)
    for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df):
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip((pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df), want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("reverse", [False, True])
def test_full_size_backward_is_deterministic(reverse):
    """The wide-batch kernels (single-pass forward, fused single-walk backward: products issued from several lanes, warps
    running ahead of the control warp, barrier-free hand-offs through mbarriers) give bit-identical results on every run —
    a missing ordering between a producer and a consumer shows up here as run-to-run differences."""
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, dh = (x.cuda() for x in make(32, 4, 400, 64, torch.bfloat16, "rand", seed=9))
    pl = ops.MLSTMPlan(q, k, v, i, f, dh, reverse=reverse)
    assert pl.variant_bwd == "fused_walk"
    pl.forward(); pl.backward()
    torch.cuda.synchronize()
    want = [t.clone() for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df)]
    for t in want:
        assert torch.isfinite(t).all()
    for _ in range(25):
        for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df):
            t.fill_(float("nan"))
        pl.forward(); pl.backward()
        torch.cuda.synchronize()
        for a, b in zip((pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df), want):
            assert torch.equal(a, b)
