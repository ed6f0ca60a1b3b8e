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


KINK_STATS = []   # (what, fraction of rows within 2 % of the kink, rows whose side the kernel flipped)


def oracle_on_kernel_side(inputs, eps=1e-6, reverse=False, states=None, input_gate="exp", batch_rows=None, what=""):
    """fp64 oracle (h, dq, dk, dv, di, df) for a bf16 kernel run.

    The reference normaliser max(|n_t|, exp(-m_t)) (backends.py:249-252) has a kink at |n_t| = exp(-m_t): dh/dn jumps
    there, so for a row within rounding distance of it the gradient is whichever side the implementation's own
    rounding of n_t picks (the reference's fp32 rounding included), and a flipped row changes its dq by O(1) and leaks
    into dk / di / df of every earlier key.  Instead of masking such rows and relaxing the tolerance for the rest
    (round 1), the oracle evaluates the one-sided derivative on the side the kernel's saved n_t, m_t landed on — for
    the rows within 2 % of the kink only; every other row uses the oracle's own side — and then EVERY row of every
    gradient is held to the north-star tolerance.  The fraction of near-kink rows and the number of flipped rows are
    recorded (printed by test_zz_kink_report)."""
    from xlstm_yolo_b200 import ops
    dbl = [x.double() for x in inputs]
    kw = {} if states is None else {n: s.double() for n, s in states.items()}
    q, k, v, i, f = (x.cuda() for x in inputs[:5])
    st = [None, None, None] if states is None else [states["c_initial"].float().cuda().contiguous(),
                                                     states["n_initial"].float().cuda().contiguous(),
                                                     states["m_initial"].float().reshape(q.shape[0], q.shape[1]).cuda().contiguous()]
    _, n_g, m_g, _, _ = ops.mlstm_fwd_raw(ops._prep_act(q), ops._prep_act(k), ops._prep_act(v), i.float(), f.float(), *st,
                                          eps=eps, reverse=reverse, gate_mode=ops.gate_mode_of(input_gate))
    n_g, m_g = n_g.double().cpu(), m_g.double().cpu()
    if batch_rows is not None:
        dbl = [x[batch_rows] for x in dbl]
        kw = {n: s[batch_rows] for n, s in kw.items()}
        n_g, m_g = n_g[batch_rows], m_g[batch_rows]
    with torch.no_grad():
        if input_gate == "sigmoid":
            _, (n_o, m_o) = O.mlstm_siging_parallel(*dbl[:5], eps=eps, reverse=reverse, return_rows=True)
        else:
            _, (n_o, m_o) = O.mlstm_chunkwise(*dbl[:5], chunk_size=64, eps=eps, reverse=reverse, return_rows=True, **kw)
    side_o, near = O.kink_rows(n_o, m_o)
    side_g, _ = O.kink_rows(n_g, m_g)
    side = torch.where(near, side_g, side_o)
    KINK_STATS.append((what, near.float().mean().item(), int((side != side_o).sum())))
    assert near.float().mean().item() < 0.02, "too many near-kink rows for a meaningful comparison"
    return O.mlstm_fwbw(*dbl, chunk_size=64, eps=eps, reverse=reverse, input_gate=input_gate, kink_side=side, **kw)


def check(got, ref, dtype, what=""):
    th, tg = TOL[dtype]
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
    (2, 8, 256, 16, torch.bfloat16, "rand", False, 1e-6),       # reference default qkv_block_size=16 -> tcgen05, padded to 64
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
    (1, 2, 200, 256, torch.bfloat16, "rand", False, 1e-6),      # secondary-reading head dim: slice-streaming tcgen05 family
    (1, 2, 200, 256, torch.bfloat16, "refinit", True, 5e-5),
    (2, 2, 400, 256, torch.bfloat16, "forget", True, 1e-6),
    (3, 2, 129, 256, torch.bfloat16, "rand", True, 1e-6),
    (1, 4, 1600, 256, torch.bfloat16, "rand", False, 1e-6),     # cfg3 secondary reading (d = 512, expansion 2 -> DH 256)
    (1, 2, 3200, 256, torch.bfloat16, "rand", True, 1e-6),
    (2, 2, 1, 256, torch.bfloat16, "rand", False, 1e-6),        # a single token: one item per output half, no state step
    (1, 1, 128, 256, torch.bfloat16, "refinit", True, 5e-5),    # exactly one full chunk, one (batch, head): a 2-CTA grid
    (1, 8, 300, 256, torch.bfloat16, "forget", False, 1e-6),
    (1, 2, 70, 256, torch.float32, "rand", True, 1e-6),
    (1, 2, 96, 192, torch.float32, "forget", False, 1e-6),      # three slices
]


@pytest.mark.parametrize("B,NH,S,DH,dtype,regime,reverse,eps", CASES)
def test_cuda_matches_oracle(B, NH, S, DH, dtype, regime, reverse, eps):
    inputs = make(B, NH, S, DH, dtype, regime)
    what = f"B{B} NH{NH} S{S} DH{DH} {regime} rev={reverse}"
    if dtype == torch.bfloat16:
        ref = oracle_on_kernel_side(inputs, eps=eps, reverse=reverse, what=what)
    else:
        ref = O.mlstm_fwbw(*(x.double() for x in inputs), chunk_size=64, eps=eps, reverse=reverse)
    got = run_cuda(inputs, eps=eps, reverse=reverse)
    check(got, ref, dtype, what)


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
    what = f"siging B{B} NH{NH} S{S} DH{DH} {regime} rev={reverse}"
    if dtype == torch.bfloat16:
        ref = oracle_on_kernel_side(inputs, reverse=reverse, input_gate="sigmoid", what=what)
    else:
        ref = O.mlstm_fwbw(*(x.double() for x in inputs), eps=1e-6, reverse=reverse, input_gate="sigmoid")
    got = run_cuda(inputs, reverse=reverse, input_gate="sigmoid")
    check(got, ref, dtype, what)


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
    assert ops.kernel_family(bf(16), bf(16)) == "tcgen05"      # zero-padded to 64 inside the library (mlstm_api.cu)
    assert ops.kernel_family(bf(192), bf(192)) == "tcgen05"    # ... to 256
    assert ops.kernel_family(bf(8), bf(8)) == "simt"
    assert ops.kernel_family(bf(16).float(), bf(16).float()) == "simt"
    assert ops.kernel_family(bf(128).float(), bf(128).float()) == "simt"
    assert ops.kernel_family(bf(256), bf(256)) == "tcgen05"
    assert ops.kernel_family(bf(256).float(), bf(256).float()) == "simt"


@pytest.mark.parametrize("dtype,DH", [(torch.float32, 32), (torch.bfloat16, 64), (torch.bfloat16, 128),
                                      (torch.float32, 256), (torch.bfloat16, 256)])
@pytest.mark.parametrize("reverse", [False, True])
def test_initial_and_last_states(dtype, DH, reverse):
    B, NH, S = 2, 2, 300
    inputs = make(B, NH, S, DH, dtype, "rand")
    g = torch.Generator().manual_seed(7)
    st = dict(c_initial=torch.randn(B, NH, DH, DH, generator=g), n_initial=torch.randn(B, NH, DH, generator=g),
              m_initial=torch.randn(B, NH, 1, generator=g))
    if dtype == torch.bfloat16:
        ref = oracle_on_kernel_side(inputs, reverse=reverse, states=st, what=f"states DH{DH} rev={reverse}")
    else:
        ref = O.mlstm_fwbw(*(x.double() for x in inputs), chunk_size=64, eps=1e-6, reverse=reverse,
                           **{n: s.double() for n, s in st.items()})
    got = run_cuda(inputs, reverse=reverse, states=st)
    check(got, ref, dtype, "states")
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

FULL = [(32, 4, 400, 64), (32, 4, 1600, 128), (16, 4, 1600, 256)]


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


@pytest.mark.parametrize("B,NH,S,DH,reverse", [(32, 4, 400, 64, False), (32, 4, 400, 64, True), (32, 4, 1600, 128, False),
                                                (32, 4, 1600, 128, True), (8, 4, 1600, 128, False), (32, 4, 1600, 256, False),
                                                (8, 4, 1600, 256, True)])
def test_full_size_batch_rows_match_oracle(B, NH, S, DH, reverse):
    """The BASELINE-size launches themselves (the wide-batch kernels bench.py times, and the per-GPU DDP shape) against
    the fp64 oracle on batch rows 0, mid and last — h and all five gradients, north-star tolerances: a (batch, head)
    indexing or stride bug that is self-consistent under the property tests above shows up here."""
    inputs = make(B, NH, S, DH, torch.bfloat16, "rand", seed=21)
    rows = sorted({0, (17 * B) // 32, B - 1})
    got = [t[rows] for t in run_cuda(inputs, reverse=reverse)]
    what = f"full B{B} NH{NH} S{S} DH{DH} rev={reverse} rows {rows}"
    ref = oracle_on_kernel_side(inputs, reverse=reverse, batch_rows=rows, what=what)
    check(got, ref, torch.bfloat16, what)


REF_CONFIGS = {   # the four configs MatrixLSTMCell builds at HEAD, vision_lstm2.py:819-877 (train and infer differ in `mode` only)
    "cpu": dict(chunkwise_kernel="chunkwise--native_autograd", sequence_kernel="native_sequence__native", step_kernel="native"),
    "gpu": dict(chunkwise_kernel="chunkwise--triton_xl_chunk_siging", sequence_kernel="native_sequence__triton", step_kernel="triton"),
}


@pytest.mark.parametrize("which,mode", [("cpu", "train"), ("cpu", "inference"), ("gpu", "train"), ("gpu", "inference")])
def test_backend_built_from_the_references_own_config_strings(which, mode):
    """Through the import shim, with the literal keyword arguments of vision_lstm2.py:819-877.  Documents which gate
    arithmetic each string selects here: "chunkwise--native_autograd" -> exponential input gate with max-stabiliser
    (backends.py:149-263; the oracle, parity pinned), "chunkwise--triton_xl_chunk_siging" -> sigmoid input gate
    (upstream's published form; parity UNPINNED, oracle/mlstm_oracle.py) — on CUDA and on CPU tensors alike."""
    from xlstm_yolo_b200 import compat
    compat.install()
    from mlstm_kernels.torch.backend_module import mLSTMBackend, mLSTMBackendConfig
    cfg = mLSTMBackendConfig(**REF_CONFIGS[which], chunk_size=64, autocast_kernel_dtype="bfloat16",
                             return_last_states=False, mode=mode, eps=5e-5)
    be = mLSTMBackend(config=cfg)
    gate = {"cpu": "exp", "gpu": "sigmoid"}[which]
    assert be.input_gate == gate and cfg.input_gate == gate
    inputs = make(2, 4, 400, 64, torch.bfloat16, "rand", seed=4)
    q, k, v, i, f, dh = inputs
    ref = oracle_on_kernel_side(inputs, eps=5e-5, input_gate=gate, what=f"refcfg {which} {mode}")
    leaves = [x.cuda().detach().requires_grad_(True) for x in (q, k, v, i, f)]
    h = be(q=leaves[0], k=leaves[1], v=leaves[2], i=leaves[3], f=leaves[4])
    h.backward(dh.cuda())
    check([h.detach()] + [l.grad for l in leaves], ref, torch.bfloat16, f"refcfg {which} {mode}")
    # the same backend object on CPU tensors (the constructor-time stride probe, nn/tasks.py:353-362) uses the same gate
    hc = be(q=q.float(), k=k.float(), v=v.float(), i=i, f=f)
    assert rel(hc, ref[0]) < 1e-4


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
def test_full_size_backward_is_deterministic_dh128(reverse):
    """Same for the DH = 128 fused walk (mlstm_tc_bwd_fused128.cu: TMEM regions and the Cs tile change hands several times per
    step between the compute warps, the control warp's MMAs and the TMA proxies) at the north-star shape."""
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, dh = (x.cuda() for x in make(32, 4, 1600, 128, torch.bfloat16, "rand", seed=9))
    pl = ops.MLSTMPlan(q, k, v, i, f, dh, reverse=reverse)
    assert pl.variant_bwd == "fused_walk"
    pl.forward(); pl.backward()
    torch.cuda.synchronize()
    want = [t.clone() for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df)]
    for t in want:
        assert torch.isfinite(t).all()
    for _ in range(10):
        for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df):
            t.fill_(float("nan"))
        pl.forward(); pl.backward()
        torch.cuda.synchronize()
        for a, b in zip((pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df), want):
            assert torch.equal(a, b)


@pytest.mark.parametrize("reverse", [False, True])
def test_full_size_dh256_family_is_deterministic(reverse):
    """The slice-streaming DH = 256 family (mlstm_tc_256.cu): ring slots change hands between TMA, two MMA lanes and the
    next item's prefetch; the staged output tile doubles as the MMA2 operand buffer.  Bit-identical NaN-poisoned reruns."""
    from xlstm_yolo_b200 import ops
    q, k, v, i, f, dh = (x.cuda() for x in make(16, 4, 1600, 256, torch.bfloat16, "rand", seed=9))
    pl = ops.MLSTMPlan(q, k, v, i, f, dh, reverse=reverse)
    assert pl.family == "tcgen05" and pl.variant_bwd == "chunk_parallel"
    pl.forward(); pl.backward()
    torch.cuda.synchronize()
    want = [t.clone() for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df)]
    for t in want:
        assert torch.isfinite(t).all()
    for _ in range(8):
        for t in (pl.h, pl.dq, pl.dk, pl.dv, pl.di, pl.df):
            t.fill_(float("nan"))
        pl.forward(); pl.backward()
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


def test_zz_kink_report(capsys):
    """Not a check: prints how often the near-kink rule of oracle_on_kernel_side was in play in this session."""
    with capsys.disabled():
        tot = len(KINK_STATS)
        hit = [s for s in KINK_STATS if s[2] > 0]
        print(f"\n[kink] {tot} bf16 oracle comparisons; {len(hit)} had rows whose side of max(|n|, e^-m) the kernel flipped")
        for what, frac, flipped in KINK_STATS:
            if frac > 0:
                print(f"[kink]   {what}: {100 * frac:.3f} % of rows within 2 % of the kink, {flipped} flipped")


def test_cell_and_layer_trace_under_torch_compile():
    """The reference compiles its model (debug.py:13).  The cell is registered with torch.library (xlstm_yolo_b200::mlstm_fwd /
    ::mlstm_bwd, fake implementations + autograd formula), so Dynamo + AOT autograd trace a module that contains it into ONE
    graph; results equal eager mode.  backend="aot_eager": tracing, functionalisation and the fake kernels are what is under
    test, not a code generator."""
    from xlstm_yolo_b200 import MatrixLSTMCell, ViLBlockPair
    torch.manual_seed(0)
    cell = MatrixLSTMCell(dim=256, num_heads=4, chunk_size=64)
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.05); cell.fgate.weight.normal_(0, 0.05); cell.igate.bias.normal_(0, 1.0)
    cell = cell.cuda().to(torch.bfloat16)
    cell.fused_gates = False      # eager and compiled then run the same decomposition (F.linear gates + the cell op)
    comp = torch.compile(cell, backend="aot_eager", fullgraph=True)
    g = torch.Generator().manual_seed(1)
    base = [(torch.randn(2, 400, 256, generator=g) * s).bfloat16().cuda() for s in (0.1, 0.1, 1.0)]
    dy = torch.randn(2, 400, 256, generator=g).bfloat16().cuda()
    res = []
    for m in (cell, comp):
        leaves = [t.clone().requires_grad_(True) for t in base]
        cell.zero_grad()
        y = m(*leaves)
        y.backward(dy)
        res.append([y.detach()] + [t.grad for t in leaves] + [cell.fgate.bias.grad.clone()])
    for a, b in zip(*res):
        assert torch.isfinite(a).all() and torch.equal(a, b)
    assert hasattr(torch.ops.xlstm_yolo_b200, "mlstm_fwd") and hasattr(torch.ops.xlstm_yolo_b200, "mlstm_bwd")
    # the whole bidirectional pair (graph breaks allowed: the depthwise conv view tricks are the compiler's business)
    pair = ViLBlockPair(dim=128, chunk_size=64, qkv_block_size=64).cuda()
    cpair = torch.compile(pair, backend="aot_eager")
    x = torch.randn(2, 400, 128, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ye, yc = pair(x), cpair(x)
    assert rel(yc.float(), ye.float().cpu()) < 2e-2


@pytest.mark.parametrize("dtype,DK,DV", [(torch.float32, 32, 64), (torch.bfloat16, 32, 64), (torch.bfloat16, 64, 128)])
def test_unequal_head_dims_match_oracle(dtype, DK, DV):
    """qk_dim_factor = 0.5 of the reference's mLSTMLayerVision (mlstm_large.py:46,186-187): q, k of head dim DHqk, v and h of
    head dim DHv = 2 DHqk.  bf16 runs the tcgen05 family on q, k zero-padded to DHv inside the library (mlstm_api.cu: every
    q / k dependent quantity is an inner product over DHqk), fp32 the SIMT family; the reference's own PyTorch chunkwise_simple
    cannot run this shape at all (it views q, k, v with one DH, backends.py:163-171), so the oracle restatement is the pin."""
    B, NH, S = 2, 2, 200
    g = torch.Generator().manual_seed(5)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    q, k = mk(B, S, NH, DK, sc=DK ** -0.5).to(dtype).transpose(1, 2), mk(B, S, NH, DK, sc=DK ** -0.5).to(dtype).transpose(1, 2)
    v, dh = mk(B, S, NH, DV).to(dtype).transpose(1, 2), mk(B, S, NH, DV).to(dtype).transpose(1, 2)
    i, f = mk(B, S, NH).transpose(1, 2), (torch.linspace(3, 6, NH).view(1, 1, NH) + mk(B, S, NH)).transpose(1, 2)
    inputs = [q, k, v, i, f, dh]
    ref = O.mlstm_fwbw(*(x.double() for x in inputs), chunk_size=64, eps=1e-6)
    got = run_cuda(inputs)
    from xlstm_yolo_b200 import ops
    assert ops.kernel_family(q.cuda(), v.cuda()) == ("simt" if dtype == torch.float32 else "tcgen05")
    check(got, ref, torch.float32 if dtype == torch.float32 else torch.bfloat16, f"DHqk {DK} DHv {DV}")


@pytest.mark.parametrize("DK,DV,S,reverse", [(64, 128, 300, False), (32, 64, 700, True), (128, 256, 200, False)])
def test_unequal_head_dims_with_states(DK, DV, S, reverse):
    """The padded tensor-core path (bf16, DHqk < DHv) with carried state: c_initial (B, NH, DHqk, DHv) / n_initial are padded
    with zero rows inside the library and the last states cropped back; both scan directions, single-pass and chunk-parallel
    lengths, DHv = 256 (the slice-streaming family)."""
    from xlstm_yolo_b200 import ops
    B, NH = 2, 2
    g = torch.Generator().manual_seed(9)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    bf = torch.bfloat16
    q, k = mk(B, S, NH, DK, sc=DK ** -0.5).to(bf).transpose(1, 2), mk(B, S, NH, DK, sc=DK ** -0.5).to(bf).transpose(1, 2)
    v, dh = mk(B, S, NH, DV).to(bf).transpose(1, 2), mk(B, S, NH, DV).to(bf).transpose(1, 2)
    i, f = mk(B, S, NH).transpose(1, 2), (torch.linspace(3, 6, NH).view(1, 1, NH) + mk(B, S, NH)).transpose(1, 2)
    inputs = [q, k, v, i, f, dh]
    st = dict(c_initial=mk(B, NH, DK, DV), n_initial=mk(B, NH, DK), m_initial=mk(B, NH, 1))
    assert ops.kernel_family(q.cuda(), v.cuda()) == "tcgen05"
    ref = oracle_on_kernel_side(inputs, reverse=reverse, states=st, what=f"states DHqk{DK} DHv{DV} rev={reverse}")
    got = run_cuda(inputs, reverse=reverse, states=st)
    check(got, ref, bf, f"states DHqk {DK} DHv {DV}")
    h, (C, n, m) = ops.mlstm(*(x.cuda() for x in inputs[:5]), *(s_.cuda() for s_ in st.values()), return_last_states=True,
                             reverse=reverse)
    _, (Cr, nr, mr) = O.mlstm_chunkwise(*(x.double() for x in inputs[:5]), *(s_.double() for s_ in st.values()),
                                        chunk_size=64, return_last_states=True, reverse=reverse)
    assert C.shape == (B, NH, DK, DV) and n.shape == (B, NH, DK)
    th = TOL[bf][0]
    assert rel(C, Cr) < th and rel(n, nr) < th and rel(m, mr) < 1e-4


def test_reference_mlstm_layer_vision_runs_on_the_shim():
    """The reference's own mLSTMLayerVision (mlstm_large.py:135-330: soft-capped gates, sigmoid output gate, DHqk = DHv / 2,
    carried state) imported unmodified from baseline/_ref, its mLSTMBackend being this repo's through the import shim: CUDA
    forward + backward against the same module on CPU tensors (the native-PyTorch branch of the backend)."""
    import os
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.isdir(os.path.join(root, "ultralytics")):
        pytest.skip("baseline/_ref not present (made by baseline/make_ref.py where the reference tree exists)")
    from xlstm_yolo_b200.compat import reference_loader as RL
    RL.import_reference(root)
    from ultralytics.nn.modules.vision_lstm import mlstm_large as ML
    torch.manual_seed(0)
    cfg = ML.mLSTMVisionBlockConfig(embedding_dim=128, num_heads=2, chunkwise_kernel="chunkwise--triton_xl_chunk",
                                    sequence_kernel="native_sequence__triton", step_kernel="triton", autocast_kernel_dtype="float32")
    layer = ML.mLSTMLayerVision(cfg, seqlens=[14, 14])
    with torch.no_grad():
        for n, p in layer.named_parameters():
            if "gate" in n and p.dim() > 1:
                p.normal_(0, 0.05)
    x = torch.randn(2, 196, 128)
    res = []
    import copy
    torch.backends.cudnn.allow_tf32 = False          # the CPU run is plain fp32: keep the conv / linears of the CUDA run there too
    torch.backends.cuda.matmul.allow_tf32 = False
    for dev in ("cpu", "cuda"):
        m = copy.deepcopy(layer).to(dev)
        xi = x.detach().clone().to(dev).requires_grad_(True)
        y = m(xi)
        y.square().mean().backward()
        res.append((y.detach().cpu(), xi.grad.cpu(), m.igate_preact.weight.grad.cpu()))
    for a, b in zip(res[1], res[0]):
        assert torch.isfinite(a).all() and rel(a, b) < 2e-4
