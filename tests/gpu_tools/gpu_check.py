"""Quick GPU parity sweep (developer tool): CUDA op vs the CPU oracle on seeded inputs."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mlstm_oracle as O  # noqa: E402
from xlstm_yolo_b200 import ops  # noqa: E402


def rel(a, b):
    return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def make(B, NH, S, DH, dtype, regime, qk_std=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(B, S, NH, DH, generator=g) * qk_std).to(dtype)
    k = (torch.randn(B, S, NH, DH, generator=g) * qk_std).to(dtype)
    v = torch.randn(B, S, NH, DH, generator=g).to(dtype)
    dh = torch.randn(B, S, NH, DH, generator=g).to(dtype)
    if regime == "rand":
        i = torch.randn(B, S, NH, generator=g)
    elif regime == "refinit":
        i = -10 + 0.1 * torch.randn(B, S, NH, generator=g)
    else:
        i = 2 * torch.randn(B, S, NH, generator=g)
    f = torch.linspace(3, 6, NH).view(1, 1, NH) + torch.randn(B, S, NH, generator=g)
    if regime == "forget":
        f = f - 4
    t = lambda x: x.transpose(1, 2)
    return t(q), t(k), t(v), t(i), t(f), t(dh)


def run(B, NH, S, DH, dtype, regime, reverse=False, eps=1e-6, qk_std=1.0, states=False):
    q, k, v, i, f, dh = make(B, NH, S, DH, dtype, regime, qk_std)
    d = [x.double() for x in (q, k, v, i, f, dh)]
    kw = {}
    if states:
        g = torch.Generator().manual_seed(7)
        kw = dict(c_initial=torch.randn(B, NH, DH, DH, generator=g).double(), n_initial=torch.randn(B, NH, DH, generator=g).double(),
                  m_initial=torch.randn(B, NH, 1, generator=g).double())
    ref = O.mlstm_fwbw(*d, chunk_size=64, eps=eps, reverse=reverse, **kw)
    cq, ck, cv, ci, cf, cdh = (x.cuda() for x in (q, k, v, i, f, dh))
    leaves = [x.detach().requires_grad_(True) for x in (cq, ck, cv, ci, cf)]
    kwc = {kk: vv.float().cuda() for kk, vv in kw.items()}
    torch.cuda.synchronize()
    t0 = time.time()
    h = ops.mlstm(*leaves, eps=eps, reverse=reverse, **kwc)
    h.backward(cdh)
    torch.cuda.synchronize()
    dt = time.time() - t0
    errs = [rel(h, ref[0])] + [rel(l.grad, r) for l, r in zip(leaves, ref[1:])]
    fam = ops.kernel_family(leaves[0].to(torch.bfloat16 if dtype != torch.float32 else torch.float32), leaves[2])
    print(f"B{B} NH{NH} S{S} DH{DH} {str(dtype)[6:]:8s} {regime:8s} rev={int(reverse)} st={int(states)} [{fam}] "
          + " ".join(f"{n}:{e:.1e}" for n, e in zip(["h", "dq", "dk", "dv", "di", "df"], errs)) + f"  ({dt*1e3:.1f} ms)", flush=True)
    return errs


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "case":  # case B NH S DH dtype regime reverse states qk_std
        a = sys.argv[2:]
        run(int(a[0]), int(a[1]), int(a[2]), int(a[3]), getattr(torch, a[4]), a[5], reverse=bool(int(a[6])),
            states=bool(int(a[7])), qk_std=float(a[8]))
        sys.exit(0)
    if which in ("all", "simt"):
        for regime in ["rand", "refinit", "forget"]:
            run(2, 2, 100, 16, torch.float32, regime)
        run(2, 4, 400, 64, torch.float32, "rand", reverse=True)
        run(1, 2, 70, 32, torch.float32, "rand", states=True)
        run(2, 2, 300, 128, torch.float32, "refinit", eps=5e-5)
        run(2, 8, 256, 16, torch.bfloat16, "rand", qk_std=0.25)
    if which in ("all", "tc"):
        run(1, 1, 128, 64, torch.bfloat16, "rand", qk_std=0.125)
        run(1, 1, 128, 128, torch.bfloat16, "rand", qk_std=0.09)
        run(2, 4, 400, 64, torch.bfloat16, "rand", qk_std=0.125)
        run(2, 4, 400, 64, torch.bfloat16, "refinit", qk_std=0.125, eps=5e-5, reverse=True)
        run(2, 4, 400, 128, torch.bfloat16, "forget", qk_std=0.09, reverse=True)
        run(1, 4, 1600, 128, torch.bfloat16, "rand", qk_std=0.09)
        run(1, 2, 300, 128, torch.bfloat16, "rand", qk_std=0.09, states=True)
        run(1, 2, 300, 64, torch.bfloat16, "rand", qk_std=0.125, states=True, reverse=True)
        run(2, 2, 1, 64, torch.bfloat16, "rand", qk_std=0.125)
