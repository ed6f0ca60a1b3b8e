"""Bisect the slow e2e compute loop seen inside bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from xlstm_yolo_b200 import ops
from xlstm_yolo_b200.backend import mLSTMBackend, mLSTMBackendConfig

B, NH, S, DH = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
be = mLSTMBackend(mLSTMBackendConfig(chunk_size=64, eps=1e-6, autocast_kernel_dtype="bfloat16"))
host = [t.pin_memory() for t in bench.make_inputs(torch, B, NH, S, DH, 77, "cpu", torch.bfloat16)]
dev_in = [torch.empty_like(t, device=dev) for t in host]
for d_, h_ in zip(dev_in, host):
    d_.copy_(h_)
res_host = torch.empty(4, dtype=torch.float32).pin_memory()

def compute(reduce):
    q, k, v, i, f, dh = bench.as_heads(dev_in)
    leaves = [t.detach().requires_grad_(True) for t in (q, k, v, i, f)]
    h = be(*leaves)
    h.backward(dh)
    if reduce:
        res = torch.stack([h.float().abs().mean(), leaves[0].grad.float().abs().mean(),
                           leaves[3].grad.abs().mean(), leaves[4].grad.abs().mean()])
        res_host.copy_(res, non_blocking=True)

def timed(tag, reduce, n=20):
    for _ in range(3):
        compute(reduce)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        compute(reduce)
    th = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    tw = (time.perf_counter() - t0) / n * 1e3
    print(f"{tag}: host {th:.3f} ms, wall {tw:.3f} ms", flush=True)

timed("baseline", False)
timed("with reductions", True)
plans = []
for s_ in range(4):
    q, k, v, i, f, dh = bench.as_heads(bench.make_inputs(torch, B, NH, S, DH, s_, dev, torch.bfloat16))
    plans.append(ops.MLSTMPlan(q, k, v, i, f, dh, eps=1e-6, chunk_size=64))
for it in range(40):
    pl = plans[it % 4]; pl.forward(); pl.backward(0); pl.backward(1)
torch.cuda.synchronize()
timed("after plans", True)
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(200)]
for e4 in evs:
    for e in e4:
        e.record()
torch.cuda.synchronize()
timed("after 800 timing events", True)
samp = bench.ClockSampler(0)
samp.start(); time.sleep(0.3); samp.stop_flag = True; samp.join(timeout=1.0)
print("sampler alive:", samp.is_alive(), samp.result())
timed("after nvml sampler", True)
