"""Where does the e2e step time go?  H2D alone, host enqueue alone, compute alone."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from xlstm_yolo_b200.backend import mLSTMBackend, mLSTMBackendConfig

B, NH, S, DH = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
be = mLSTMBackend(mLSTMBackendConfig(chunk_size=64, eps=1e-6, autocast_kernel_dtype="bfloat16"))
host = [t.pin_memory() for t in bench.make_inputs(torch, B, NH, S, DH, 77, "cpu", torch.bfloat16)]
dev_in = [torch.empty_like(t, device=dev) for t in host]
nbytes = sum(t.numel() * t.element_size() for t in host)

def ev():
    return torch.cuda.Event(enable_timing=True)

def h2d():
    for d_, h_ in zip(dev_in, host):
        d_.copy_(h_, non_blocking=True)

def compute():
    q, k, v, i, f, dh = bench.as_heads(dev_in)
    leaves = [t.detach().requires_grad_(True) for t in (q, k, v, i, f)]
    h = be(*leaves)
    h.backward(dh)
    return h, leaves

for rep in range(4):
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(20):
        h2d()
    b.record(); torch.cuda.synchronize()
    t_h2d = a.elapsed_time(b) / 20
    a, b = ev(), ev()
    t0 = time.perf_counter()
    a.record()
    for _ in range(20):
        compute()
    t_host = (time.perf_counter() - t0) / 20 * 1e3
    b.record(); torch.cuda.synchronize()
    t_wall = (time.perf_counter() - t0) / 20 * 1e3
    t_dev = a.elapsed_time(b) / 20
    print(f"rep {rep}: h2d {t_h2d:.3f} ms ({nbytes / t_h2d / 1e6:.1f} GB/s)  compute: host-enqueue {t_host:.3f} ms, device {t_dev:.3f} ms, wall {t_wall:.3f} ms", flush=True)
print("cpus", os.cpu_count(), "load", os.getloadavg())
