import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from xlstm_yolo_b200.backend import mLSTMBackend, mLSTMBackendConfig
print("ALLOC_CONF", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"), os.environ.get("PYTORCH_ALLOC_CONF"))
B, NH, S, DH = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
dev = torch.device("cuda", 0)
be = mLSTMBackend(mLSTMBackendConfig(chunk_size=64, eps=1e-6, autocast_kernel_dtype="bfloat16"))
dev_in = [t for t in bench.make_inputs(torch, B, NH, S, DH, 77, dev, torch.bfloat16)]
res_host = torch.empty(4, dtype=torch.float32).pin_memory()
res_dev = torch.zeros(4, device=dev)

def compute(mode):
    q, k, v, i, f, dh = bench.as_heads(dev_in)
    leaves = [t.detach().requires_grad_(True) for t in (q, k, v, i, f)]
    h = be(*leaves)
    h.backward(dh)
    if mode == "float":
        x = h.float()
    elif mode == "mean_bf16":
        x = h.abs().mean()
    elif mode == "reduce":
        res = torch.stack([h.float().abs().mean(), leaves[0].grad.float().abs().mean(),
                           leaves[3].grad.abs().mean(), leaves[4].grad.abs().mean()])
    elif mode == "reduce_d2h":
        res = torch.stack([h.float().abs().mean(), leaves[0].grad.float().abs().mean(),
                           leaves[3].grad.abs().mean(), leaves[4].grad.abs().mean()])
        res_host.copy_(res, non_blocking=True)
    elif mode == "sync":
        torch.cuda.synchronize()
    elif mode == "d2h":
        res_host.copy_(res_dev, non_blocking=True)
    elif mode == "grad_only":
        x = leaves[0].grad.float().abs().mean()
    elif mode == "gates_only":
        x = leaves[3].grad.abs().mean()

def timed(tag, n=20):
    for _ in range(3):
        compute(tag)
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    for _ in range(n):
        compute(tag)
    th = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    tw = (time.perf_counter() - t0) / n * 1e3
    st1 = torch.cuda.memory_stats()
    print(f"{tag}: host {th:.3f} ms; wall {tw:.3f} ms; cudaMalloc calls {st1['num_device_alloc'] - st0['num_device_alloc']}, frees {st1['num_device_free'] - st0['num_device_free']}", flush=True)

for m in ["none", "float", "mean_bf16", "grad_only", "gates_only", "d2h", "reduce", "reduce_d2h", "sync", "none"]:
    timed(m)
