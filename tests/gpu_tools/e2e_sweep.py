"""Developer probe: the bench's end-to-end arm at several step counts, with the caching allocator's device-malloc count."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_WORKLOAD
for full in (True, False):
    for K in (3, 10, 50, 50):
        s0 = torch.cuda.memory_stats(dev)
        r = bench.measure_e2e(torch, None, name, K, dev, 0, 1, full_result=full)
        s1 = torch.cuda.memory_stats(dev)
        print(f"{name} full={full} K={K}: {r['ms_per_step']:.3f} ms/step, {r['value'] / 1e6:.1f} M tok/s, "
              f"device mallocs {s1['num_device_alloc'] - s0['num_device_alloc']}, reserved {s1['reserved_bytes.all.current'] / 1e6:.0f} MB",
              flush=True)
