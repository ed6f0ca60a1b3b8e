"""The whole detector's tensor path (the reference's unmodified DetectionModel on the authored s-scale YAML with this repo's
ViLBlockPair drop-in) replayed from CUDA graphs gives what the eager program gives: every op of this repo on that path —
producer, gate projection, cell, tail, their backward kernels — is a plain launch on the current stream with static shapes
and no host synchronisation, so forward and backward capture (bench_detector.py --graph 1 relies on it).

Needs the reference tree under baseline/_ref (placed by baseline/make_ref.py; it travels to the GPU box)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ultralytics")), reason="baseline/_ref/ultralytics not present")
def test_detector_forward_backward_from_cuda_graphs_matches_eager(tmp_path, monkeypatch):
    import yaml
    monkeypatch.setenv("YOLO_CONFIG_DIR", str(tmp_path))
    from xlstm_yolo_b200.compat import reference_loader as RL
    RL.import_reference(REF)
    RL.use_b200_dropins(pair_level=True, head_compat=False)
    from ultralytics.nn.tasks import DetectionModel

    cfg = yaml.safe_load(open(os.path.join(ROOT, "xlstm_yolo_b200", "compat", "yamls", "xlstm-yolo-s.yaml")))
    cfg["scale"] = "s"
    torch.manual_seed(0)
    model = DetectionModel(cfg, ch=3, nc=80, verbose=False).cuda().train()
    for m in model.modules():                      # frozen batch statistics: eager and replayed runs see the same network
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()

    class TensorPath(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
                return tuple(self.m.predict(x))

    path = TensorPath(model)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(2, 3, 320, 320, device="cuda", generator=g)
    probes = [p for n, p in model.named_parameters() if any(k in n for k in ("q_proj.weight", "igate.weight", "conv.weight", "proj_up.weight"))]
    assert len(probes) >= 8

    def run(fn):
        for p in model.parameters():
            p.grad = None
        outs = fn(x)
        sum((o.float() * w).sum() for o, w in zip(outs, (1.0, 0.5, 0.25))).backward()
        torch.cuda.synchronize()
        # (the reference builds modules its forward never calls: their parameters get no gradient in either run)
        return ([o.detach().float().clone() for o in outs],
                [torch.zeros(1, device="cuda") if p.grad is None else p.grad.detach().float().clone() for p in probes])

    eager_out, eager_grad = run(path)
    graphed = torch.cuda.make_graphed_callables(path, (torch.rand_like(x),), allow_unused_input=True)
    for _ in range(2):                              # replayed twice: static buffers are reused correctly
        got_out, got_grad = run(graphed)
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    for a, b in zip(got_out, eager_out):
        assert torch.isfinite(a).all() and rel(a, b) < 1e-3
    assert sum(float(b.abs().sum()) > 0 for b in eager_grad) >= 8
    for a, b in zip(got_grad, eager_grad):
        assert a.shape == b.shape and torch.isfinite(a).all() and rel(a, b) < 1e-3
