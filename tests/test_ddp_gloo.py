"""world_size-2 gloo (CPU) test of the data-parallel path: the cell shards by batch with no
data-path collective; only its parameter gradients ride in DDP's all-reduce (engine/trainer.py:274)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xlstm_yolo_b200 import MatrixLSTMCell


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(seed=0):
    torch.manual_seed(seed)
    cell = MatrixLSTMCell(dim=32, num_heads=2, chunk_size=8)
    with torch.no_grad():
        cell.igate.weight.normal_(0, 0.05)
        cell.fgate.weight.normal_(0, 0.05)
        cell.igate.bias.fill_(0.0)
    return cell


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return [torch.randn(3, 20, 32, generator=g) for _ in range(3)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    ddp = torch.nn.parallel.DistributedDataParallel(_make())
    q, k, v = _data(rank)
    ddp(q, k, v).square().mean().backward()
    grads = {n: p.grad.clone() for n, p in ddp.module.named_parameters() if p.grad is not None}
    if rank == 0:
        torch.save(grads, out)
    # every rank ends with the same (averaged) gradient
    t = grads["igate.weight"].clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert torch.allclose(t, grads["igate.weight"])
    dist.destroy_process_group()


def test_ddp_two_ranks_average_gradients(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    want = None
    for rank in range(2):
        cell = _make()
        q, k, v = _data(rank)
        cell(q, k, v).square().mean().backward()
        g = {n: p.grad for n, p in cell.named_parameters() if p.grad is not None}
        want = g if want is None else {n: want[n] + g[n] for n in g}
    for n in want:
        assert torch.allclose(got[n], want[n] / 2, atol=1e-6), n
