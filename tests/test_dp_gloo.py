"""CPU, world_size 2, gloo: the data-parallel host logic (2g-gcn_b200/dp.py) — parameter broadcast, one all-reduce of
the flat gradient buffer, averaging, rebinding of the parameter gradients, and batch sharding after padding.  The CUDA
model cannot run here, so a stand-in with the same ``flat_grad`` / ``bind_flat_grads`` contract is used."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _StandIn(torch.nn.Module):
    """Same gradient contract as TGGCN: backward fills one flat buffer, parameters get views of it."""

    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.a = torch.nn.Parameter(torch.randn(5, 3, generator=g))
        self.b = torch.nn.Parameter(torch.randn(7, generator=g))
        self.dead = torch.nn.Parameter(torch.randn(2, generator=g))      # off the gradient path: stays grad=None
        self.register_buffer('stat', torch.randn(4, generator=g))
        self.flat_grad = None

    def fake_backward(self, scale):
        params = [self.a, self.b]
        self.flat_grad = torch.cat([torch.full((p.numel(),), float(scale) * (i + 1)) for i, p in enumerate(params)])
        offs = [0, self.a.numel()]
        self._views = [self.flat_grad[o:o + p.numel()].view(p.shape) for o, p in zip(offs, params)]
        for p in params:
            p.grad = torch.zeros_like(p)          # what autograd may have left there (a copy)

    def bind_flat_grads(self):
        for p, g in zip([self.a, self.b], self._views):
            p.grad = g


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        dp = importlib.import_module('2g-gcn_b200.dp')
        model = _StandIn(seed=100 + rank)                # replicas start different ...
        red = dp.GradientAllReduce(model)
        red.sync_parameters()                            # ... and are made identical to rank 0
        ref = _StandIn(seed=100)
        same = all(torch.equal(p, q) for p, q in zip(list(model.parameters()) + list(model.buffers()),
                                                     list(ref.parameters()) + list(ref.buffers())))
        model.fake_backward(scale=rank + 1)              # rank 0: 1,2 ; rank 1: 2,4
        flat = red.reduce()
        ok_avg = torch.allclose(model.a.grad, torch.full((5, 3), 1.5)) and torch.allclose(model.b.grad, torch.full((7,), 3.0))
        bound = model.a.grad.data_ptr() == flat.data_ptr() and model.dead.grad is None
        batch = {'x': torch.arange(8 * 3).view(8, 3), 'mask': torch.ones(8, 2), 'n': 5}
        sh = dp.shard_batch(batch, rank, world)
        shard_ok = sh['x'].shape == (4, 3) and int(sh['x'][0, 0]) == rank * 12 and sh['n'] == 5
        try:
            dp.shard_batch({'x': torch.zeros(7, 2)}, rank, world)
            uneven = False
        except ValueError:
            uneven = True
        out[rank] = (same, ok_avg, bound, shard_ok, uneven)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        assert out[r] == (True, True, True, True, True), (r, out[r])
