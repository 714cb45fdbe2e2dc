"""CPU, world_size 2, gloo: the data-parallel host logic — parameter broadcast, bucketed all-reduce of the flat gradient buffer,
valid-count loss weights (the DP(2) update must equal the single-process update on the global batch even when the ranks hold
different numbers of valid target elements), the sharded sampler and the epoch driver (2g-gcn_b200/dp.py, trainer.py).  The CUDA
model cannot run here, so a small torch model with the same ``flat_grad`` / ``grad_buckets`` contract stands in."""
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
    """Same gradient contract as TGGCN: after backward, every trainable parameter's .grad is a view of ONE flat buffer."""

    def __init__(self, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.a = torch.nn.Parameter(torch.randn(5, 3, generator=g))
        self.b = torch.nn.Parameter(torch.randn(7, generator=g))
        self.dead = torch.nn.Parameter(torch.randn(2, generator=g))      # off the gradient path: stays grad=None
        self.register_buffer('stat', torch.randn(4, generator=g))
        self.flat_grad, self.grad_buckets, self.grad_ready_callback = None, [], None

    def fake_backward(self, scale):
        params = [self.a, self.b]
        self.flat_grad = torch.cat([torch.full((p.numel(),), float(scale) * (i + 1)) for i, p in enumerate(params)])
        offs = [0, self.a.numel()]
        for o, p in zip(offs, params):
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self.grad_buckets = [(0, self.a.numel(), None), (self.a.numel(), self.flat_grad.numel(), None)]
        if self.grad_ready_callback is not None:
            self.grad_ready_callback(self)


class _Regressor(torch.nn.Module):
    """Per-video model for the trainer test: output list like the real model's (one tensor per loss term)."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(4, 3)
        self.flat_grad, self.grad_buckets, self.grad_ready_callback = None, [], None
        self.lin.weight.register_hook(lambda g: None)

    def forward(self, x):
        y = self.lin(x)                        # (B, T, 3)
        return [y[..., 0], y[..., 1:]]

    def collect(self):
        """What TGGCN._backward does: gradients into one flat buffer, .grad = views, then the reducer's callback."""
        ps = [self.lin.weight, self.lin.bias]
        self.flat_grad = torch.cat([p.grad.reshape(-1) for p in ps])
        off = 0
        for p in ps:
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.grad_buckets = [(0, off, None)]
        if self.grad_ready_callback is not None:
            self.grad_ready_callback(self)


def _masked_mse(outputs, targets, reduction='mean'):
    """Two terms, each a mean over the valid (target != -1) elements of the LOCAL batch like pyrutils/torch/losses.py."""
    out = []
    for o, t in zip(outputs, targets):
        m = t != -1.0
        out.append(((o - t) ** 2 * m).sum() / m.sum().clamp(min=1))
    return out


def _dataset(n=8, T=6):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, T, 4, generator=g)
    t0 = torch.randn(n, T, generator=g)
    t1 = torch.randn(n, T, 2, generator=g)
    lengths = torch.tensor([6, 2, 5, 3, 6, 1, 4, 6])[:n]
    for i, L in enumerate(lengths):            # unequal padding: ranks see different numbers of valid elements
        t0[i, L:] = -1.0
        t1[i, L:] = -1.0
    return x, t0, t1


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        pkg = importlib.import_module('2g-gcn_b200')
        dp, trainer = pkg.dp, pkg.trainer
        # --- broadcast + bucketed all-reduce + rebinding --------------------------------------------------------------------
        model = _StandIn(seed=100 + rank)                # replicas start different ...
        red = dp.GradientAllReduce(model).attach()
        red.sync_parameters()                            # ... and are made identical to rank 0
        ref = _StandIn(seed=100)
        same = all(torch.equal(p, q) for p, q in zip(list(model.parameters()) + list(model.buffers()),
                                                     list(ref.parameters()) + list(ref.buffers())))
        model.fake_backward(scale=rank + 1)              # rank 0: 1,2 ; rank 1: 2,4
        flat = red.reduce()
        ok_avg = torch.allclose(model.a.grad, torch.full((5, 3), 1.5)) and torch.allclose(model.b.grad, torch.full((7,), 3.0))
        bound = model.a.grad.data_ptr() == flat.data_ptr() and model.dead.grad is None
        # --- batch sharding -------------------------------------------------------------------------------------------------
        batch = {'x': torch.arange(8 * 3).view(8, 3), 'mask': torch.ones(8, 2), 'n': 5}
        sh = dp.shard_batch(batch, rank, world)
        shard_ok = sh['x'].shape == (4, 3) and int(sh['x'][0, 0]) == rank * 12 and sh['n'] == 5
        try:
            dp.shard_batch({'x': torch.zeros(7, 2)}, rank, world)
            uneven = False
        except ValueError:
            uneven = True
        # --- sampler: ranks partition every global batch, all ranks run the same number of steps ---------------------------
        s = trainer.ShardedBatchSampler(10, 4, rank, world, shuffle=True, seed=3)
        s.set_epoch(2)
        mine = list(s)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        flat_idx = [i for step in range(len(s)) for r in range(world) for i in gathered[r][step]]
        sampler_ok = (len(s) == 3 and all(len(b) == 2 for b in mine) and sorted(set(flat_idx)) == list(range(10))
                      and len(flat_idx) == 12)
        # --- DP(2) step == single-process step on the global batch, with unequal valid counts ------------------------------
        x, t0, t1 = _dataset()
        torch.manual_seed(0)
        single = _Regressor()
        losses = _masked_mse(single(x), [t0, t1])
        sum(losses).backward()
        want = torch.cat([single.lin.weight.grad.reshape(-1), single.lin.bias.grad.reshape(-1)])
        torch.manual_seed(0)
        m2 = _Regressor()
        red2 = dp.GradientAllReduce(m2).attach()
        sl = slice(rank * 4, rank * 4 + 4)
        w = dp.loss_term_weights([t0[sl], t1[sl]])
        local = _masked_mse(m2(x[sl]), [t0[sl], t1[sl]])
        sum(l * wi for l, wi in zip(local, w.unbind(0))).backward()
        m2.collect()
        got = red2.reduce()
        weighted_ok = torch.allclose(got, want, rtol=1e-5, atol=1e-7)
        plain = torch.cat([g.reshape(-1) for g in torch.autograd.grad(sum(_masked_mse(m2(x[sl]), [t0[sl], t1[sl]])),
                                                                      [m2.lin.weight, m2.lin.bias])])
        dist.all_reduce(plain)
        plain_differs = not torch.allclose(plain / world, want, rtol=1e-3, atol=1e-6)      # why the weighting exists
        # --- the epoch driver on the stand-in (CPU tensors, gloo) ----------------------------------------------------------
        torch.manual_seed(1)
        m3 = _Regressor()
        opt = torch.optim.SGD(m3.parameters(), lr=0.05)

        class Crit:
            def __call__(self, output, target, reduction='mean'):
                return _masked_mse(output, target)

        orig_backward = torch.Tensor.backward

        def fetch(ds, device):
            return [ds[0]], [ds[1], ds[2]]

        def feed(m, data):
            return m(data[0])

        tr = trainer.DataParallelTrainer(m3, opt, Crit(), ['a', 'b'], 'cpu', fetch, feed, verbose=False, log_interval=2)
        # the stand-in has no custom autograd Function: collect the flat buffer right after backward like TGGCN._backward does
        red3 = tr.reducer
        real_reduce = red3.reduce
        red3.reduce = lambda: (m3.collect(), real_reduce())[1]
        ck = tr.fit(torch.utils.data.TensorDataset(x, t0, t1), epochs=4, global_batch=4, val_dataset=(x, t0, t1), seed=1)
        params_equal = [p.detach().clone() for p in m3.parameters()]
        gathered_p = [None] * world
        dist.all_gather_object(gathered_p, params_equal)
        replicas_ok = all(torch.equal(a, b) for a, b in zip(gathered_p[0], gathered_p[1]))
        ck_ok = (set(ck) >= {'epoch', 'model_state_dict', 'train_losses', 'val_losses', 'train_raw_losses', 'val_raw_losses'}
                 and len(ck['train_losses']) == 4 and ck['val_losses'][-1][0] < ck['val_losses'][0][0])
        out[rank] = (same, ok_avg, bound, shard_ok, uneven, sampler_ok, weighted_ok, plain_differs, replicas_ok, ck_ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_host_logic_two_ranks():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    for r in range(2):
        assert out[r] == (True,) * 10, (r, out[r])


def test_sampler_single_rank_covers_every_video_once_per_epoch():
    sys.path.insert(0, ROOT)
    trainer = importlib.import_module('2g-gcn_b200.trainer')
    s = trainer.ShardedBatchSampler(9, 4, shuffle=True, seed=0)
    s.set_epoch(1)
    a = [i for b in s for i in b]
    s.set_epoch(2)
    b = [i for bb in s for i in bb]
    assert len(a) == 12 and sorted(set(a)) == list(range(9)) and a != b          # 3 steps; the tail wraps; epochs differ
    with pytest.raises(ValueError):
        trainer.ShardedBatchSampler(9, 5, 0, 2)
