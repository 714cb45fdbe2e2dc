"""GPU: the drop-in model inside the data-parallel training driver (2g-gcn_b200/trainer.py), through the reference's contract
(pyrutils/torch/train_utils.py:12-165): dataset tuples -> fetcher -> feeder -> criterion list -> backward -> clip -> Adam.  The loss
must go down when the same small dataset is revisited, the checkpoint dictionary must have the reference's keys, and — on a box
with two GPUs — the DP(2) gradient must equal the single-process gradient on the same global batch (SURVEY.md §4 item vi)."""
import importlib
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fetch(ds, device):           # stand-in for gcn_fetcher (vhoi/data_loading.py:1282-1315)
    ds = [t.to(device) for t in ds]
    return ds[:8], ds[8:]


def _feed(m, data):               # stand-in for gcn_forward (vhoi/data_loading.py:1233-1279), stage-2 settings
    return m(x_human=data[0], x_objects=data[1], objects_mask=data[2], human_segmentation=None,
             steps_per_example=data[7], inspect_model=False)


def _dataset(synth, shape, n_videos, T, seed=21):
    batch = synth.make_batch(shape, n_videos, T, seed=seed)
    targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=seed + 1))
    zeros = torch.zeros(n_videos, 1)
    # tuple layout of the reference's TensorDataset (vhoi/data_loading.py:517-519): 8 inputs then the targets
    tensors = [batch['x_human'], batch['x_objects'], batch['objects_mask'], zeros, zeros, zeros, zeros, batch['steps_per_example']]
    return torch.utils.data.TensorDataset(*tensors, *targets)


def test_trainer_reduces_the_loss_and_returns_the_reference_checkpoint(orc, synth, pkg, tmp_path):
    shape = synth.SHAPES['mphoi']
    torch.manual_seed(3)
    model = pkg.TGGCN(**synth.model_kwargs(shape, hidden_size=32, stage=2)).cuda()
    dataset = _dataset(synth, shape, 8, 12)

    def criterion(output, target, reduction='mean'):
        return orc.multi_task_loss(output, target, 'mphoi', 2)

    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    tr = pkg.trainer.DataParallelTrainer(model, opt, criterion, ['hb', 'hs', 'fr', 'fp', 'sr', 'sp'], 'cuda', _fetch, _feed,
                                         clip_gradient_at=5.0, verbose=False, log_interval=1)
    torch.manual_seed(5)                               # the model draws its Gumbel noise from the global CPU generator
    path = str(tmp_path / 'ckpt.tar')
    ck = tr.fit(dataset, epochs=6, global_batch=4, val_dataset=dataset, seed=2, checkpoint_path=path)
    totals = [t for t, _ in ck['train_losses']]
    assert all(torch.isfinite(torch.tensor(totals)))
    assert totals[-1] < 0.92 * totals[0], totals
    assert set(ck) >= {'epoch', 'model_state_dict', 'train_losses', 'val_losses', 'train_raw_losses', 'val_raw_losses'}
    assert 1 <= ck['epoch'] <= 6 and len(ck['val_losses']) == 6
    # the file is a reference-format checkpoint: a fresh model loads it strictly (train.py:37 / predict.py:43 use strict=False)
    loaded = torch.load(path, weights_only=False)
    fresh = pkg.TGGCN(**synth.model_kwargs(shape, hidden_size=32, stage=2))
    fresh.load_state_dict(loaded['model_state_dict'], strict=True)


def test_device_resident_dataset_and_pipeline(synth, pkg):
    shape = synth.SHAPES['cad120']
    ds = _dataset(synth, shape, 6, 5)
    res = pkg.feeder.DeviceResidentDataset.from_tensor_dataset(ds, 'cuda')
    got = res.batch([4, 1])
    for g, t in zip(got, ds.tensors):
        assert g.is_cuda and torch.equal(g.cpu(), t[[4, 1]])
    assert len(res) == 6 and res.staged_bytes == sum(t.numel() * t.element_size() for t in ds.tensors)
    # double-buffered pipeline: slot discipline (a batch taken by get() but not released must keep its slot)
    host = {'a': torch.arange(6, dtype=torch.float32).pin_memory(), 'b': torch.ones(3).pin_memory()}
    pipe = pkg.feeder.DeviceBatchPipeline('cuda', host, depth=2)
    pipe.submit(host)
    pipe.submit({'a': host['a'] * 2, 'b': host['b']})
    first = pipe.get()
    with pytest.raises(RuntimeError, match='released'):
        pipe.submit(host)                           # both slots hold unreleased batches
    with pytest.raises(RuntimeError, match='release'):
        pipe.get()                                  # one batch at a time
    assert torch.equal(first['a'].cpu(), host['a'])
    pipe.release()
    pipe.submit({'a': host['a'] * 3, 'b': host['b']})
    second = pipe.get()
    torch.cuda.synchronize()
    assert torch.equal(second['a'].cpu(), host['a'] * 2)
    pipe.release()
    with pytest.raises(RuntimeError, match='matching'):
        pipe.release()


def _dp_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import torch.distributed as dist
    import tggcn_oracle as orc
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        pkg = importlib.import_module('2g-gcn_b200')
        synth = pkg.synth
        shape = synth.SHAPES['mphoi']
        B, T, D = 4, 10, 64
        kw = synth.model_kwargs(shape, hidden_size=D, stage=2)
        model = pkg.TGGCN(**kw)
        synth.deterministic_fill(model.state_dict(), seed=11, gain=2.0)
        model = model.cuda().eval()           # eval-mode BatchNorm: running statistics, so replicas and the single process agree
        for p in model.parameters():
            p.requires_grad_(True)
        batch = synth.make_batch(shape, B, T, seed=31)          # unequal lengths -> unequal valid counts per rank
        targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=32))
        noise = orc.draw_noise(T * (shape.H + shape.O), B, torch.Generator().manual_seed(33))

        def run(sl, weights):
            model.set_gumbel_noise(noise[:, sl])
            model.zero_grad(set_to_none=True)
            outp = model(x_human=batch['x_human'][sl].cuda(), x_objects=batch['x_objects'][sl].cuda(),
                         objects_mask=batch['objects_mask'][sl].cuda())
            losses = orc.multi_task_loss(outp, [t[sl].cuda() for t in targets], 'mphoi', 2)
            sum(l * w for l, w in zip(losses, weights)).backward()
            return model.flat_grad

        full = run(slice(0, B), [1.0] * 6).clone()              # the single-process gradient of the global batch
        red = pkg.dp.GradientAllReduce(model).attach()
        per = B // world
        sl = slice(rank * per, (rank + 1) * per)
        w = pkg.dp.loss_term_weights([t[sl].cuda() for t in targets])
        run(sl, list(w.unbind(0)))
        got = red.reduce()
        torch.cuda.synchronize()
        model.check_persistent_kernels()
        scale = float(full.abs().max())
        err = float((got - full).abs().max())
        out[rank] = (err, scale, len(model.grad_buckets))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (gpurun --gpus 2)')
def test_dp2_gradient_equals_single_process_gradient():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    for r in range(2):
        err, scale, n_buckets = out[r]
        assert n_buckets >= 3
        assert err <= 2e-3 * scale + 1e-7, (r, err, scale)       # same tolerance as the backward parity tests
