"""GPU: FlatAdam (one-launch Adam over the flat parameter / gradient buffers) against torch.optim.Adam on the same gradients."""
import copy
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(weight_decay):
    pkg = importlib.import_module('2g-gcn_b200')
    shape = pkg.synth.SHAPES['mphoi']
    kw = pkg.synth.model_kwargs(shape, hidden_size=32, stage=2)
    torch.manual_seed(3)
    a = pkg.TGGCN(**kw).cuda().train()
    b = copy.deepcopy(a)
    batch = pkg.synth.make_batch(shape, 2, 8, seed=5)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    noise = pkg.TGGCN.draw_gumbel_noise(8 * (shape.H + shape.O), 2).cuda()
    a.set_gumbel_noise(noise), b.set_gumbel_noise(noise)
    return pkg, a, b, x, weight_decay


@pytest.mark.parametrize('weight_decay', [0.0, 0.01])
def test_flat_adam_matches_torch_adam(weight_decay):
    pkg, a, b, x, wd = _setup(weight_decay)
    ref = torch.optim.Adam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    ours = pkg.optim.FlatAdam(b, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    names = [n for n, _ in a.named_parameters()]
    for it in range(6):
        for model, opt in ((a, ref), (b, ours)):
            opt.zero_grad(set_to_none=True)
            out = model(**x)
            loss = sum(o.float().pow(2).mean() for o in out[2:]) + out[1].mean()
            loss.backward()
            opt.step()
        for n, pa, pb in zip(names, a.parameters(), b.parameters()):
            # entries whose gradient is of the order of eps move by a fraction of lr that depends on the last bits of the gradient (the
            # backward's split-K atomics do not sum in a fixed order): absolute tolerance = a few 1e-3 of lr per step
            torch.testing.assert_close(pb, pa, rtol=2e-5, atol=6e-6 * (it + 1), msg=lambda m: f'step {it}, {n}: {m}')
    # parameters keep their identity and shapes; their storage is one buffer now
    assert [n for n, _ in b.named_parameters()] == names
    trained = [p for p in b.parameters() if p.grad is not None]
    base = ours._p.data_ptr()
    assert all(base <= p.data_ptr() < base + 4 * ours._p.numel() for p in trained)
    # state_dict interchanges with torch.optim.Adam's
    sd = ours.state_dict()
    fresh = torch.optim.Adam(a.parameters(), lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    fresh.load_state_dict(sd)
    again = pkg.optim.FlatAdam(b, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    again.load_state_dict(ref.state_dict())
    for model, opt in ((a, fresh), (b, again)):
        opt.zero_grad(set_to_none=True)
        out = model(**x)
        (sum(o.float().pow(2).mean() for o in out[2:]) + out[1].mean()).backward()
        opt.step()
    for n, pa, pb in zip(names, a.parameters(), b.parameters()):
        torch.testing.assert_close(pb, pa, rtol=2e-5, atol=5e-5, msg=lambda m: f'after state_dict exchange, {n}: {m}')
