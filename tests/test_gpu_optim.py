"""GPU: FlatAdam (one-launch Adam over the flat parameter / gradient buffers) against torch.optim.Adam fed the SAME gradients
(the backward's split-K atomics do not sum in a fixed order, and Adam's first steps amplify last-bit differences of gradients of
the order of eps into fractions of lr — so the reference optimiser gets copies of the gradients our model produced)."""
import copy
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup():
    pkg = importlib.import_module('2g-gcn_b200')
    shape = pkg.synth.SHAPES['mphoi']
    kw = pkg.synth.model_kwargs(shape, hidden_size=32, stage=2)
    torch.manual_seed(3)
    ours_model = pkg.TGGCN(**kw).cuda().train()
    ref_params = copy.deepcopy(ours_model)                      # only a parameter holder for torch.optim.Adam
    batch = pkg.synth.make_batch(shape, 2, 8, seed=5)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    ours_model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(8 * (shape.H + shape.O), 2).cuda())
    return pkg, ours_model, ref_params, x


def _one_step(model, holder, x, ours, ref):
    ours.zero_grad(set_to_none=True)
    out = model(**x)
    (sum(o.float().pow(2).mean() for o in out[2:]) + out[1].mean()).backward()
    for pm, ph in zip(model.parameters(), holder.parameters()):
        ph.grad = None if pm.grad is None else pm.grad.detach().clone()
    ours.step()
    ref.step()


def _assert_same(model, holder, what):
    for (n, pm), ph in zip(model.named_parameters(), holder.parameters()):
        torch.testing.assert_close(pm, ph, rtol=1e-5, atol=1e-8, msg=lambda m: f'{what}, {n}: {m}')


@pytest.mark.parametrize('weight_decay', [0.0, 0.01])
def test_flat_adam_matches_torch_adam(weight_decay):
    pkg, model, holder, x = _setup()
    hp = dict(lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=weight_decay)
    ref = torch.optim.Adam(holder.parameters(), **hp)
    ours = pkg.optim.FlatAdam(model, **hp)
    names = [n for n, _ in model.named_parameters()]
    for it in range(6):
        _one_step(model, holder, x, ours, ref)
        _assert_same(model, holder, f'step {it}')
    # parameters keep their identity, names and shapes; the storage of the trained ones is one buffer now
    assert [n for n, _ in model.named_parameters()] == names
    base = ours._p.data_ptr()
    assert all(base <= p.data_ptr() < base + 4 * ours._p.numel() for p in model.parameters() if p.grad is not None)
    # state_dict interchanges with torch.optim.Adam's, both ways
    fresh_ref = torch.optim.Adam(holder.parameters(), **hp)
    fresh_ref.load_state_dict(ours.state_dict())
    fresh_ours = pkg.optim.FlatAdam(model, **hp)
    fresh_ours.load_state_dict(ref.state_dict())
    _one_step(model, holder, x, fresh_ours, fresh_ref)
    _assert_same(model, holder, 'after the state_dict exchange')
