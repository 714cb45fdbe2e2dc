"""CPU: the oracle (oracle/tggcn_oracle.py) against every golden vector produced by running the
unmodified reference (oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from golden_util import CASES, FULL_CASES, GoldenCase
import importlib


def _state(case, dtype):
    pkg = importlib.import_module('2g-gcn_b200')
    model = pkg.TGGCN(**case.kwargs)
    sd = case.fill(model.state_dict())
    return {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}


@pytest.mark.parametrize('name', sorted(CASES) + sorted(FULL_CASES))
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_oracle_matches_reference(name, dtype, orc):
    case = GoldenCase(name)
    p = _state(case, dtype)
    cast = lambda t: None if t is None else t.to(dtype)
    b = case.batch
    res = orc.forward(p, case.ocfg, cast(b['x_human']), cast(b['x_objects']), cast(b['objects_mask']), cast(case.hseg),
                      cast(case.oseg), cast(case.noise), training=case.train_mode, inspect_model=case.inspect,
                      steps_per_example=b['steps_per_example'], distances=None if case.dists is None else tuple(None if d is None else cast(d) for d in case.dists))
    out, att = (res if case.inspect else (res, None))
    assert len(out) == len(case.outputs)
    n_gate = 2 if case.shape.num_classes[1] is None else 4
    for i, (o, g) in enumerate(zip(out, case.outputs)):
        assert tuple(o.shape) == tuple(g.shape)
        if i < n_gate:      # hard / soft gates: discrete decisions must be identical
            torch.testing.assert_close(o.float(), g, rtol=0, atol=2e-6)
        else:               # log-probabilities: 1e-3 relative is north_star's bar; the oracle is far inside it
            torch.testing.assert_close(o.float(), g, rtol=1e-4, atol=1e-5)
            assert torch.equal(o.argmax(1), g.argmax(1))
    if att is not None:
        for i, a in enumerate(att):
            torch.testing.assert_close(a.float(), torch.from_numpy(case.blob[f'att{i}']), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('name', sorted(CASES) + sorted(FULL_CASES))
def test_oracle_losses_and_f1(name, orc):
    case = GoldenCase(name)
    losses = orc.multi_task_loss(case.outputs, case.targets, case.shape.dataset, case.stage)
    got = np.array([float(l) for l in losses])
    np.testing.assert_allclose(got, case.blob['losses'], rtol=1e-5, atol=1e-7)
    rec_idx = 4 if case.shape.num_classes[1] is None else 8
    pred = case.outputs[rec_idx].argmax(dim=1).numpy()
    tgt = case.targets[rec_idx].numpy()
    f1 = [orc.f1_at_k(orc.labels_for_f1(tgt), orc.labels_for_f1(pred), case.shape.num_classes[0], k) for k in (0.10, 0.25, 0.50)]
    np.testing.assert_allclose(f1, case.blob['f1'], rtol=0, atol=1e-12)


def test_oracle_gcn_and_bn_update(orc):
    case = GoldenCase('mphoi_s2_train_bn')
    p = _state(case, torch.float64)
    taps = {}
    b = case.batch
    orc.forward(p, case.ocfg, b['x_human'].double(), b['x_objects'].double(), b['objects_mask'].double(), None, None,
                case.noise.double(), training=True, taps=taps)
    np.testing.assert_allclose(taps['gcn_out'].float().numpy(), case.blob['gcn_out'], rtol=1e-4, atol=1e-5)
    key = 'geometry_embedding_gcn.joint_embed.cnn.0.bn.'
    new_mean = 0.9 * p[key + 'running_mean'] + 0.1 * taps['batch_mean']
    new_var = 0.9 * p[key + 'running_var'] + 0.1 * taps['batch_var_unbiased']
    np.testing.assert_allclose(new_mean.float().numpy(), case.blob['bn_after.running_mean'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(new_var.float().numpy(), case.blob['bn_after.running_var'], rtol=1e-5, atol=1e-6)


def test_reorder_and_filter_edge_cases(orc):
    # no segment end at all: every frame keeps its own state; all ends: identity
    hx = torch.arange(12.0).reshape(1, 6, 2)
    assert torch.equal(orc.reorder(hx, torch.zeros(1, 6)), hx)
    assert torch.equal(orc.reorder(hx, torch.ones(1, 6)), hx)
    u = torch.tensor([[0., 0., 1., 0., 1., 0.]])
    exp = hx[:, [2, 2, 2, 4, 4, 5]]
    assert torch.equal(orc.reorder(hx, u), exp)
    y = torch.tensor([[0.05, 0.3, 0.2, 0.2, 0.6, 0.7]])
    f = orc.filter_soft(y, 0.1)
    assert f.ne(0).tolist() == [[False, True, False, False, False, True]]


from golden_util import GRAD_CASES, FULL_GRAD_CASES, alias_shared_heads, dists64      # noqa: E402


def _summarize(g):
    f = g.detach().double().reshape(-1)
    if f.numel() <= 4096:
        return f.numpy()
    idx = torch.linspace(0, f.numel() - 1, 256).long()
    return np.concatenate([[float(f.sum()), float(f.abs().sum()), float((f * f).sum())], f[idx].numpy()])


@pytest.mark.parametrize('name', sorted(GRAD_CASES) + sorted(FULL_GRAD_CASES))
def test_oracle_gradients_match_reference(name, orc, synth, pkg):
    """Backward pins for the next round: autograd through the oracle (train mode, multi_task_loss summed) against the
    gradients of the unmodified reference (oracle/gen_golden.py::run_grad_case), including which parameters get none."""
    import os
    from golden_util import GOLDEN_DIR
    spec = GRAD_CASES[name] if name in GRAD_CASES else FULL_GRAD_CASES[name]
    shape_name, D, B, T, stage = spec[:5]
    extra = spec[5] if len(spec) > 5 else {}
    blob = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    data_seed, noise_seed, target_seed, weight_seed = [int(v) for v in blob['meta']]
    shape = synth.SHAPES[shape_name]
    kw = synth.model_kwargs(shape, hidden_size=D, stage=stage, **extra)
    model = pkg.TGGCN(**kw)
    synth.deterministic_fill(model.state_dict(), seed=weight_seed, gain=float(blob['gain'][0]))
    assert abs(synth.state_checksum(model.state_dict()) - float(blob['weights_checksum'][0])) < 1e-6
    p = {k: v.detach().double().requires_grad_(v.is_floating_point() and 'running' not in k)
         if v.is_floating_point() else v for k, v in model.state_dict().items()}
    alias_shared_heads(p, extra)
    batch = synth.make_batch(shape, B, T, seed=data_seed)
    human_given, objects_given = stage == 1, stage == 1 and shape.dataset == 'cad120'
    n_calls = orc.num_noise_draws(T, shape.H, shape.O, human_given, objects_given, kw['object_segment_update_strategy'],
                                      kw['discrete_optimization_strategy'] in ('st', 'straight-through'))
    noise = orc.draw_noise(max(n_calls, 1), B, torch.Generator().manual_seed(noise_seed))[:n_calls]
    hseg = torch.ones(B, T, shape.H).double() if human_given else None
    oseg = torch.ones(B, T, shape.O).double() if objects_given else None
    ocfg = orc.config_from_kwargs(kw)
    out = orc.forward(p, ocfg, batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(),
                      hseg, oseg, noise.double() if n_calls else None, training=True, steps_per_example=batch['steps_per_example'],
                      distances=dists64(synth.make_distances(shape, B, T, seed=data_seed + 5000) if extra.get('_distances') else None))
    targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=target_seed))
    targets = [t.double() if t.is_floating_point() else t for t in targets]
    losses = orc.multi_task_loss(out, targets, shape.dataset, stage)
    total = sum(losses)
    np.testing.assert_allclose(float(total), float(blob['loss'][0]), rtol=1e-5)
    total.backward()
    none_ref = set(str(k) for k in blob['none_grad_keys'])
    params = dict(model.named_parameters())
    for k in params:
        g = p[k].grad
        if k in none_ref:
            assert g is None or float(g.abs().max()) == 0.0, k
            continue
        assert g is not None, k
        want = blob['grad.' + k]
        got = _summarize(g)
        scale = max(float(np.abs(want).max()), 1e-6)
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-5 * scale + 1e-7, err_msg=k)   # 1e-7: fp32 noise floor of exact zeros
