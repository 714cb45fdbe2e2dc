"""GPU: the CUDA path (through the C ABI, via the drop-in TGGCN class) against the golden vectors of the
reference and against the oracle's intermediates, on identical inputs, weights and Gumbel noise.

Tolerances (north_star): log-probabilities within 1e-3 relative, identical hard gates, identical
per-frame argmax labels, identical F1@{.10,.25,.50}.
"""
import importlib

import numpy as np
import pytest
import torch

from golden_util import CASES, GoldenCase, dist_kwargs, dists64

pytestmark = pytest.mark.gpu


def _build(case, persistent=True, gemm_path=0):
    pkg = importlib.import_module('2g-gcn_b200')
    model = pkg.TGGCN(**case.kwargs)
    case.fill(model.state_dict())
    model = model.cuda()
    model.train(case.train_mode)
    model.persistent_kernels = persistent
    model.gemm_path = gemm_path
    model.set_gumbel_noise(case.noise)
    return model


def _run(model, case, inspect=False):
    b = case.batch
    dev = 'cuda'
    with torch.no_grad():
        res = model(x_human=b['x_human'].to(dev), x_objects=b['x_objects'].to(dev), objects_mask=b['objects_mask'].to(dev),
                    human_segmentation=None if case.hseg is None else case.hseg.to(dev),
                    objects_segmentation=None if case.oseg is None else case.oseg.to(dev),
                    steps_per_example=b['steps_per_example'].to(dev), inspect_model=inspect,
                    **dist_kwargs(case.dists, lambda t: t.to(dev)))
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    return res


def _assert_close(name, got, want, rtol=1e-3, atol=1e-4):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert tuple(got.shape) == tuple(want.shape), f'{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}'
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = err > tol
    assert not bad.any(), f'{name}: {int(bad.sum())}/{bad.numel()} off, max abs err {float(err.max()):.3e} (max |ref| {float(want.abs().max()):.3e})'


@pytest.mark.parametrize('persistent', [True, False])
@pytest.mark.parametrize('name', sorted(CASES))
def test_forward_matches_reference_golden(name, persistent, orc):
    case = GoldenCase(name)
    model = _build(case, persistent)
    res = _run(model, case, inspect=case.inspect)
    out, att = (res if case.inspect else (res, None))
    assert len(out) == len(case.outputs)
    n_gate = 2 if case.shape.num_classes[1] is None else 4
    for i, (o, g) in enumerate(zip(out, case.outputs)):
        if i < n_gate:
            _assert_close(f'{name}.out{i} (gates)', o, g, rtol=0, atol=5e-6)
            if i < n_gate // 2:
                assert torch.equal(o.cpu() != 0, g != 0), f'{name}.out{i}: hard gates differ'
        else:
            _assert_close(f'{name}.out{i}', o, g, rtol=1e-3, atol=1e-4)
            assert torch.equal(o.argmax(1).cpu(), g.argmax(1)), f'{name}.out{i}: argmax labels differ'
    # F1@k through the oracle's restatement of pyrutils/metrics.py (itself pinned to the golden F1 values)
    rec_idx = 4 if case.shape.num_classes[1] is None else 8
    pred = out[rec_idx].argmax(dim=1).cpu().numpy()
    tgt = case.targets[rec_idx].numpy()
    f1 = [orc.f1_at_k(orc.labels_for_f1(tgt), orc.labels_for_f1(pred), case.shape.num_classes[0], k) for k in (0.10, 0.25, 0.50)]
    np.testing.assert_allclose(f1, case.blob['f1'], rtol=0, atol=1e-12)
    # losses computed from our outputs equal the reference's losses
    losses = orc.multi_task_loss([o.cpu() for o in out], case.targets, case.shape.dataset, case.stage)
    np.testing.assert_allclose([float(l) for l in losses], case.blob['losses'], rtol=2e-3, atol=1e-5)
    if att is not None:
        for i, a in enumerate(att):
            _assert_close(f'{name}.att{i}', a, torch.from_numpy(case.blob[f'att{i}']), rtol=1e-3, atol=1e-5)
    if case.train_mode:
        bn = model.geometry_embedding_gcn.joint_embed.cnn[0].bn
        _assert_close('bn.running_mean', bn.running_mean, torch.from_numpy(case.blob['bn_after.running_mean']), 1e-4, 1e-6)
        _assert_close('bn.running_var', bn.running_var, torch.from_numpy(case.blob['bn_after.running_var']), 1e-4, 1e-6)
        assert int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize('name', ['mphoi_s2_eval', 'cad120_s2_eval', 'bimanual_s2_eval', 'mphoi_s2_train_bn'])
def test_intermediates_match_oracle(name, orc):
    """Kernel-by-kernel localisation: every named workspace region against the oracle's taps (fp64)."""
    case = GoldenCase(name)
    model = _build(case)
    _run(model, case)
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in case.fill(
        importlib.import_module('2g-gcn_b200').TGGCN(**case.kwargs).state_dict()).items()}
    taps = {}
    b = case.batch
    dd = lambda t: None if t is None else t.double()
    orc.forward(p64, case.ocfg, dd(b['x_human']), dd(b['x_objects']), dd(b['objects_mask']), dd(case.hseg), dd(case.oseg),
                dd(case.noise), training=case.train_mode, taps=taps, steps_per_example=b['steps_per_example'], distances=dists64(case.dists))
    B, T, H, O, D, V = case.B, case.T, case.shape.H, case.shape.O, case.D, case.shape.V
    nkh = 2 if case.shape.hh else 1
    ws = model.workspace_tensor
    _assert_close('gcn_out', ws('GCN_OUT', (B, 128, V, T)), taps['gcn_out'])
    _assert_close('gcn_out vs reference', ws('GCN_OUT', (B, 128, V, T)), torch.from_numpy(case.blob['gcn_out']))
    for ent, E in (('h', H), ('o', O), ('g', 1)):
        s = ws('S_' + ent.upper(), (B, T, E, 2 * D))
        _assert_close('x_' + ent, s[..., :D], taps['x_' + ent])
        _assert_close('hfr_' + ent, ws('HFR_' + ent.upper(), (B, T, E, 2 * D)), taps['hfr_' + ent])
        _assert_close('h_' + ent, s[..., D:], taps['h_' + ent])
    ts = 1 if case.ocfg.time_position == 's' else 0
    _assert_close('xx_h', ws('XX_H', (B, T, H, (1 + nkh + ts) * D)), taps['xx_h'])
    _assert_close('xx_o', ws('XX_O', (B, T, O, (4 + ts) * D)), taps['xx_o'])
    _assert_close('hx_h', ws('HX_H', (B, T, H, 2 * D)), taps['hx_h'])
    _assert_close('hx_o', ws('HX_O', (B, T, O, 2 * D)), taps['hx_o'])


def test_default_noise_reproduces_manual_seed(pkg):
    """predict.py:21 seeds torch before the forward; the default noise path must be deterministic under it."""
    case = GoldenCase('mphoi_s2_eval')
    model = _build(case)
    model.set_gumbel_noise(None)
    torch.manual_seed(42)
    a = _run(model, case)
    torch.manual_seed(42)
    b = _run(model, case)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    torch.manual_seed(43)
    c = _run(model, case)
    assert not torch.equal(a[1], c[1])


def test_default_noise_equals_the_batched_cpu_draw(pkg):
    """The default path draws in place in pinned buffers (no stream sync of a pageable copy): same bits as draw_gumbel_noise, call
    after call (the ring has three buffers: run more calls than that)."""
    case = GoldenCase('mphoi_s2_eval')
    model = _build(case)
    n_calls = case.T * (case.shape.H + case.shape.O)
    torch.manual_seed(7)
    want = []
    for _ in range(5):
        model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(n_calls, case.B))
        want.append([o.clone() for o in _run(model, case)])
    model.set_gumbel_noise(None)
    torch.manual_seed(7)
    for w in want:
        got = _run(model, case)
        for x, y in zip(got, w):
            assert torch.equal(x, y)


def test_unmasked_objects_do_not_leak(pkg):
    """Features of masked (virtual) objects must not influence the humans' outputs."""
    case = GoldenCase('mphoi_s2_eval')
    model = _build(case)
    ref = _run(model, case)
    b = dict(case.batch)
    xo = b['x_objects'].clone()
    om = b['objects_mask']
    for i in range(om.size(0)):
        for k in range(om.size(1)):
            if om[i, k] == 0:
                xo[i, :, k] = torch.randn_like(xo[i, :, k])
    case.batch = {**b, 'x_objects': xo}
    out = _run(model, case)
    for i in (0, 1):
        assert torch.equal(out[i], ref[i])
    for i in range(2, 6):
        _assert_close(f'out{i}', out[i], ref[i], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('name', ['mphoi_s1_eval', 'mphoi_s2_eval', 'mphoi_s2_d64', 'cad120_s2_eval', 'mphoi_s2_train_bn'])
def test_forward_with_tcgen05_projections(name, orc):
    """Same parity bar with every nn.Linear on the tcgen05 3xTF32 kernel (gemm_path = 1)."""
    case = GoldenCase(name)
    model = _build(case, True, gemm_path=1)
    out = _run(model, case)
    n_gate = 2 if case.shape.num_classes[1] is None else 4
    for i, (o, g) in enumerate(zip(out, case.outputs)):
        if i < n_gate:
            _assert_close(f'{name}.out{i} (gates)', o, g, rtol=0, atol=5e-6)
            if i < n_gate // 2:
                assert torch.equal(o.cpu() != 0, g != 0), f'{name}.out{i}: hard gates differ'
        else:
            _assert_close(f'{name}.out{i}', o, g, rtol=1e-3, atol=1e-4)
            assert torch.equal(o.argmax(1).cpu(), g.argmax(1)), f'{name}.out{i}: argmax labels differ'
