"""GPU: parity AT THE SIZES bench.py MEASURES (BASELINE.json configs[1], [3], [4]) and at the row counts that select the
large-batch recurrent kernels — forward against the fp64 oracle on identical weights / inputs / Gumbel noise
(vhoi/models.py:584-933), backward against fp64 oracle autograd for every parameter, plus the reference-generated golden
vectors at hidden 512 (tests/golden/mphoi_s2_d512_full.npz, grad_mphoi_s2_d512.npz; oracle/gen_golden.py).

Tolerances (north_star): log-probabilities within 1e-3 relative (+1e-4 absolute), soft gates within 5e-6, identical hard
gates, identical per-frame argmax wherever the oracle's top-2 margin exceeds the tolerance, identical F1@{.10,.25,.50}.
Videos are independent in eval mode, so a video whose sampled gates all keep a margin from every discrete decision must match
in full; the few knife-edge videos of the biggest cases are compared on their soft gates only (seeds: tools/find_safe_seeds.py).
"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MARGIN = 2e-5

# name: (shape, D, B, T, stage, seed attempt of tools/find_safe_seeds.py, min share of knife-edge-free videos, recurrent_mode)
FWD_CASES = {
    'mphoi_bench': ('mphoi', 512, 8, 128, 2, 0, 1.0, 0),          # BASELINE.json configs[1] == bench.py's workload
    'cad120_long': ('cad120', 512, 8, 512, 2, 5, 1.0, 0),         # configs[3], first point of the sweep
    'bimanual_b32': ('bimanual', 64, 32, 256, 2, 0, 0.7, 0),      # configs[4], shipped hidden size
    'bimanual_d512_rows': ('bimanual', 512, 16, 24, 2, 0, 1.0, 2),   # 32 / 144 rows per step on the large-batch recurrent kernels
    'cad120_b64_rows': ('cad120', 512, 64, 16, 2, 1, 1.0, 0),        # 64 / 320 rows per step: large-batch path chosen automatically
}


def _case(name, orc, synth, pkg, table=FWD_CASES):
    shape_name, D, B, T, stage, attempt = table[name][:6]
    shape = synth.SHAPES[shape_name]
    kw = synth.model_kwargs(shape, hidden_size=D, stage=stage)
    model = pkg.TGGCN(**kw)
    synth.deterministic_fill(model.state_dict(), seed=11, gain=1.0)
    batch = synth.make_batch(shape, B, T, seed=500 + attempt)
    n_calls = orc.num_noise_draws(T, shape.H, shape.O, False, False)
    noise = orc.draw_noise(n_calls, B, torch.Generator().manual_seed(800 + attempt))
    ocfg = orc.OracleConfig(D, shape.V, shape.num_classes, shape.hh, stage == 2, kw['update_segment_threshold'])
    return shape, kw, model, batch, noise, ocfg


def _video_safe(soft, thr, filtered):
    """(B,) bool: no sampled gate of the video within MARGIN of flipping a discrete decision."""
    m = (soft - thr).abs() > MARGIN
    if filtered:
        pad = torch.zeros_like(soft[:, :1])
        m &= (soft - torch.cat([pad, soft[:, :-1]], 1)).abs() > MARGIN
        m &= (soft - torch.cat([soft[:, 1:], pad], 1)).abs() > MARGIN
    return m.flatten(1).all(1)


@pytest.mark.parametrize('name', list(FWD_CASES))
def test_benchmark_sizes_forward_matches_oracle(name, orc, synth, pkg):
    shape, kw, model, batch, noise, ocfg = _case(name, orc, synth, pkg)
    thr, need = kw['update_segment_threshold'], FWD_CASES[name][6]
    p64 = {k: v.detach().double() if v.is_floating_point() else v for k, v in model.state_dict().items()}
    taps = {}
    with torch.no_grad():
        ref = orc.forward(p64, ocfg, batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(),
                          None, None, noise.double(), taps=taps)
    model = model.cuda().eval()
    model.recurrent_mode = FWD_CASES[name][7]
    model.set_gumbel_noise(noise)
    with torch.no_grad():
        out = model(x_human=batch['x_human'].cuda(), x_objects=batch['x_objects'].cuda(), objects_mask=batch['objects_mask'].cuda())
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    out = [o.cpu() for o in out]
    cad = shape.num_classes[1] is not None
    soft_h = ref[2 if cad else 1].float()
    soft_o = taps['y_oss'].float()
    safe = _video_safe(soft_h, thr, True) & _video_safe(soft_o, thr, True)
    assert float(safe.float().mean()) >= need, f'only {int(safe.sum())} of {safe.numel()} videos are knife-edge free'
    n_gate = 4 if cad else 2
    for i, (o, r) in enumerate(zip(out, ref)):
        r = r.float()
        assert torch.isfinite(o).all(), f'output {i} not finite'
        if i < n_gate // 2:                       # hard gates
            assert torch.equal(o[safe] != 0, r[safe] != 0), f'output {i}: hard gates differ'
        elif i < n_gate:                          # soft gates: every video, they do not depend on the discrete path
            torch.testing.assert_close(o, r, rtol=0, atol=5e-6)
        else:
            torch.testing.assert_close(o[safe], r[safe], rtol=1e-3, atol=1e-4)
            top2 = r[safe].topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2e-4          # a class tie inside the tolerance may resolve either way
            assert torch.equal(o[safe].argmax(1)[clear], r[safe].argmax(1)[clear]), f'output {i}: argmax differs'
            assert float(clear.float().mean()) > 0.9        # random weights: padded frames / masked objects are class near-ties
    # F1@k of the segment-level recognition output against synthetic labels (predict.py:229-246 convention)
    T = batch['x_human'].shape[1]
    tg = synth.make_targets(shape, batch['lengths'], T, seed=900)
    rec = 8 if cad else 4
    tgt = tg['rec_h'][safe].numpy()
    for k in (0.10, 0.25, 0.50):
        f_ref = orc.f1_at_k(orc.labels_for_f1(tgt), orc.labels_for_f1(ref[rec].float()[safe].argmax(1).numpy()), shape.num_classes[0], k)
        f_got = orc.f1_at_k(orc.labels_for_f1(tgt), orc.labels_for_f1(out[rec][safe].argmax(1).numpy()), shape.num_classes[0], k)
        assert abs(f_ref - f_got) <= 1e-12, f'F1@{k}: {f_got} vs {f_ref}'


def test_reference_golden_at_hidden_512(orc, synth, pkg):
    """Outputs of the UNMODIFIED reference at the benchmarked configuration (MPHOI, B=8, T=128, hidden 512, stage 2)."""
    from golden_util import GoldenCase
    case = GoldenCase('mphoi_s2_d512_full')
    model = pkg.TGGCN(**case.kwargs)
    case.fill(model.state_dict())
    model = model.cuda().eval()
    model.set_gumbel_noise(case.noise)
    b = case.batch
    with torch.no_grad():
        out = model(x_human=b['x_human'].cuda(), x_objects=b['x_objects'].cuda(), objects_mask=b['objects_mask'].cuda())
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    for i, (o, g) in enumerate(zip(out, case.outputs)):
        o = o.cpu()
        if i == 0:
            assert torch.equal(o != 0, g != 0), 'hard gates differ from the reference'
        elif i == 1:
            torch.testing.assert_close(o, g, rtol=0, atol=5e-6)
        else:
            torch.testing.assert_close(o, g, rtol=1e-3, atol=1e-4)
            top2 = g.topk(2, dim=1).values
            clear = (top2[:, 0] - top2[:, 1]) > 2e-4
            assert torch.equal(o.argmax(1)[clear], g.argmax(1)[clear]), f'output {i}: argmax differs from the reference'
    tg = synth.make_targets(case.shape, b['lengths'], case.T, seed=int(case.blob['meta'][2]))
    pred = out[4].cpu().argmax(1).numpy()
    f1 = [orc.f1_at_k(orc.labels_for_f1(tg['rec_h'].numpy()), orc.labels_for_f1(pred), case.shape.num_classes[0], k)
          for k in (0.10, 0.25, 0.50)]
    np.testing.assert_allclose(f1, case.blob['f1'], rtol=0, atol=1e-12)


# ---- backward at hidden 512 -------------------------------------------------------------------------------------------------
BWD_CASES = {
    # name: (shape, D, B, T, stage, seed attempt)
    'mphoi_d512_T48': ('mphoi', 512, 8, 48, 2, 0),
    'mphoi_d512_T128': ('mphoi', 512, 8, 128, 2, 0),          # the train step bench.py times
}


@pytest.mark.parametrize('name', list(BWD_CASES))
def test_benchmark_sizes_backward_matches_oracle(name, orc, synth, pkg):
    shape, kw, model, batch, noise, ocfg = _case(name, orc, synth, pkg, BWD_CASES)
    T = BWD_CASES[name][3]
    targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=900))
    p = {k: (v.detach().double().requires_grad_('running' not in k) if v.is_floating_point() else v)
         for k, v in model.state_dict().items()}
    out64 = orc.forward(p, ocfg, batch['x_human'].double(), batch['x_objects'].double(), batch['objects_mask'].double(), None, None,
                        noise.double(), training=True)
    t64 = [t.double() if t.is_floating_point() else t for t in targets]
    want_losses = orc.multi_task_loss(out64, t64, shape.dataset, 2)
    sum(want_losses).backward()
    want = {k: v.grad for k, v in p.items() if torch.is_tensor(v) and v.is_floating_point()}

    model = model.cuda().train()
    model.set_gumbel_noise(noise)
    out = model(x_human=batch['x_human'].cuda(), x_objects=batch['x_objects'].cuda(), objects_mask=batch['objects_mask'].cuda())
    losses = orc.multi_task_loss(out, [t.cuda() for t in targets], shape.dataset, 2)
    np.testing.assert_allclose([float(l) for l in losses], [float(l) for l in want_losses], rtol=2e-4, atol=1e-6)
    sum(losses).backward()
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    bad, checked = [], 0
    for k, prm in model.named_parameters():
        w = want.get(k)
        if w is None or float(w.abs().max()) == 0.0:
            assert prm.grad is None or float(prm.grad.abs().max()) <= 1e-7, f'{k}: oracle has no gradient here'
            continue
        assert prm.grad is not None, f'{k}: gradient missing'
        got = prm.grad.detach().double().cpu()
        assert torch.isfinite(got).all(), k
        scale = float(w.abs().max())
        if scale < 1e-10:
            continue
        rel2 = float((got - w).norm() / w.norm())
        err = float((got - w).abs().max())
        checked += 1
        # L2 error tight; max error loose (a ReLU pre-activation within rounding of zero may sit on the other side of the kink).
        # 128 reverse steps of 3xTF32 products against fp64: 5e-3 at T = 128, 2e-3 (the bar of tests/test_gpu_backward.py) below
        if not (rel2 <= (5e-3 if T > 64 else 2e-3) and err <= 0.1 * scale + 1e-7):
            bad.append(f'{k}: rel L2 err {rel2:.3e}, max err {err:.3e} vs scale {scale:.3e}')
    assert checked > 80 and not bad, '\n'.join(bad)


def test_reference_gradients_at_hidden_512(orc, synth, pkg):
    """Gradients of the UNMODIFIED reference's train-mode forward -> multi_task_loss -> backward at hidden 512."""
    name = 'grad_mphoi_s2_d512'
    tb = importlib.import_module('test_gpu_backward')
    c = tb._setup(name, orc, synth, pkg)
    blob = c['blob']
    model = c['model'].cuda().train()
    model.set_gumbel_noise(c['noise'])
    b = c['batch']
    out = model(x_human=b['x_human'].cuda(), x_objects=b['x_objects'].cuda(), objects_mask=b['objects_mask'].cuda())
    losses = orc.multi_task_loss(out, [t.cuda() for t in c['targets']], c['shape'].dataset, c['stage'])
    total = sum(losses)
    np.testing.assert_allclose(float(total), float(blob['loss'][0]), rtol=2e-4)
    total.backward()
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    none_ref = set(str(k) for k in blob['none_grad_keys'])
    for k, prm in model.named_parameters():
        if k in none_ref:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, f'{k}: the reference gives no gradient'
            continue
        assert prm.grad is not None, f'{k}: gradient missing'
        ref = blob['grad.' + k]
        rscale = max(float(np.abs(ref).max()), 1e-6)
        # the fixture is the reference's fp32 autograd: at this size it carries ~3e-4 (of a tensor's largest entry) of
        # summation noise of its own — the CUDA gradients sit closer to the fp64 oracle (test above) than the fixture does
        # A ReLU pre-activation within fp32 rounding of zero may sit on the other side of the kink in the reference's fp32 run and
        # move single entries (seen: 1 of 512 entries of a bias gradient by 5e-3 of the largest entry): entries outside the
        # tolerance must be rare (<= 0.5 %) and bounded (<= 0.1 of the largest entry) — the rule of the oracle comparison above
        got = tb._summarize(prm.grad)
        diff = np.abs(got - ref)
        outside = diff > 1e-2 * np.abs(ref) + 1e-3 * rscale + 1e-7
        assert outside.mean() <= 0.005 and diff.max() <= 0.1 * rscale + 1e-7, \
            f'{k}: {int(outside.sum())} of {outside.size} entries outside tolerance, max diff {diff.max():.3e} vs scale {rscale:.3e}'
