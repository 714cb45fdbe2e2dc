"""The fused criterion (2g-gcn_b200/losses.py): weights / names vs the reference's select_loss (CPU, needs no GPU), and on the
GPU the loss values and output gradients vs the oracle restatement of pyrutils/torch/losses.py (autograd)."""
import pytest
import torch


class Cfg(dict):
    def get(self, k, default_value=None):
        return dict.get(self, k, default_value)


STAGE_MISC = {
    1: dict(impose_segmentation_pattern=1, segmentation_loss=dict(add=False, sigma=0.0, weight=1.0)),
    2: dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0)),
}


@pytest.mark.parametrize('dataset', ['mphoi', 'cad120', 'bimanual'])
@pytest.mark.parametrize('stage', [1, 2])
def test_select_loss_weights_match_the_shipped_configs(dataset, stage, pkg, orc):
    crit, names = pkg.losses.select_loss('2G-GCN', 'multiple', dataset, Cfg(misc=STAGE_MISC[stage]))
    assert crit.weight == orc.loss_weights(dataset, stage)
    assert len(names) == len(crit.kinds) == (12 if dataset == 'cad120' else 6)
    assert pkg.losses.decide_num_main_losses('2G-GCN', dataset, STAGE_MISC[stage]) == (4 if dataset == 'cad120' else 2)
    with pytest.raises(ValueError):
        pkg.losses.select_loss('bimanual_baseline', 'multiple', dataset, None)


@pytest.mark.gpu
@pytest.mark.parametrize('dataset,stage', [('mphoi', 2), ('mphoi', 1), ('cad120', 2)])
def test_fused_losses_and_gradients_match_the_oracle(dataset, stage, pkg, orc, synth):
    shape = synth.SHAPES[dataset]
    B, T = 3, 11
    g = torch.Generator().manual_seed(7)
    batch = synth.make_batch(shape, B, T, seed=31)
    targets = synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=32))
    n_sub, n_aff = shape.num_classes
    def gates(E): return torch.rand(B, T, E, generator=g).clamp(1e-4, 1 - 1e-4)
    def logp(C, E): return torch.log_softmax(torch.randn(B, C, T, E, generator=g), dim=1)
    if n_aff is None:
        outs = [gates(shape.H), gates(shape.H)] + [logp(n_sub, shape.H) for _ in range(4)]
    else:
        outs = [gates(shape.H), gates(shape.O), gates(shape.H), gates(shape.O), logp(n_sub, shape.H), logp(n_sub, shape.H),
                logp(n_aff, shape.O), logp(n_aff, shape.O), logp(n_sub, shape.H), logp(n_sub, shape.H), logp(n_aff, shape.O),
                logp(n_aff, shape.O)]
    # non-trivial weights everywhere so that every term's gradient is exercised
    weights = [0.5 + 0.1 * i for i in range(len(outs))]
    crit, _ = pkg.losses.select_loss('2G-GCN', 'multiple', dataset, Cfg(misc=STAGE_MISC[stage]))
    crit.weight = weights
    # oracle (CPU, fp64 autograd)
    o64 = [o.double().requires_grad_() for o in outs]
    t64 = [t.double() if t.is_floating_point() else t for t in targets]
    fns = ([orc.budget_loss, orc.budget_loss, orc.bce_loss, orc.bce_loss] + [orc.nll_loss] * 8) if n_aff is not None else \
          ([orc.budget_loss, orc.bce_loss] + [orc.nll_loss] * 4)
    want = [w * fn(o, t) for o, t, fn, w in zip(o64, t64, fns, weights)]
    coef = torch.linspace(0.5, 1.5, len(want), dtype=torch.float64)
    (torch.stack(want) * coef).sum().backward()
    # fused
    og = [o.cuda().requires_grad_() for o in outs]
    got = crit(og, [t.cuda() for t in targets], reduction='mean')
    assert len(got) == len(want)
    for a, b in zip(got, want):
        torch.testing.assert_close(a.detach().cpu().double(), b.detach(), rtol=2e-5, atol=1e-6)
    (torch.stack(got) * coef.float().cuda()).sum().backward()
    for a, b in zip(og, o64):
        torch.testing.assert_close(a.grad.cpu().double(), b.grad, rtol=2e-4, atol=1e-7)


@pytest.mark.gpu
def test_fused_loss_all_targets_ignored_is_zero(pkg):
    out = [torch.rand(2, 5, 2, device='cuda', requires_grad=True), torch.log_softmax(torch.randn(2, 4, 5, 2, device='cuda'), 1).requires_grad_()]
    tgt = [torch.full((2, 5, 2), -1.0, device='cuda'), torch.full((2, 5, 2), -1, dtype=torch.int64, device='cuda')]
    losses = pkg.losses.multi_task_loss(out, tgt, [pkg.losses.BCE, pkg.losses.NLL], [1.0, 1.0])
    assert [float(l) for l in losses] == [0.0, 0.0]
    sum(losses).backward()
    assert float(out[0].grad.abs().max()) == 0.0 and float(out[1].grad.abs().max()) == 0.0


@pytest.mark.gpu
def test_fused_nll_class_index_out_of_range_is_loud(pkg):
    """F.nll_loss raises for a target >= C; the fused kernel must not read out of bounds and must not return a plausible
    number: the term turns NaN."""
    out = [torch.log_softmax(torch.randn(2, 4, 5, 2, device='cuda'), 1)]
    tgt = [torch.randint(0, 4, (2, 5, 2), device='cuda')]
    tgt[0][1, 3, 0] = 4
    losses = pkg.losses.multi_task_loss(out, tgt, [pkg.losses.NLL], [1.0])
    assert torch.isnan(losses[0])
