"""GPU: the bf16 configuration (BASELINE.json configs[2], model.set_precision('bf16')): projections and the large-batch recurrent
kernels round their operands to bf16 and accumulate in fp32; gates, softmaxes, losses and the optimiser stay fp32.  Parity here is
STATISTICAL, against the fp32-class path of the same model on the same inputs and noise:
  * forward: log-probabilities close in the mean, per-frame argmax and hard segmentation gates agree on almost every frame;
  * training: the loss after k Adam steps stays within a few per cent of the fp32 run's, and both go down."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(pkg, synth, shape_name, D, B, T, mode):
    shape = synth.SHAPES[shape_name]
    torch.manual_seed(0)
    model = pkg.TGGCN(**synth.model_kwargs(shape, hidden_size=D, stage=2)).cuda()
    model.recurrent_mode = mode
    batch = synth.make_batch(shape, B, T, seed=9)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    noise = pkg.TGGCN.draw_gumbel_noise(T * (shape.H + shape.O), B)
    targets = [t.cuda() for t in synth.target_list(shape, synth.make_targets(shape, batch['lengths'], T, seed=10))]
    return shape, model, x, noise, targets


# (shape, D, B, T, recurrent_mode): mode 2 = bf16 recurrent step kernels as well, mode 1 = bf16 projections only
@pytest.mark.parametrize('cfg', [('mphoi', 128, 6, 24, 2), ('cad120', 256, 8, 20, 2), ('mphoi', 128, 4, 16, 1)])
def test_bf16_forward_tracks_fp32(cfg, pkg, synth):
    shape_name, D, B, T, mode = cfg
    shape, model, x, noise, _ = _setup(pkg, synth, shape_name, D, B, T, mode)
    model.eval()
    outs = {}
    for prec in ('fp32', 'bf16'):
        model.set_precision(prec)
        model.set_gumbel_noise(noise)
        with torch.no_grad():
            outs[prec] = [o.float().cpu() for o in model(**x)]
        model.check_persistent_kernels()
    n_gate = 2 if shape.num_classes[1] is None else 4
    ref, got = outs['fp32'], outs['bf16']
    assert any(not torch.equal(a, b) for a, b in zip(ref, got)), 'bf16 run is bit-identical to fp32: the bf16 kernels did not run'
    for i, (r, g) in enumerate(zip(ref, got)):
        assert torch.isfinite(g).all()
        if i < n_gate // 2:
            assert ((r != 0) == (g != 0)).float().mean() > 0.97, f'output {i}: hard gates'
        elif i < n_gate:
            assert (r - g).abs().mean() < 2e-2, f'output {i}: soft gates'
        else:
            assert (r - g).abs().mean() < 5e-2, f'output {i}: mean |dlogp| {(r - g).abs().mean():.3e}'
            assert (r.argmax(1) == g.argmax(1)).float().mean() > 0.9, f'output {i}: argmax agreement'


def test_bf16_training_tracks_fp32(pkg, synth, orc):
    shape_name, D, B, T = 'mphoi', 128, 6, 16
    curves = {}
    for prec in ('fp32', 'bf16'):
        shape, model, x, noise, targets = _setup(pkg, synth, shape_name, D, B, T, 2)
        model.train()
        model.set_precision(prec)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        losses = []
        for it in range(12):
            model.set_gumbel_noise(noise)
            opt.zero_grad(set_to_none=True)
            loss = sum(orc.multi_task_loss(model(**x), targets, 'mphoi', 2))
            loss.backward()
            opt.step()
            losses.append(float(loss))
        model.check_persistent_kernels()
        curves[prec] = losses
    a, b = curves['fp32'], curves['bf16']
    assert all(torch.isfinite(torch.tensor(b)))
    assert a[-1] < 0.9 * a[0] and b[-1] < 0.9 * b[0], (a, b)
    assert abs(b[0] - a[0]) <= 0.02 * abs(a[0]), (a[0], b[0])             # same weights, same batch: first loss within 2 %
    assert abs(b[-1] - a[-1]) <= 0.10 * abs(a[-1]), (a[-1], b[-1])       # after 12 Adam steps within 10 %
