"""Device-side evaluation post-processing against the oracle restatement of pyrutils/metrics.py and the
torch/numpy formulation predict.py uses."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _labels(rng, B, T, E, C, seg=(3, 12)):
    y = np.zeros((B, T, E), dtype=np.int64)
    for b in range(B):
        for e in range(E):
            t = 0
            while t < T:
                n = int(rng.integers(seg[0], seg[1]))
                y[b, t:t + n, e] = int(rng.integers(0, C))
                t += n
    return y


@pytest.mark.parametrize('B,T,E,C,ds', [(5, 37, 2, 13, 1), (4, 23, 5, 10, 3), (3, 16, 9, 14, 2), (2, 5, 1, 4, 7)])
def test_upsample_argmax_matches_predict_py(pkg, B, T, E, C, ds):
    g = torch.Generator().manual_seed(B * 100 + T)
    out = torch.log_softmax(torch.randn(B, C, T, E, generator=g), dim=1)
    out[0, 1, :, 0] = 5.0                                   # ties at the maximum: the first one wins
    out[0, 2, :, 0] = 5.0
    for Tt in (T * ds, T * ds - 2, T * ds + 3):             # cropped and padded targets (match_shape)
        if Tt <= 0:
            continue
        tgt = torch.zeros(B, Tt, E, dtype=torch.int64)
        up = torch.repeat_interleave(out, repeats=ds, dim=-2)
        if up.size(2) >= Tt:
            up = up[:, :, :Tt]
        else:
            up = torch.cat([up, torch.repeat_interleave(up[:, :, -1:], Tt - up.size(2), dim=-2)], dim=-2)
        want = np.argmax(up.numpy(), axis=1)
        got = pkg.evaluate.predict_labels(out.cuda(), tgt.cuda(), ds).cpu().numpy()
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize('B,T,E,C', [(6, 120, 2, 13), (3, 64, 6, 12), (8, 33, 1, 10)])
def test_f1_at_k_matches_reference_metric(pkg, orc, B, T, E, C):
    rng = np.random.default_rng(B * 1000 + T)
    tgt = _labels(rng, B, T, E, C)
    prd = tgt.copy()
    flip = rng.random(tgt.shape) < 0.15                     # noisy prediction: short spurious segments + shifted boundaries
    prd[flip] = rng.integers(0, C, size=int(flip.sum()))
    prd = np.roll(prd, 2, axis=1)
    for b in range(B):                                      # padding beyond the video length, one fully padded row
        L = int(rng.integers(T // 2, T + 1))
        tgt[b, L:] = -1
    tgt[0, :, 0] = -1
    overlaps = (0.10, 0.25, 0.50)
    got = pkg.evaluate.f1_at_k(torch.from_numpy(tgt).cuda(), torch.from_numpy(prd).cuda(), C, overlaps)
    for ov in overlaps:
        want = orc.f1_at_k(orc.labels_for_f1(tgt.astype(np.float64)), orc.labels_for_f1(prd.astype(np.float64)), C, ov)
        assert got[ov] == pytest.approx(want, rel=0, abs=1e-15)


def test_golden_f1_reproduced_on_device(pkg):
    """The reference's own F1@{.10,.25,.50} of its own outputs (tests/golden) from the GPU path's outputs."""
    from golden_util import GoldenCase
    case = GoldenCase('mphoi_s2_eval')
    model = pkg.TGGCN(**case.kwargs)
    case.fill(model.state_dict())
    model = model.cuda().eval()
    model.set_gumbel_noise(case.noise)
    b = case.batch
    with torch.no_grad():
        out = model(x_human=b['x_human'].cuda(), x_objects=b['x_objects'].cuda(), objects_mask=b['objects_mask'].cuda())
    rec_idx = 4 if case.shape.num_classes[1] is None else 8          # segment-level recognition head, as test_gpu_parity.py
    tgt = case.targets[rec_idx].cuda()
    labels = pkg.evaluate.predict_labels(out[rec_idx], tgt, 1)
    assert torch.equal(labels.cpu(), case.outputs[rec_idx].argmax(1))  # identical per-frame labels to the reference's
    got = pkg.evaluate.f1_at_k(tgt, labels, case.shape.num_classes[0])
    np.testing.assert_allclose([got[0.10], got[0.25], got[0.50]], case.blob['f1'], rtol=0, atol=1e-12)
