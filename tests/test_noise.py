"""CPU: the one-call Gumbel pre-draw equals the reference's per-call sampling
(pyrutils/torch/distributions.py:16 -> torch.distributions.gumbel.Gumbel(0,1).sample((B,2)))."""
import pytest
import torch


@pytest.mark.parametrize('B', [1, 2, 8, 13, 64, 128])
def test_batched_draw_equals_per_call_draws(B, pkg, orc):
    n_calls = 37
    torch.manual_seed(42)
    per_call = torch.stack([torch.distributions.gumbel.Gumbel(0.0, 1.0).sample((B, 2)) for _ in range(n_calls)])
    torch.manual_seed(42)
    batched = pkg.TGGCN.draw_gumbel_noise(n_calls, B)
    assert torch.equal(per_call, batched)
    g = torch.Generator().manual_seed(42)
    assert torch.equal(orc.draw_noise(n_calls, B, g), batched)
