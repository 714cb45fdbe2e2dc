"""GPU: the single-group BiGRU recurrence (forward) and its backward through time against fp64 autograd."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(gi, whh, bhh, G, orc):
    """fp64 recurrence with autograd; gi (B,T,E,2,3D)."""
    B, T, E, _, D3 = gi.shape
    D = D3 // 3
    gi = gi.double().requires_grad_()
    whh = [w.double().requires_grad_() for w in whh]
    bhh = [b.double().requires_grad_() for b in bhh]
    outs = []
    for d, order in enumerate((range(T), range(T - 1, -1, -1))):
        h = torch.zeros(B * E, D, dtype=torch.float64)
        hs = [None] * T
        for t in order:
            h = orc.gru_step(gi[:, t, :, d].reshape(B * E, D3), h, whh[d], bhh[d])
            hs[t] = h
        outs.append(torch.stack(hs, 1).reshape(B, E, T, D).permute(0, 2, 1, 3))
    hfr = torch.cat(outs, -1)                               # (B,T,E,2D)
    (hfr * G.double()).sum().backward()
    return hfr.detach(), gi.grad, [w.grad for w in whh], [b.grad for b in bhh]


@pytest.mark.parametrize('persistent', [1, 0])
@pytest.mark.parametrize('dims', [(3, 7, 2, 32), (8, 12, 4, 64), (5, 6, 1, 48), (9, 5, 5, 32),
                                  (20, 6, 9, 64), (11, 5, 3, 128)])      # the last two: several row blocks per weight slice
def test_bigru_forward_backward(dims, persistent, pkg, orc):
    B, T, E, D = dims
    g = torch.Generator().manual_seed(B * 100 + T * 10 + E)
    gi = torch.randn(B, T, E, 2, 3 * D, generator=g)
    whh = [torch.randn(3 * D, D, generator=g) / D ** 0.5 for _ in range(2)]
    bhh = [torch.randn(3 * D, generator=g) * 0.1 for _ in range(2)]
    G = torch.randn(B, T, E, 2 * D, generator=g)
    hfr_ref, dgi_ref, dw_ref, db_ref = _reference(gi, whh, bhh, G, orc)

    lib = pkg.abi.lib()
    dev = 'cuda'
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    gi_d, G_d = gi.to(dev), G.to(dev)
    whh_d, bhh_d = [w.to(dev) for w in whh], [b.to(dev) for b in bhh]
    hfr = torch.zeros(B, T, E, 2 * D, device=dev)
    gates = torch.zeros(B, T, E, 2, 4 * D, device=dev)
    sync = torch.zeros(4, dtype=torch.int32, device=dev)
    pkg.abi.check(lib.tggcn_bigru_fwd(gi_d.data_ptr(), whh_d[0].data_ptr(), whh_d[1].data_ptr(), bhh_d[0].data_ptr(),
                                      bhh_d[1].data_ptr(), hfr.data_ptr(), gates.data_ptr(), sync.data_ptr(), B, T, E, D,
                                      persistent, stream), 'tggcn_bigru_fwd')
    torch.cuda.synchronize()
    assert int(sync[1]) == 0, 'grid barrier timed out'
    torch.testing.assert_close(hfr.cpu().double(), hfr_ref, rtol=1e-4, atol=1e-5)

    dgi = torch.zeros_like(gi_d)
    dgh = torch.zeros_like(gi_d)
    dw = [torch.zeros(3 * D, D, device=dev) for _ in range(2)]
    db = [torch.zeros(3 * D, device=dev) for _ in range(2)]
    scratch = torch.zeros(lib.tggcn_bigru_bwd_scratch_floats(B, T, E, D), device=dev)
    pkg.abi.check(lib.tggcn_bigru_bwd(G_d.data_ptr(), hfr.data_ptr(), gates.data_ptr(), whh_d[0].data_ptr(), whh_d[1].data_ptr(),
                                      dgi.data_ptr(), dgh.data_ptr(), dw[0].data_ptr(), dw[1].data_ptr(), db[0].data_ptr(),
                                      db[1].data_ptr(), scratch.data_ptr(), B, T, E, D, 0, stream), 'tggcn_bigru_bwd')
    torch.cuda.synchronize()

    def close(name, got, want):
        err = (got.cpu().double() - want).abs().max().item()
        scale = want.abs().max().item() + 1e-6
        assert err <= 2e-4 * scale + 1e-5, f'{name}: max err {err:.3e} (scale {scale:.3e})'
    close('dgi', dgi, dgi_ref)
    for d in range(2):
        close(f'dW_hh[{d}]', dw[d], dw_ref[d])
        close(f'db_hh[{d}]', db[d], db_ref[d])
