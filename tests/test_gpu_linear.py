"""GPU: the projection kernels (tggcn_linear_fwd) against a float64 CPU matmul, including ragged tiles,
strided operands and sliced outputs as the forward uses them."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # M, N, K, lda_pad, ldw_pad, ldc_pad, relu, bias
    (1, 16, 16, 0, 0, 0, 0, 1),
    (37, 48, 32, 0, 0, 0, 1, 1),
    (130, 130, 64, 8, 4, 3, 0, 1),
    (128, 128, 32, 0, 0, 0, 0, 0),
    (300, 200, 96, 0, 0, 0, 1, 1),
    (257, 96, 48, 104, 0, 32, 1, 0),
    (2048, 512, 2048, 104, 0, 512, 1, 1),      # human ROI embedding, B=8,T=128
    (1024, 2048, 3328, 0, 0, 0, 1, 1),         # geometry MLP layer 0
    (4096, 1536, 512, 512, 0, 1536, 0, 1),     # BiGRU input gates, objects
    (5000, 4000, 96, 8, 4, 32, 1, 1),          # enough 128x256 tiles for the wide-tile tcgen05 variant, ragged in M and N
    (4096, 3072, 64, 0, 0, 0, 0, 0),
]


def _linear(pkg, A, W, bias, C_out, M, N, K, relu, path):
    lib = pkg.abi.lib()
    rc = lib.tggcn_linear_fwd(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0),
                              bias.data_ptr() if bias is not None else None, C_out.data_ptr(), C_out.stride(0),
                              M, N, K, relu, path, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg.abi.check(rc, 'tggcn_linear_fwd')


@pytest.mark.parametrize('path', [0, 1, 3])
@pytest.mark.parametrize('shape', SHAPES)
def test_linear_matches_fp64(shape, path, pkg):
    """path 0: fp32 SIMT, 1: tcgen05 3xTF32 (fp32-class), 3: tcgen05 with bf16 operands (dims.precision = 1; SIMT where K % 32 != 0)."""
    M, N, K, pa, pw, pc, relu, has_bias = shape
    if path == 1 and K % 32 != 0:
        pytest.skip('tcgen05 path needs K % 32 == 0')
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A_full = torch.randn(M, K + pa, generator=g)
    W_full = torch.randn(N, K + pw, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g) if has_bias else None
    ref = A_full[:, :K].double() @ W_full[:, :K].double().t()
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    A_d, W_d = A_full.cuda(), W_full.cuda()
    C_full = torch.full((M, N + pc), -7.0, device='cuda')
    _linear(pkg, A_d[:, :K], W_d[:, :K], None if bias is None else bias.cuda(), C_full[:, :N], M, N, K, relu, path)
    torch.cuda.synchronize()
    got = C_full[:, :N].cpu().double()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    if path == 3 and K % 32 == 0:
        # bf16 operands (8 mantissa bits each), fp32 accumulation: the error of a length-K dot product of unit-variance terms is
        # ~2^-8 * sqrt(K) * |w|; bound it by 2 % of the largest output and require the error to be unbiased
        assert err <= 2e-2 * scale, f'bf16 path: max err {err:.3e} (scale {scale:.3e})'
        assert (got - ref).mean().abs().item() <= 2e-4 * scale
        assert err > 1e-6 * scale, 'bf16 path requested but the result is fp32-exact: the bf16 kernel did not run'
    else:
        assert err <= 2e-5 * scale + 1e-5, f'max err {err:.3e} (scale {scale:.3e})'
    if pc:
        assert torch.all(C_full[:, N:] == -7.0), 'wrote outside the N columns'


SHAPES16 = [
    # M, N, K, lda_pad, ldw_pad, ldc_pad, relu, bias        (K % 64 == 0, N % 16 == 0: what the forward sends to gemm16.cu)
    (1, 16, 64, 0, 0, 0, 0, 1),
    (130, 144, 64, 8, 4, 4, 0, 1),
    (300, 208, 192, 0, 0, 0, 1, 1),
    (257, 96, 128, 104, 0, 32, 1, 0),
    (2048, 512, 2048, 104, 0, 512, 1, 1),      # human ROI embedding, B=8,T=128
    (1024, 2048, 3328, 0, 0, 0, 1, 1),         # geometry MLP layer 0
    (4096, 1536, 512, 512, 0, 1536, 0, 1),     # BiGRU input gates, objects
    (5000, 4000, 128, 8, 4, 32, 1, 1),         # many 128x256 tiles, ragged in M and N
    (4096, 3072, 64, 0, 0, 0, 0, 0),
]


@pytest.mark.parametrize('precision', [0, 1])
@pytest.mark.parametrize('shape', SHAPES16)
def test_linear16_matches_fp64(shape, precision, pkg):
    """The TMA-fed tcgen05 projection kernel on 16-bit operand planes (csrc/gemm16.cu): precision 0 = fp16 (hi, lo) split,
    fp32-class accuracy (same bound as the 3xTF32 kernel); 1 = bf16 operands.  Rows of very different magnitudes exercise the
    split (small values keep their absolute accuracy through the lo plane's subnormals)."""
    M, N, K, pa, pw, pc, relu, has_bias = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + 1)
    A_full = torch.randn(M, K + pa, generator=g)
    A_full[::3] *= 30.0
    A_full[1::3] *= 1e-3
    W_full = torch.randn(N, K + pw, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g) if has_bias else None
    ref = A_full[:, :K].double() @ W_full[:, :K].double().t()
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    A_d, W_d = A_full.cuda(), W_full.cuda()
    C_full = torch.full((M, N + pc), -7.0, device='cuda')
    lib = pkg.abi.lib()
    nbytes = lib.tggcn_linear16_scratch_bytes(M, N, K)
    scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device='cuda')
    sptr = (scratch.data_ptr() + 255) // 256 * 256
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    rc = lib.tggcn_linear16_fwd(A_d.data_ptr(), A_d.stride(0), W_d.data_ptr(), W_d.stride(0),
                                bias.cuda().data_ptr() if bias is not None else None, C_full.data_ptr(), C_full.stride(0),
                                M, N, K, relu, precision, C.c_void_p(sptr), nbytes, C.c_void_p(status.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg.abi.check(rc, 'tggcn_linear16_fwd')
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    got = C_full[:, :N].cpu().double()
    # per-row scale: the rows differ by 4.5 orders of magnitude
    scale = ref.abs().amax(dim=1, keepdim=True) + 1e-6
    rel = ((got - ref).abs() / scale).max().item()
    if precision == 1:
        assert rel <= 2e-2, f'bf16: max row-relative err {rel:.3e}'
        assert rel > 1e-6, 'bf16 requested but the result is fp32-exact'
    else:
        big = (got - ref).abs()[::3].max().item() / (ref.abs()[::3].max().item() + 1e-6)
        assert big <= 2e-5, f'fp16 split: max err on the large rows {big:.3e}'
        # the small rows keep fp32-class ABSOLUTE accuracy (their lo parts are fp16 subnormals)
        assert ((got - ref).abs()[1::3].max().item() if M > 1 else 0.0) <= 2e-5 * (ref.abs().max().item() + 1e-6) + 1e-5
        assert rel <= 5e-4, f'fp16 split: max row-relative err {rel:.3e}'
    if pc:
        assert torch.all(C_full[:, N:] == -7.0), 'wrote outside the N columns'


def test_linear16_reports_range_violation(pkg):
    M, N, K = 64, 32, 64
    A = torch.ones(M, K, device='cuda')
    A[5, 7] = 1e6                                   # beyond fp16
    W = torch.ones(N, K, device='cuda') * 0.01
    Cout = torch.empty(M, N, device='cuda')
    lib = pkg.abi.lib()
    nbytes = lib.tggcn_linear16_scratch_bytes(M, N, K)
    scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device='cuda')
    sptr = (scratch.data_ptr() + 255) // 256 * 256
    status = torch.zeros(1, dtype=torch.int32, device='cuda')
    rc = lib.tggcn_linear16_fwd(A.data_ptr(), K, W.data_ptr(), K, None, Cout.data_ptr(), N, M, N, K, 0, 0, C.c_void_p(sptr), nbytes,
                                C.c_void_p(status.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg.abi.check(rc, 'tggcn_linear16_fwd')
    torch.cuda.synchronize()
    assert int(status.item()) & 2


BWD_SHAPES = [
    # M, N, K, relu, lda_pad, ldy_pad
    (37, 48, 32, 1, 0, 0),
    (300, 64, 96, 0, 8, 16),
    (1030, 512, 1024, 1, 0, 512),
    (2048, 1536, 512, 0, 512, 1536),
]


@pytest.mark.parametrize('path', [0, 1])
@pytest.mark.parametrize('shape', BWD_SHAPES)
def test_linear_backward_matches_autograd(shape, path, pkg):
    M, N, K, relu, pa, py = shape
    if path == 1 and (K % 32 or N % 32):
        pytest.skip('tcgen05 path needs multiples of 32')
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    X = torch.randn(M, K + pa, generator=g)
    W = (torch.randn(N, K, generator=g) / K ** 0.5)
    b = torch.randn(N, generator=g) * 0.1
    dY = torch.randn(M, N + py, generator=g)
    dX0 = torch.randn(M, K, generator=g)
    x64, w64, b64 = X[:, :K].double().requires_grad_(), W.double().requires_grad_(), b.double().requires_grad_()
    y64 = x64 @ w64.t() + b64
    if relu:
        y64 = y64.clamp_min(0)
    y64.backward(dY[:, :N].double())
    Xd, Wd, dYd = X.cuda(), W.cuda(), dY.cuda()
    Yd = y64.detach().float().cuda() if relu else None
    dX = dX0.clone().cuda()
    dW = torch.empty(N, K, device='cuda')
    db = torch.empty(N, device='cuda')
    wt = torch.empty(K * N, device='cuda')
    lib = pkg.abi.lib()
    rc = lib.tggcn_linear_bwd(dYd.data_ptr(), dYd.stride(0), Yd.data_ptr() if relu else None, N, Xd.data_ptr(), Xd.stride(0),
                              Wd.data_ptr(), K, dX.data_ptr(), K, 1, dW.data_ptr(), K, db.data_ptr(), wt.data_ptr(), M, N, K, path,
                              C.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg.abi.check(rc, 'tggcn_linear_bwd')
    torch.cuda.synchronize()
    for name, got, want in (('dX', dX.cpu().double() - dX0.double(), x64.grad), ('dW', dW.cpu().double(), w64.grad),
                            ('db', db.cpu().double(), b64.grad)):
        err = (got - want).abs().max().item()
        scale = want.abs().max().item() + 1e-6
        assert err <= 3e-5 * scale + 1e-5, f'{name}: max err {err:.3e} (scale {scale:.3e})'
