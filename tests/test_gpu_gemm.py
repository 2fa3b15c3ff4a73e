"""GPU: the tcgen05 GEMM (through the C ABI) against a plain PyTorch fp32 reference of the same op on the
same bf16-rounded operands.  Tolerances: fp32 outputs differ only by accumulation order (rtol 2e-5 of the
row's |a|.|w| mass); bf16 outputs additionally by one bf16 rounding (2^-8 relative)."""
import ctypes

import pytest
import torch

from clip_based_cross_modal_hash_b200 import _lib

pytestmark = pytest.mark.gpu
EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_F32, EPI_TANH_F32 = range(5)


def run_gemm(a, w, bias, epi, resid=None):
    M, K = a.shape
    N = w.shape[0]
    out_bf16 = epi in (EPI_BF16, EPI_GELU_BF16)
    out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=a.device)
    if resid is not None:  # the residual epilogue is an in-place reduce-add into the stream
        out.copy_(resid)
        resid = out
    rc = _lib.lib().cmh_gemm_bf16(a.data_ptr(), M, K, a.stride(0), w.data_ptr(), N, w.stride(0),
                                  None if bias is None else bias.data_ptr(), epi, out.data_ptr(), out.stride(0),
                                  None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0),
                                  torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    return out


def gelu_slack(a, w, bias):
    """QuickGELU is evaluated as 0.5x + 0.5x*tanh(0.851x) with MUFU tanh.approx (relative error 2^-11 on the tanh):
    absolute error <= 2.5e-4 |x| on the pre-activation x, before the bf16 rounding of the result."""
    x = a.float() @ w.float().t() + (0 if bias is None else bias)
    return 3e-4 * x.abs() + 1e-6


def reference(a, w, bias, epi, resid=None):
    acc = a.float() @ w.float().t()
    if bias is not None:
        acc = acc + bias
    if epi == EPI_GELU_BF16:
        acc = acc * torch.sigmoid(1.702 * acc)
    if epi == EPI_TANH_F32:
        acc = torch.tanh(acc)
    if epi == EPI_RESID_F32:
        acc = acc + resid
    return acc


SHAPES = [
    (128, 256, 64), (128, 128, 128), (256, 768, 768), (100, 2304, 768), (12800, 2304, 768), (12800, 768, 3072),
    (8192, 1536, 512), (8192, 512, 2048), (50, 512, 768), (257, 64, 512), (300, 136, 192), (1, 16, 64), (777, 640, 320),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", [EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_F32, EPI_TANH_F32])
def test_gemm_matches_fp32_reference(M, N, K, epi):
    if M * N * K > 2e9 and epi in (EPI_TANH_F32, EPI_F32):
        pytest.skip("large shapes are covered by the other epilogues")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + epi)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == EPI_RESID_F32 else None
    got = run_gemm(a, w, bias, epi, resid).float()
    want = reference(a, w, bias, epi, resid)
    mass = a.float().abs() @ w.float().abs().t() + 1.0
    err = (got - want).abs()
    tol = 2e-5 * mass + (2.0 ** -8) * want.abs() * (1 if epi in (EPI_BF16, EPI_GELU_BF16) else 0) + 1e-6
    if epi == EPI_GELU_BF16:
        tol = tol + gelu_slack(a, w, bias)
    assert bool((err <= tol).all()), (float(err.max()), float((err / tol).max()))
    # no bias / in-place residual variants
    if epi == EPI_F32:
        got = run_gemm(a, w, None, epi)
        assert bool(((got - reference(a, w, None, epi)).abs() <= 2e-5 * mass + 1e-6).all())
@pytest.mark.parametrize("bn,cg", [(128, 1), (192, 1), (256, 1), (128, 2), (192, 2), (256, 2)])
@pytest.mark.parametrize("M,N,K,epi", [(12800, 2304, 768, EPI_BF16), (12800, 768, 3072, EPI_RESID_F32), (8192, 2048, 512, EPI_GELU_BF16),
                                       (300, 136, 192, EPI_F32), (1, 16, 64, EPI_F32), (257, 520, 320, EPI_BF16), (129, 768, 128, EPI_RESID_F32)])
def test_gemm_every_tile_configuration(bn, cg, M, N, K, epi):
    """Each (tile width, CTA group) kernel instance, pinned through cmh_gemm_force_tile, including ragged edges where the
    second CTA of a pair has few or no rows / columns."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + bn + cg)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == EPI_RESID_F32 else None
    _lib.check(_lib.lib().cmh_gemm_force_tile(bn, cg))
    try:
        got = run_gemm(a, w, bias, epi, resid).float()
    finally:
        _lib.check(_lib.lib().cmh_gemm_force_tile(0, 0))
    want = reference(a, w, bias, epi, resid)
    mass = a.float().abs() @ w.float().abs().t() + 1.0
    tol = 2e-5 * mass + (2.0 ** -8) * want.abs() * (1 if epi in (EPI_BF16, EPI_GELU_BF16) else 0) + 1e-6
    if epi == EPI_GELU_BF16:
        tol = tol + gelu_slack(a, w, bias)
    assert bool(((got - want).abs() <= tol).all()), float(((got - want).abs() / tol).max())


def test_gemm_repeated_launches_are_deterministic():
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(4096, 768, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(3072, 768, device="cuda", generator=g).to(torch.bfloat16)
    outs = [run_gemm(a, w, None, EPI_F32) for _ in range(5)]
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_gemm_rejects_bad_arguments():
    a = torch.zeros(8, 60, dtype=torch.bfloat16, device="cuda")   # K stride not a multiple of 8 elements
    w = torch.zeros(8, 60, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_lib.CmhError):
        run_gemm(a, w, None, EPI_F32)
