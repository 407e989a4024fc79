"""GPU unit tests of individual tensor-core kernels against a plain PyTorch fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, W, bias):
    from danspeech_b200 import _native as N
    M, K = A.shape
    Nn = W.shape[0]
    C = torch.empty((M, Nn), dtype=torch.float32, device=A.device)
    N.check(N.lib().dsb_gemm_bf16(N.ptr(A), A.stride(0), N.ptr(W), W.stride(0), N.ptr(bias), N.ptr(C), C.stride(0),
                                  M, Nn, K, N.current_stream()), "dsb_gemm_bf16")
    return C


@pytest.mark.parametrize("M,N,K", [(128, 240, 64), (256, 240, 128), (1000, 7200, 1200), (333, 1200, 2016),
                                   (4096, 2400, 400), (130, 33, 1200), (64, 96, 8), (777, 500, 1312)])
def test_gemm_bf16_matches_fp32_reference(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn((M, K), generator=g, device="cuda")).to(torch.bfloat16)
    W = (torch.randn((N, K), generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g, device="cuda")
    C = _gemm(A, W, bias)
    ref = A.float() @ W.float().t() + bias            # exact products of the bf16 inputs, fp32 accumulate
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, "tcgen05 GEMM differs from the fp32 reference: %g" % err
