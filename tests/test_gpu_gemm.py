"""The library's GEMM (every nn.Linear on the path lowers to it): fp32 CUDA-core path at 1e-5, tcgen05 TF32 path at
its stated tolerance, on ragged shapes (partial M / N / K tiles)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TC_REL = 2e-3      # TF32 operands (10-bit mantissa), fp32 accumulate: stated tolerance of AVHOT_PREC_TC


def run(pkg, a, w, bias, relu, prec):
    dev = a.device
    M, K = a.shape
    N = w.shape[0]
    c = torch.full((M, N), float("nan"), device=dev)
    lib = pkg._lib.load()
    pkg._lib.check(lib.avhot_gemm(pkg._lib.ptr(a), pkg._lib.ptr(w), pkg._lib.ptr(bias), pkg._lib.ptr(c), M, N, K,
                                  1 if relu else 0, prec, pkg._lib.stream_ptr(dev)), "gemm")
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (256, 256, 64), (300, 1024, 104), (130, 128, 256), (1, 256, 8),
                                   (515, 384, 72), (2048, 256, 12544)])
@pytest.mark.parametrize("prec", ["fp32", "tc"])
def test_gemm(M, N, K, prec):
    import multimodal_emotion_detection_b200 as pkg
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.rand(M, K, generator=g) - 0.3
    w = (torch.rand(N, K, generator=g) - 0.5) / np.sqrt(K)
    bias = torch.rand(N, generator=g) - 0.5
    for relu in (False, True):
        want = a.double() @ w.double().T + bias.double()
        if relu:
            want = want.clamp_min(0)
        got = run(pkg, a.to(dev), w.to(dev), bias.to(dev), relu, pkg._lib.PREC_FP32 if prec == "fp32" else pkg._lib.PREC_TC)
        assert torch.isfinite(got).all()
        err = float((got.cpu().double() - want).abs().max() / want.abs().max())
        assert err <= (1e-5 if prec == "fp32" else TC_REL), err
