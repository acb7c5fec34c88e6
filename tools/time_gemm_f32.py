"""Time the fp32 CUDA-core GEMM (avhot_gemm, AVHOT_PREC_FP32) at the LSTM input-projection shapes of a near-tie flush."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_emotion_detection_b200 as pkg

lib = pkg._lib.load()
dev = torch.device("cuda:0")
for M, N, K in ((1565 * 301, 1024, 104), (1565 * 301, 1024, 256), (1565 * 16, 256, 12544)):
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    c = torch.empty(M, N, device=dev)
    def run():
        pkg._lib.check(lib.avhot_gemm(pkg._lib.ptr(a), pkg._lib.ptr(w), pkg._lib.ptr(bias), pkg._lib.ptr(c), M, N, K, 0,
                                      pkg._lib.PREC_FP32, pkg._lib.stream_ptr(dev)), "gemm")
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"M={M} N={N} K={K}: {ms:.3f} ms, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
