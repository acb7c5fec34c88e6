import torch
dev = torch.device("cuda:0")
n = 12 * 1024**3 // 4
t = torch.empty(n, device=dev, dtype=torch.float32)
src = torch.empty(n, device=dev, dtype=torch.float32).normal_()
def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: t.fill_(1.0)); print("fill  (write only)  %.0f GB/s" % (n * 4 / ms / 1e6))
ms = timeit(lambda: t.zero_());    print("zero  (write only)  %.0f GB/s" % (n * 4 / ms / 1e6))
ms = timeit(lambda: t.copy_(src)); print("copy  (read+write)  %.0f GB/s" % (2 * n * 4 / ms / 1e6))
ms = timeit(lambda: src.sum());    print("sum   (read only)   %.0f GB/s" % (n * 4 / ms / 1e6))
h = src[: n // 2].half()
o = torch.empty(n // 2, device=dev, dtype=torch.float32)
ms = timeit(lambda: o.copy_(h)); print("f16->f32 (2B in, 4B out) %.0f GB/s" % (6 * (n // 2) / ms / 1e6))
