"""Pure-write and copy bandwidth of the device (context for the scoring kernel's full-matrix mode)."""
import torch
n = 723 * 1000000
x = torch.empty(n, device="cuda", dtype=torch.float32)
y = torch.empty(n, device="cuda", dtype=torch.float32)
def t(fn, it=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.fill_(1.5)); print("fill_  %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_  %.3f ms  %.0f GB/s written" % (ms, n * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy_  %.3f ms  %.0f GB/s read+written" % (ms, 2 * n * 4 / ms / 1e6))
ms = t(lambda: x.sum()); print("sum    %.3f ms  %.0f GB/s read" % (ms, n * 4 / ms / 1e6))
