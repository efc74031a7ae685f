"""What plain library kernels reach at K2's sizes on this GPU (floor for a ~100 MB streaming kernel):
memset of 64 MB, copy of 51 MB -> 51 MB, read-only sum of 38 MB; mean of 50 back-to-back launches."""
import torch

def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

MB = 1 << 20
for size in (64, 128, 512):
    a = torch.empty(size * MB // 4, dtype=torch.float32, device="cuda")
    us = timed(lambda: a.zero_())
    print("memset %4d MB: %7.2f us  %6.0f GB/s" % (size, us, size * MB / us / 1e3))
for size in (51, 128, 512):
    a = torch.empty(size * MB // 4, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    us = timed(lambda: b.copy_(a))
    print("copy   %4d MB: %7.2f us  %6.0f GB/s (read+write)" % (size, us, 2 * size * MB / us / 1e3))
for size in (38, 128, 512):
    a = torch.ones(size * MB // 4, dtype=torch.float32, device="cuda")
    us = timed(lambda: a.sum())
    print("sum    %4d MB: %7.2f us  %6.0f GB/s" % (size, us, size * MB / us / 1e3))
# K2-shaped mix: read 38 MB + write 64 MB in one elementwise op is not available in torch; the three above bracket it
