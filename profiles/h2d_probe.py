"""Host->device copy rate of this box for the e2e leg's 64 MB of head outputs: one stream vs several."""
import torch

MB = 1 << 20
src = torch.empty(64 * MB, dtype=torch.uint8).pin_memory()
dst = torch.empty(64 * MB, dtype=torch.uint8, device="cuda")

def run(nstreams, parts, reps=20):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    n = src.numel() // parts
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        for i in range(parts):
            with torch.cuda.stream(streams[i % nstreams]):
                dst[i * n:(i + 1) * n].copy_(src[i * n:(i + 1) * n], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("streams %d parts %2d: %.3f ms  %.1f GB/s" % (nstreams, parts, ms, src.numel() / ms / 1e6))

for ns, parts in ((1, 1), (1, 8), (2, 2), (2, 8), (4, 4), (4, 16)):
    run(ns, parts)
d2h = torch.empty(64 * MB, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    d2h.copy_(dst, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print("d2h: %.1f GB/s" % (64 * MB * 10 / e0.elapsed_time(e1) / 1e6))
