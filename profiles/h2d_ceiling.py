import torch, time
n = 51516160
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(5): d.copy_(h, non_blocking=True)
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(50): d.copy_(h, non_blocking=True)
    e1.record(s)
    s.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"one 51.5 MB pinned H2D copy: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s  -> ceiling {64 / ms * 1e3:.0f} samples/s at 64 samples per copy")
# 8 pieces like the real input set
sizes = [4194304, 16777216, 16777216, 13762560, 768, 2304, 768, 1024]
hs = [torch.empty(x, dtype=torch.uint8).pin_memory() for x in sizes]
ds = [torch.empty(x, dtype=torch.uint8, device="cuda") for x in sizes]
with torch.cuda.stream(s):
    e0.record(s)
    for _ in range(50):
        for a, b in zip(ds, hs): a.copy_(b, non_blocking=True)
    e1.record(s)
    s.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"same bytes as 8 copies: {ms:.3f} ms = {sum(sizes) / ms / 1e6:.1f} GB/s -> ceiling {64 / ms * 1e3:.0f} samples/s")
