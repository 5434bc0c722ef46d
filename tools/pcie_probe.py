import torch, time
d = torch.empty(700_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(700_000_000, dtype=torch.uint8).pin_memory()
for name, src, dst in (("d2h", d, h), ("h2d", h, d)):
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(name, f"{0.7 / dt:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(700_000_000, dtype=torch.uint8).pin_memory(); d2 = torch.empty_like(d)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
print("bidirectional each", f"{0.7 / dt:.1f} GB/s")
import os; print("cpus", os.cpu_count())
