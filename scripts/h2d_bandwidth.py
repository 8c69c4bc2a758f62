"""Pinned host <-> device copy bandwidth of the box for the byte counts of one bench.py step (H2D 19 MB, D2H 3.5 MB)."""
import torch
x = torch.empty(19_070_566, dtype=torch.uint8).pin_memory()
y = torch.empty_like(x, device="cuda")
z = torch.empty(3_487_488, dtype=torch.uint8, device="cuda"); zh = torch.empty(3_487_488, dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream()
for name, src, dst in (("h2d 19MB", x, y), ("d2h 3.5MB", z, zh)):
    with torch.cuda.stream(s):
        for _ in range(3): dst.copy_(src, non_blocking=True)
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(20): dst.copy_(src, non_blocking=True)
        e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(name, "%.3f ms" % ms, "%.1f GB/s" % (src.numel() / ms / 1e6))
# many small chunks like the library's upload (4 arrays)
parts = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in (12_800_000, 6_400_000, 42_000, 520)]
dparts = [torch.empty_like(p, device="cuda") for p in parts]
with torch.cuda.stream(s):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20):
        for p, d in zip(parts, dparts): d.copy_(p, non_blocking=True)
    e1.record(s); s.synchronize()
    print("4-part upload %.3f ms" % (e0.elapsed_time(e1) / 20))
