"""Pinned host->device copy bandwidth on the bench's batch size: the floor of bench.py's e2e number."""
import torch, time
n = 1048576 * 1536
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunks in (1, 16):
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step = n // chunks
        for c in range(chunks): d[c*step:(c+1)*step].copy_(h[c*step:(c+1)*step], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"chunks={chunks} H2D {n/1e9:.2f} GB in {ms:.2f} ms = {n/ms/1e6:.1f} GB/s")
hb = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
db = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): hb.copy_(db, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"D2H 64 MiB: {(64<<20)*5/e0.elapsed_time(e1)/1e6:.1f} GB/s")
