"""Raw pinned host -> device copy rate of this box (development tool): bounds bench.py's e2e number."""
import torch
n = 618218496
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D {n/1e6:.0f} MB pinned: {ms:.2f} ms = {n/ms/1e6:.1f} GB/s")
o = torch.empty(34345472, dtype=torch.uint8).pin_memory()
s = torch.empty(34345472, dtype=torch.uint8, device="cuda")
e0.record()
for _ in range(5): o.copy_(s, non_blocking=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"D2H 34 MB pinned: {ms:.2f} ms = {34345472/ms/1e6:.1f} GB/s")
