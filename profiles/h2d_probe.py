"""Host->device copy bandwidth of the GPU box (pinned and pageable), to interpret bench.py's e2e leg."""
import time
import torch
n = 104_202_240
for pinned in (True, False):
    h = torch.empty(n, dtype=torch.uint8)
    if pinned:
        h = h.pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(f"H2D {'pinned' if pinned else 'pageable'}: {n / dt / 1e9:.2f} GB/s ({dt * 1e3:.2f} ms for {n / 1e6:.0f} MB)")
    t = time.perf_counter()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(f"D2H {'pinned' if pinned else 'pageable'}: {n / dt / 1e9:.2f} GB/s")
