"""PCIe probe: pinned H2D / D2H bandwidth and small-copy latency on this box (context for the e2e numbers)."""
import time, torch
dev = torch.device("cuda", 0)
for mb in (0.05, 1, 5.3, 21, 64):
    n = int(mb * 1e6 / 8)
    h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); th = (time.perf_counter() - t0) / 50
    t0 = time.perf_counter()
    for _ in range(50): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); td = (time.perf_counter() - t0) / 50
    print(f"{mb:6.2f} MB: H2D {th*1e6:8.1f} us ({mb/1e3/th:5.1f} GB/s)   D2H {td*1e6:8.1f} us ({mb/1e3/td:5.1f} GB/s)")
