import torch, time
n = 370 * 1024 * 1024 // 4
d = torch.empty(n, dtype=torch.float32, device="cuda")
h = torch.empty(n, dtype=torch.float32, pin_memory=True)
for _ in range(2): h.copy_(d, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"D2H pinned {n*4/1e6:.0f} MB: {dt*1e3:.2f} ms -> {n*4/dt/1e9:.1f} GB/s")
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"H2D pinned: {dt*1e3:.2f} ms -> {n*4/dt/1e9:.1f} GB/s")
