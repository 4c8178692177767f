import torch, time
x = torch.empty(64 * 1024 * 1024 // 8, dtype=torch.float64).pin_memory()   # 64 MB
d = torch.empty_like(x, device="cuda")
for n in (1.6e6, 12.8e6, 64e6):
    k = int(n // 8)
    for _ in range(3): d[:k].copy_(x[:k], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): d[:k].copy_(x[:k], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"H2D pinned {n/1e6:.1f} MB: {dt*1e6:.1f} us -> {n/dt/1e9:.1f} GB/s")
