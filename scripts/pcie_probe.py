import torch, time
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, (a, b) in (('d2h', (h, d)), ('h2d', (d, h))):
    for _ in range(2):
        a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, 3 * n / (time.perf_counter() - t) / 1e9, 'GB/s')
