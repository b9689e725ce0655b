import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in (256 << 20, 10 << 20, 1 << 20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): d[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10): d[:n].copy_(x[:n], non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print("H2D %d MB: %.1f GB/s" % (n >> 20, 10 * n / a.elapsed_time(b) / 1e6))
