import torch
x = torch.empty((3650, 1000000), dtype=torch.float64, device="cuda")
for _ in range(2): x.fill_(1.0)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.fill_(2.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("fill 29.2 GB: %.2f ms  %.0f GB/s" % (ms, x.numel() * 8 / ms / 1e6))
y = torch.empty_like(x)
e0.record()
for _ in range(3): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("copy 29.2 GB: %.2f ms  %.0f GB/s (r+w)" % (ms, 2 * x.numel() * 8 / ms / 1e6))
