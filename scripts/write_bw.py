import torch
y = torch.empty(256 * 400 * 80 * 128, dtype=torch.bfloat16, device='cuda')
x = torch.empty_like(y)
for name, fn in (('fill_', lambda: y.fill_(1.0)), ('zero_', lambda: y.zero_()), ('copy_', lambda: y.copy_(x))):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, 'us', round(ms * 1e3, 1), 'GB/s (bytes written)', round(y.numel() * 2 / ms / 1e6))
