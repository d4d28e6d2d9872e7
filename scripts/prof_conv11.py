import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops
x = torch.randn(256, 400, 80, device='cuda')
w = torch.randn(128, 1, 3, 3, device='cuda') * 0.3
b = torch.randn(128, device='cuda') * 0.1
for name, fn in (('direct', lambda: ops.conv11_direct(x, w, b, out_dtype=torch.bfloat16)), ('tc', lambda: ops.conv11_tc(x, w, b))):
    for i in range(4):
        y = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        y = fn()
    e1.record(); torch.cuda.synchronize()
    print('conv11', name, 'us', round(e0.elapsed_time(e1) * 100, 1), 'GB/s written', round(y.numel() * 2 / (e0.elapsed_time(e1) * 1e-4) / 1e9))
