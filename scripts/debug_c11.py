import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from doubleattentionspeakerverification_b200 import ops
for (B, T, F, Cout) in [(2, 64, 80, 64), (1, 64, 80, 64), (2, 64, 80, 128), (2, 63, 80, 64), (1, 16, 80, 64), (4, 64, 80, 8)]:
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(B, T, F, device='cuda', generator=g) * 2
    w = torch.randn(Cout, 1, 3, 3, device='cuda', generator=g) * 0.5
    b = torch.randn(Cout, device='cuda', generator=g) * 0.1
    for rep in range(3):
        y = ops.conv11_tc(x, w, b).float()
        yd = ops.conv11_direct(x, w, b, out_dtype=torch.bfloat16).float()
        d = (y - yd).abs()
        bad = (d > 0.05 * yd.abs().max()).nonzero()
        print((B, T, F, Cout), rep, 'max diff %.4f' % float(d.max()), 'nbad', len(bad), bad[:4].tolist(), bad[-2:].tolist(), flush=True)
