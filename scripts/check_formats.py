#!/usr/bin/env python
"""Which operand-format combinations does tcgen05.mma kind::f16 accept?  One combination per process (a fault is sticky)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from doubleattentionspeakerverification_b200 import ops
xdt = {'bf16': torch.bfloat16, 'f16': torch.float16}[sys.argv[1]]
wdt = {'bf16': torch.bfloat16, 'f16': torch.float16}[sys.argv[2]]
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.relu(torch.randn(2, 12, 20, 128, device='cuda', generator=g)).to(xdt)
w = (torch.randn(128, 128, 3, 3, device='cuda', generator=g) * 0.03).to(wdt).float()
b = torch.zeros(128, device='cuda')
y = ops.conv3x3_igemm_bf16(x, ops.pack_conv_weight_bf16(w, wdt), b, 128)
torch.cuda.synchronize()
ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1)).permute(0, 2, 3, 1)
print(sys.argv[1:], 'max rel err', float((y.float() - ref).abs().max() / ref.abs().max()))
