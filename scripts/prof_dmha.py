"""Small driver for ncu: the DoubleMHA pooling microbench shape (BASELINE configs[1]), a few launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops

dt = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == 'bf16' else torch.float32
B, T, D, H = 512, 200, 1024, 16
g = torch.Generator(device='cuda').manual_seed(0)
xs = [torch.randn(B, T, D, device='cuda', generator=g).to(dt) for _ in range(2)]
q = torch.randn(D // H, H, device='cuda', generator=g) * 0.3
a = torch.randn(D // H, device='cuda', generator=g) * 0.3
for i in range(6):
    r = ops.dmha_fwd(xs[i & 1], q, a, need_align=False)
gout = torch.randn(B, D // H, device='cuda', generator=g)
for i in range(4):
    ops.dmha_bwd(xs[i & 1], q, a, gout, None, r['ctx'], r['lse'], r['headw'])
torch.cuda.synchronize()
print('ok')
