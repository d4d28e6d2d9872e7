"""One training step of the front-end + pooling on the package's kernels (for the ncu launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import CNNs, poolings
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
net = CNNs.VGG4L(1024, precision='bf16', train_kernels=True).cuda()
pool = poolings.DoubleMHA(5120, 32, mask_prob=0.3).cuda().train()
x = torch.randn(B, 400, 80, device='cuda') * 2
for _ in range(2):
    net.zero_grad(set_to_none=True); pool.zero_grad(set_to_none=True)
    out, _ = pool(net(x))
    out.square().mean().backward()
torch.cuda.synchronize()
print('ok')
