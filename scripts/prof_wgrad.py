"""Driver for ncu: the weight-gradient kernel at two exampleModel layer sizes (conv22, conv42), batch 128."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops
B = 128
for (T, F, Cin, Cout) in ((200, 40, 256, 256), (50, 10, 1024, 1024)):
    x = torch.randn(B, T, F, Cin, device='cuda').to(torch.bfloat16)
    g = torch.randn(B, T, F, Cout, device='cuda').to(torch.bfloat16)
    for _ in range(2):
        ops.conv3x3_wgrad(x, g)
torch.cuda.synchronize()
print('ok')
