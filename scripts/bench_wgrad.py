"""conv3x3 weight-gradient kernel at the exampleModel layer sizes (batch 256 x 4 s): time and TFLOP/s per layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
layers = [('conv12', 400, 80, 128, 128), ('conv21', 200, 40, 128, 256), ('conv22', 200, 40, 256, 256),
          ('conv31', 100, 20, 256, 512), ('conv32', 100, 20, 512, 512), ('conv41', 50, 10, 512, 1024), ('conv42', 50, 10, 1024, 1024)]
tot_ms = tot_fl = 0.0
for name, T, F, Cin, Cout in layers:
    x = torch.randn(B, T, F, Cin, device='cuda').to(torch.bfloat16)
    g = torch.randn(B, T, F, Cout, device='cuda').to(torch.bfloat16)
    for _ in range(2):
        ops.conv3x3_wgrad(x, g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3x3_wgrad(x, g)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * T * F * Cout * 9 * Cin
    tot_ms += ms; tot_fl += fl
    print('%s  %.3f ms  %.0f TFLOP/s' % (name, ms, fl / ms / 1e9), flush=True)
    del x, g
print('total %.2f ms  %.0f TFLOP/s' % (tot_ms, tot_fl / tot_ms / 1e9))
