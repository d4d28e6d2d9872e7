"""Time the igemm conv on the exampleModel layer shapes (B=256, T=400 input), optionally forcing pool off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops
B = int(os.environ.get('BATCH', 256))
layers = [('conv12', 400, 80, 128, 128, True, False), ('conv21', 200, 40, 128, 256, False, False), ('conv22', 200, 40, 256, 256, True, False),
          ('conv31', 100, 20, 256, 512, False, False), ('conv32', 100, 20, 512, 512, True, False),
          ('conv41', 50, 10, 512, 1024, False, False), ('conv42', 50, 10, 1024, 1024, True, True)]
force_nopool = os.environ.get('NOPOOL') == '1'
g = torch.Generator(device='cuda').manual_seed(0)
res = {}
for name, T, F, Cin, Cout, pool, ref in layers:
    if force_nopool: pool, ref = False, False
    x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
    wp = ops.pack_conv_weight_bf16(w); bias = torch.zeros(Cout, device='cuda')
    for _ in range(2): ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=torch.float32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=torch.float32)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    res[name] = round(2.0 * B * T * F * Cout * 9 * Cin / ms / 1e9, 0)
    del x, w, wp
print({k: int(v) for k, v in res.items()})
