"""Pair-mode (cta_group::2) conv vs the single-CTA kernel on the same inputs (correctness + plan print)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from doubleattentionspeakerverification_b200 import ops
g = torch.Generator(device='cuda').manual_seed(3)
for (B, T, F, Cin, Cout, pool, ref, with_len) in [(3, 13, 20, 128, 256, False, False, True), (3, 12, 20, 256, 256, True, False, False),
                                                  (4, 50, 10, 64, 256, True, True, False), (5, 7, 10, 128, 512, True, True, True),
                                                  (8, 200, 40, 128, 256, False, False, False), (8, 100, 20, 512, 512, True, False, True),
                                                  (16, 50, 10, 1024, 1024, True, True, False)]:
    x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
    bias = torch.randn(Cout, device='cuda', generator=g) * 0.1
    L = None
    if with_len:
        L = torch.randint(1, T + 1, (B,), device='cuda', generator=g, dtype=torch.int32); L[0] = T
        x = x * (torch.arange(T, device='cuda')[None, :, None, None] < L[:, None, None, None])
    wp = ops.pack_conv_weight_bf16(w)
    outs = {}
    for pair in ('0', '1'):
        os.environ['DASV_CONV_PAIR'] = pair
        outs[pair] = ops.conv3x3_igemm_bf16(x, wp, bias, Cout, lengths=L, pool=pool, ref_layout=ref, out_dtype=torch.float32).float()
    torch.cuda.synchronize()
    d = (outs['0'] - outs['1']).abs().max().item()
    print((B, T, F, Cin, Cout, pool, ref, with_len), 'max |pair - single| = %.4g' % d, 'scale %.3g' % outs['0'].abs().max().item(), flush=True)
