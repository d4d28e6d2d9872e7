"""Where the time of one conv3x3_igemm launch goes at small batch: per-CTA %globaltimer stamps (dasv_debug_conv_trace) of the
exampleModel layers at BATCH (default 1), launched back to back on a warm L2 like the steps of a stream.  Times in us from
the first CTA's start; min / median / max over the CTAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doubleattentionspeakerverification_b200 import _lib, ops
B = int(os.environ.get('BATCH', 1))
layers = [('conv12', 400, 80, 128, 128, True, False), ('conv21', 200, 40, 128, 256, False, False), ('conv22', 200, 40, 256, 256, True, False),
          ('conv31', 100, 20, 256, 512, False, False), ('conv32', 100, 20, 512, 512, True, False),
          ('conv41', 50, 10, 512, 1024, False, False), ('conv42', 50, 10, 1024, 1024, True, True)]
names = ['CTA start', 'set-up done', 'prev. kernel done', 'first operands', 'last MMA issued', 'accumulator done', 'epilogue done', 'CTA end']
g = torch.Generator(device='cuda').manual_seed(0)
buf = torch.zeros(8 * 1024, dtype=torch.int64, device='cuda')
L = _lib.lib()
for name, T, F, Cin, Cout, pool, ref in layers:
    x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
    wp = ops.pack_conv_weight_bf16(w); bias = torch.zeros(Cout, device='cuda')
    od = torch.float32 if ref else torch.bfloat16
    run = lambda: ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=od)
    for _ in range(5): run()
    torch.cuda.synchronize()
    buf.zero_()
    L.dasv_debug_conv_trace(buf.data_ptr())
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(6): run()                    # the last launch's stamps stay in the buffer
    gr.replay(); gr.replay()
    torch.cuda.synchronize()
    L.dasv_debug_conv_trace(None)
    t = buf.cpu().numpy().reshape(-1, 8)
    t = t[t[:, 0] > 0].astype(np.float64)
    t0 = t[:, 0].min()
    print(f'{name}: {len(t)} CTAs')
    for i, n in enumerate(names):
        v = (t[:, i] - t0) / 1e3
        print(f'   {n:18s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}')
