"""Patch shapes of the conv layers at B = 256 (single-CTA tiles): every valid (BF, BT, BB) with >= 112 accumulator columns
forced through DASV_CONV_PLAN against the plan the cost model picks.  One process; 10 launches per graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from doubleattentionspeakerverification_b200 import ops
B = int(os.environ.get('BATCH', 256))
layers = {'conv12': (400, 80, 128, 128, True, False), 'conv21': (200, 40, 128, 256, False, False), 'conv22': (200, 40, 256, 256, True, False),
          'conv31': (100, 20, 256, 512, False, False), 'conv32': (100, 20, 512, 512, True, False),
          'conv41': (50, 10, 512, 1024, False, False), 'conv42': (50, 10, 1024, 1024, True, True)}
g = torch.Generator(device='cuda').manual_seed(0)


def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(10): fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    gr.replay(); gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3


for name in sys.argv[1:] or list(layers):
    T, F, Cin, Cout, pool, ref = layers[name]
    x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
    wp = ops.pack_conv_weight_bf16(w); bias = torch.zeros(Cout, device='cuda')
    od = torch.float32 if ref else torch.bfloat16
    fn = lambda: ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=od)
    fl = 2.0 * B * T * F * Cout * 9 * Cin
    os.environ.pop('DASV_CONV_PLAN', None)
    us = timed(fn)
    print(f'{name}: plan of the model {us:8.1f} us {fl / us / 1e6:6.0f} TFLOP/s', flush=True)
    res = []
    for BF in range(2, F + 1, 2):
        if F % BF: continue
        for BT in range(2 if pool else 1, 256 // BF + 1, 2 if pool else 1):
            if BT > T + 1: break
            for BB in range(1, 256 // (BF * BT) + 1):
                N = (BB - 1) * (BT + 2) * BF + BT * BF
                if N > 256 or N < 112: continue
                waste = ((T + BT - 1) // BT * BT) / T
                if waste > 1.09: continue
                os.environ['DASV_CONV_PLAN'] = f'{BF},{BT},{BB}'
                try:
                    res.append((timed(fn), BF, BT, BB, N))
                except Exception as e:
                    print('   ', BF, BT, BB, 'failed', str(e)[:80])
    os.environ.pop('DASV_CONV_PLAN', None)
    for us, BF, BT, BB, N in sorted(res)[:6]:
        print(f'    BF={BF:2d} BT={BT:3d} BB={BB} N={N:3d}: {us:8.1f} us {fl / us / 1e6:6.0f} TFLOP/s')
