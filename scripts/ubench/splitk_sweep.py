"""Split-K factor of the conv layers at small batches: every layer alone (20 launches per graph, warm L2) with the factor the
plan picks and with DASV_CONV_SPLITK = 2, 4, 8, 16 forced."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from doubleattentionspeakerverification_b200 import ops
layers = [('conv12', 400, 80, 128, 128, True, False), ('conv21', 200, 40, 128, 256, False, False), ('conv22', 200, 40, 256, 256, True, False),
          ('conv31', 100, 20, 256, 512, False, False), ('conv32', 100, 20, 512, 512, True, False),
          ('conv41', 50, 10, 512, 1024, False, False), ('conv42', 50, 10, 1024, 1024, True, True)]
g = torch.Generator(device='cuda').manual_seed(0)


def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 100 * 1e3


for B in (1, 2, 4, 8):
    for name, T, F, Cin, Cout, pool, ref in layers:
        x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
        w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
        wp = ops.pack_conv_weight_bf16(w); bias = torch.zeros(Cout, device='cuda')
        od = torch.float32 if ref else torch.bfloat16
        fn = lambda: ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=od)
        res = []
        for sk in (None, '0', '2', '4', '8', '16'):
            os.environ.pop('DASV_CONV_SPLITK', None); os.environ.pop('DASV_CONV_NOSPLITK', None)
            if sk == '0': os.environ['DASV_CONV_NOSPLITK'] = '1'
            elif sk: os.environ['DASV_CONV_SPLITK'] = sk
            if sk and sk != '0' and int(sk) > Cin // 64:
                res.append('   -  ')
                continue
            res.append('%6.1f' % timed(fn))
        os.environ.pop('DASV_CONV_SPLITK', None); os.environ.pop('DASV_CONV_NOSPLITK', None)
        print(f'B={B} {name}: plan {res[0]}  unsplit {res[1]}  x2 {res[2]}  x4 {res[3]}  x8 {res[4]}  x16 {res[5]}  us')
