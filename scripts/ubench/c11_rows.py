"""conv11 frames per CTA (DASV_C11_ROWS) at batch 1-8: one layer, 20 launches per graph."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from doubleattentionspeakerverification_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
for B in (1, 2, 4, 8):
    x0 = torch.randn(B, 400, 80, device='cuda', generator=g)
    w0 = torch.randn(128, 1, 3, 3, device='cuda', generator=g); b0 = torch.zeros(128, device='cuda')
    fn = lambda: ops.conv11_direct(x0, w0, b0, out_dtype=torch.bfloat16)
    for _ in range(5): fn()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(20): fn()
    gr.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print('R', os.environ.get('DASV_C11_ROWS'), 'B', B, round(e0.elapsed_time(e1) / 200 * 1e3, 2), 'us')
