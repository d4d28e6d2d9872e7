"""Per-kernel time of the batch-1 step (one 4 s utterance): every layer launched back to back ITERS times on a warm
cache, against its FLOP floor at the sustained tensor rate; then the whole step eager and replayed from its graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import model, ops, synth
B = int(os.environ.get('BATCH', 1)); ITERS = int(os.environ.get('ITERS', 200))
layers = [('conv12', 400, 80, 128, 128, True, False), ('conv21', 200, 40, 128, 256, False, False), ('conv22', 200, 40, 256, 256, True, False),
          ('conv31', 100, 20, 256, 512, False, False), ('conv32', 100, 20, 512, 512, True, False),
          ('conv41', 50, 10, 512, 1024, False, False), ('conv42', 50, 10, 1024, 1024, True, True)]
g = torch.Generator(device='cuda').manual_seed(0)

def timed(fn, iters=ITERS, graph=True):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if graph:                                       # 20 launches per graph: the Python call (~20 us) would otherwise bound the loop
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20): fn()
        run, per = gr.replay, 20
        iters = max(1, iters // 20)
    else:
        run, per = fn, 1
    run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (iters * per) * 1e3        # us

total = 0.0
# calibration: the fixed cost of a launch in these loops (a one-CTA conv11, a one-tile tensor-core conv)
xt = torch.randn(1, 2, 2, device='cuda', generator=g)
wt = torch.randn(128, 1, 3, 3, device='cuda', generator=g); bt = torch.zeros(128, device='cuda')
print(f'one-CTA conv11 launch      {timed(lambda: ops.conv11_direct(xt, wt, bt, out_dtype=torch.bfloat16)):7.2f} us')
xq = torch.randn(1, 2, 4, 64, device='cuda', generator=g).to(torch.bfloat16)
wq = ops.pack_conv_weight_bf16(torch.randn(128, 64, 3, 3, device='cuda', generator=g))
print(f'one-tile igemm conv launch {timed(lambda: ops.conv3x3_igemm_bf16(xq, wq, bt, 128)):7.2f} us   (K = 576: 36 MMAs)')
x0 = torch.randn(B, 400, 80, device='cuda', generator=g)
w0 = torch.randn(128, 1, 3, 3, device='cuda', generator=g); b0 = torch.zeros(128, device='cuda')
us = timed(lambda: ops.conv11_direct(x0, w0, b0, out_dtype=torch.bfloat16)); total += us
print(f'conv11  {us:7.2f} us')
for name, T, F, Cin, Cout, pool, ref in layers:
    x = torch.randn(B, T, F, Cin, device='cuda', generator=g).relu_().to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device='cuda', generator=g) * (2.0 / (9 * Cin)) ** 0.5
    wp = ops.pack_conv_weight_bf16(w); bias = torch.zeros(Cout, device='cuda')
    od = torch.float32 if ref else torch.bfloat16
    pair = pool and Cin >= 256 and Cout % 256 == 0            # what CNNs.py asks for (split-K launches ignore it)
    us = timed(lambda: ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=od, pair=pair)); total += us
    if pair:
        us0 = timed(lambda: ops.conv3x3_igemm_bf16(x, wp, bias, Cout, pool=pool, ref_layout=ref, out_dtype=od, pair=False))
        print(f'        (without CTA pairs {us0:7.2f} us)')
    fl = 2.0 * B * T * F * Cout * 9 * Cin
    print(f'{name}  {us:7.2f} us   floor {fl / 1.4e15 * 1e6:5.2f} us   weights {wp.numel() * 2 / 1e6:5.1f} MB')
print(f'sum of conv kernels {total:.1f} us')
# pooling (B x 50 x 10240 reference-layout fp32... the step feeds it what conv42 wrote) and the FC/BN tail, alone
cfg = synth.example_config(); cfg.precision = 'bf16'
net0 = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, 1234)).cuda().eval()
enc = torch.randn(B, 25, 5120, device='cuda', generator=g)
pl = net0.poolingLayer
with torch.no_grad():
    for dt in (torch.float32, torch.bfloat16):
        e = enc.to(dt)
        us = timed(lambda: pl.pooled(e))
        print(f'pooling {str(dt)[6:]:9s} {us:7.2f} us   ({e.numel() * e.element_size() / 1e6:.2f} MB)')
    pooled = pl.pooled(enc)
    tp = net0._tail_params()
    print(f'fc tail           {timed(lambda: ops.fc_tail(pooled, *tp)):7.2f} us')
    xin = torch.randn(B, 400, 80, device='cuda'); xst = torch.empty_like(xin)
    print(f'input copy        {timed(lambda: xst.copy_(xin)):7.2f} us')
del net0

net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, 1234)).cuda().eval()
xs = torch.from_numpy(synth.make_logmel(B, 400, seed=1)).cuda()
with torch.no_grad():
    net.use_graphs = False
    print(f'step eager   {timed(lambda: net.getEmbedding(xs), 100, graph=False):7.1f} us')
    net.use_graphs = True
    print(f'step graph   {timed(lambda: net.getEmbedding(xs), 100, graph=False):7.1f} us')
