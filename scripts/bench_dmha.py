"""DoubleMHA pooling microbench (BASELINE configs[1]: B=512, T=200, D=1024, H=16), graph-timed, forward only.
usage: python scripts/bench_dmha.py [bf16|fp32|both]   (kernel knobs come from the DASV_DMHA* environment variables)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import ops, synth

which = sys.argv[1] if len(sys.argv) > 1 else 'both'
B, T, D, H = int(os.environ.get('BENCH_B', 512)), 200, 1024, 16     # BENCH_B: other utterance counts (tail / quantisation studies)
PEAK = 6538.3
dev = 'cuda'
gen = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(D // H, H, device=dev, generator=gen) * 0.3
a = torch.randn(D // H, device=dev, generator=gen) * 0.3
lens = torch.from_numpy(synth.make_lengths(B, 100, 200, seed=0)).to(dev)


def time_us(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


out = {}
for name, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
    if which not in ('both', name):
        continue
    xs = [torch.randn(B, T, D, device=dev, generator=gen).to(dt) for _ in range(2)]
    es = 4 if dt == torch.float32 else 2
    lens_sorted = torch.sort(lens, descending=True).values       # the order extract.bucket_plan hands batches over in
    for case, L in (('full', None), ('masked', lens), ('masked_sorted', lens_sorted)):
        nbytes = (B * T if L is None else int(L.sum().item())) * D * es
        us = time_us(lambda i: ops.dmha_fwd(xs[i & 1], q, a, lengths=L, need_align=False))
        out['%s_%s' % (name, case)] = (round(us, 1), round(nbytes / us / 1e3 / PEAK, 3))
    del xs
print({k: os.environ[k] for k in os.environ if k.startswith(('DASV_', 'BENCH_'))}, out, flush=True)
