#!/usr/bin/env python
"""Where the e2e step (pinned host in -> pinned host out) loses time against the resident step: variants of bench.py's
HostPipeline loop with one part removed at a time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from doubleattentionspeakerverification_b200 import extract, model, synth

dev = torch.device('cuda', 0)
cfg = synth.example_config(); cfg.precision = 'bf16'
net = synth.load_state_dict(model.SpeakerClassifier(cfg, dev), synth.make_state_dict(cfg, 1234)).to(dev).eval()
B = 256
x_host = torch.from_numpy(synth.make_logmel(B, 400, seed=3)).pin_memory()
x_dev = [x_host.to(dev), x_host.to(dev) * 1.01]
emb_host = torch.empty((B, 400)).pin_memory()


def timed(fn, n=20):
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def embed(xd):
    with torch.no_grad():
        return net.getEmbedding(xd)


for graphs in (True, False):
    net.use_graphs = graphs
    print('graphs', graphs)
    print('  resident                     : %.3f ms' % timed(lambda i: embed(x_dev[i & 1])))
    pipe = extract.HostPipeline(embed, (B, 400, 80), 400, dev)
    print('  H2D + step + D2H (bench e2e) : %.3f ms' % timed(lambda i: pipe.submit(x_host, emb_host)))
    print('  step + D2H (no H2D)          : %.3f ms' % timed(lambda i: emb_host.copy_(embed(x_dev[i & 1]), non_blocking=True)))
    cs = torch.cuda.Stream()
    stage = torch.empty((B, 400, 80), device=dev)

    def h2d_only(i):
        with torch.cuda.stream(cs):
            stage.copy_(x_host, non_blocking=True)
        embed(x_dev[i & 1])
    print('  step with a concurrent, unrelated H2D on another stream: %.3f ms' % timed(h2d_only))
cs.synchronize()
print('  H2D alone: %.3f ms' % timed(lambda i: stage.copy_(x_host, non_blocking=True)))
