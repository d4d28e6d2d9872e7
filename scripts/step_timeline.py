#!/usr/bin/env python
"""Where does a getEmbedding step spend its time?  Per-slot device times (an event after every C call of one step), the
eager step time, and the same step replayed from a CUDA graph.  Diagnostic for the gap between the sum of the kernel
durations and the step time.

    python scripts/step_timeline.py [--batch 256] [--steps 20]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from doubleattentionspeakerverification_b200 import model, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--frames', type=int, default=400)
    ap.add_argument('--steps', type=int, default=20)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    cfg = synth.example_config()
    cfg.precision = 'bf16'
    net = synth.load_state_dict(model.SpeakerClassifier(cfg, dev), synth.make_state_dict(cfg, 1234)).to(dev).eval()
    net.use_graphs = False            # this script drives the graph itself
    xs = [torch.from_numpy(synth.make_logmel(args.batch, args.frames, seed=100 + i)).to(dev) for i in range(2)]

    def step(i):
        with torch.no_grad():
            return net.getEmbedding(xs[i & 1])

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for i in range(5):
        step(i)
    print('eager            : %.3f ms/step' % timed(step, args.steps))

    # ---- per-slot times: an event after every op of the step
    names = ['conv11_direct', 'conv12_fused', 'conv3x3_igemm_bf16', 'dmha_fwd', 'fc_tail']
    orig = {n: getattr(ops, n) for n in names}
    marks = []

    def wrap(n):
        def f(*a, **k):
            y = orig[n](*a, **k)
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((n, e))
            return y
        return f

    for n in names:
        setattr(ops, n, wrap(n))
    nrep = 10
    starts = []
    torch.cuda.synchronize()
    for i in range(nrep):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        starts.append((len(marks), e))
        step(i)
    torch.cuda.synchronize()
    for n in names:
        setattr(ops, n, orig[n])
    per = len(marks) // nrep
    slot = [0.0] * per
    for r, (m0, e0) in enumerate(starts):
        prev = e0
        for j in range(per):
            n, e = marks[m0 + j]
            slot[j] += prev.elapsed_time(e) / nrep
            prev = e
    tot = 0.0
    for j in range(per):
        print('  slot %2d %-22s %.4f ms' % (j, marks[j][0], slot[j]))
        tot += slot[j]
    print('sum of slots     : %.3f ms (events between every launch)' % tot)

    # ---- the same step from a CUDA graph
    static_x = xs[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(3):
            with torch.no_grad():
                net.getEmbedding(static_x)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        with torch.no_grad():
            out = net.getEmbedding(static_x)
    g.replay()
    torch.cuda.synchronize()
    ref = step(0)
    static_x.copy_(xs[0])
    g.replay()
    torch.cuda.synchronize()
    print('graph == eager   :', bool(torch.equal(out, ref)))
    print('cuda graph       : %.3f ms/step' % timed(lambda i: g.replay(), args.steps))
    print('eager again      : %.3f ms/step' % timed(step, args.steps))
    net.use_graphs = True
    for i in range(4):
        step(i)
    print('module graphs    : %.3f ms/step (getEmbedding with use_graphs, incl. the input copy and output clone)' % timed(step, args.steps))
    net.use_graphs = False
    for b in (1, 8):
        xb = xs[0][:b].contiguous()
        f = lambda i: net.getEmbedding(xb)
        with torch.no_grad():
            for i in range(4):
                f(i)
            te = timed(f, 50)
            net.use_graphs = True
            for i in range(4):
                f(i)
            tg = timed(f, 50)
            net.use_graphs = False
        print('batch %d latency  : eager %.3f ms, graph %.3f ms' % (b, te, tg))

    # ---- conv layers alone, back to back
    fe = net.front_end
    h0 = ops.conv11_direct(xs[0], fe.conv11.weight, fe.conv11.bias, None, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    print('conv11 alone     : %.3f ms' % timed(lambda i: ops.conv11_direct(xs[i & 1], fe.conv11.weight, fe.conv11.bias, None, out_dtype=torch.bfloat16), 20))
    wp = fe._pack('conv12', 'bf16')
    print('conv12 alone     : %.3f ms' % timed(lambda i: ops.conv3x3_igemm_bf16(h0, wp, fe.conv12.bias, 128, None, pool=True), 20))


if __name__ == '__main__':
    main()
