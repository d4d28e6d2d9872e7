#!/usr/bin/env python
"""BASELINE configs[3] on one GPU (256 utterances, U[2,20] s): the batch planner's knobs swept, and a per-batch breakdown
(batch size, padded length, valid frames, time, useful conv TFLOP/s) of the default plan."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from doubleattentionspeakerverification_b200 import extract, model, synth

dev = torch.device('cuda', 0)
cfg = synth.example_config(); cfg.precision = 'bf16'
net = synth.load_state_dict(model.SpeakerClassifier(cfg, dev), synth.make_state_dict(cfg, 1234)).to(dev).eval()
rs = np.random.RandomState(0)
frames = (100 * rs.uniform(2.0, 20.0, size=256)).astype(np.int64)
base = synth.make_logmel(1, 2000, seed=1)[0]
packed = extract.PackedUtterances([np.ascontiguousarray(np.roll(base, i * 7, axis=0)[:int(T)]) for i, T in enumerate(frames)])
useful = 12.99e9 * frames.sum() / 100.0
rec = []


def embed(x, L):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        y = net.getEmbedding(x, lengths=L)
    e1.record()
    rec.append((x.shape[0], x.shape[1], L, e0, e1))        # no host sync in here
    return y


def run(**kw):
    best = None
    for _ in range(4):
        rec.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        extract.extract_sharded(embed, packed, dev, embedding_size=400, **kw)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        if best is None or t < best[0]:
            gaps = [e0.elapsed_time(rec[0][3])] + [rec[k][4].elapsed_time(rec[k + 1][3]) for k in range(len(rec) - 1)] + [rec[-1][4].elapsed_time(e1)]
            best = (t, [(b, T, int(v.sum()), a.elapsed_time(z)) for b, T, v, a, z in rec], gaps)
    return best


t, per, gaps = run()
print(f'default plan: {t:.2f} ms, {256 / t * 1e3:.0f} emb/s, {useful / t / 1e9:.0f} useful TFLOP/s')
for b, T, v, ms in per:
    print(f'   batch {b:3d} x {T:4d} frames, valid {v:6d} ({v / (b * T):.2f}), {ms:6.2f} ms, {12.99e9 * v / 100 / ms / 1e9:6.0f} useful TFLOP/s')
print(f'   sum of batches {sum(p[3] for p in per):.2f} ms; GPU time before / between / after the batches: ' + ' '.join(f'{g:.2f}' for g in gaps) + ' ms')
for mf in (96, 128, 192, 256, 384, 512):
    for mr in (0.5, 0.7, 0.85):
        t, per, gaps = run(max_frames=mf * 400, min_ratio=mr)
        print(f'max_frames {mf:3d}x400 min_ratio {mr:.2f}: {len(per):2d} batches {t:6.2f} ms  {useful / t / 1e9:5.0f} useful TFLOP/s')
