"""Log-mel feature extraction on the GPU: throughput at BASELINE configs[2]'s shape (256 utterances x 4 s, 16 kHz).
(The accuracy against the oracle is measured in tests/test_gpu_features.py.)  usage: python scripts/bench_features.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from doubleattentionspeakerverification_b200 import featureExtractor as fe, synth

B, sfr = 256, 16000
n = 512 + 160 * 399                                 # exactly 400 frames
wave = np.stack([synth.make_waveform(n, sfr, seed=i) for i in range(8)] * (B // 8)).astype(np.float32)
wd = torch.from_numpy(wave).cuda()
ns = [n] * B
for _ in range(3):
    feat, frames = fe.logmel_batch(wd, ns, sfr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    feat, frames = fe.logmel_batch(wd, ns, sfr)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = B * n * 4 + B * 400 * 80 * 4 * 3           # wave read + features written, read and rewritten by the CMN pass
print(json.dumps({'utterances': B, 'frames_per_utt': int(frames[0]), 'ms_per_batch_incl_python': round(ms, 3),
                  'utt_per_s': round(B / ms * 1e3), 'algorithmic_GBps': round(nbytes / ms / 1e6, 1)}))
