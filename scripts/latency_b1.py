#!/usr/bin/env python
"""One 4 s utterance at batch 1 (BASELINE configs[0] on the GPU): a few eager getEmbedding calls, for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from doubleattentionspeakerverification_b200 import model, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = synth.example_config(); cfg.precision = 'bf16'; cfg.use_graphs = False
net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, 1234)).cuda().eval()
x = torch.from_numpy(synth.make_logmel(B, 400, seed=1)).cuda()
with torch.no_grad():
    for _ in range(4):
        e = net.getEmbedding(x)
torch.cuda.synchronize()
print('ok', tuple(e.shape))
