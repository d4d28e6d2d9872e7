"""Latency / throughput of getEmbedding (exampleModel, bf16, 4 s utterances) versus batch size, inputs resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import model, synth
cfg = synth.example_config(); cfg.precision = 'bf16'
net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, 1234)).cuda().eval()
out = {}
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
    x = torch.from_numpy(synth.make_logmel(B, 400, seed=B)).cuda()
    with torch.no_grad():
        for _ in range(3): net.getEmbedding(x)
        reps = max(5, min(50, 2048 // B))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps): net.getEmbedding(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[B] = (round(ms, 3), round(B / ms * 1e3))
print({b: {'ms': v[0], 'emb_per_s': v[1]} for b, v in out.items()})
