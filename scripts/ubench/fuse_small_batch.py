"""getEmbedding at batch 1-8 with and without conv11 fused into conv12 (front_end.fuse_first)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from doubleattentionspeakerverification_b200 import model, synth
cfg = synth.example_config(); cfg.precision = 'bf16'
net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, 1234)).cuda().eval()
def timed(fn, iters=200):
    for _ in range(10): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for B in (1, 2, 4, 8):
    xs = torch.from_numpy(synth.make_logmel(B, 400, seed=1)).cuda()
    res = []
    for fuse in (False, True):
        net.front_end.fuse_first = fuse
        net._graphs.clear(); net._shape_hits.clear()
        with torch.no_grad():
            res.append(timed(lambda: net.getEmbedding(xs)))
    print(B, 'unfused %.1f us  fused %.1f us' % tuple(res))
