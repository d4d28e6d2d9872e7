import os, sys, ast
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from argparse import Namespace
from conftest import golden, min_cosine
from doubleattentionspeakerverification_b200 import model, ops, synth
g = golden('embed_k512.npz')
cfg = Namespace(**ast.literal_eval(str(g['cfg']))); cfg.precision = 'bf16'
B, T, seed = [int(v) for v in g['spec']]
net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, seed)).cuda().eval()
x = torch.from_numpy(synth.make_logmel(B, T, seed)).cuda()
tc = ops.conv11_tc
def direct(x, w, b, L=None): return ops.conv11_direct(x, w, b, L, out_dtype=torch.bfloat16)
for name, fn in (('tc', tc), ('direct', direct), ('tc', tc)):
    ops.conv11_tc = fn
    with torch.no_grad():
        for rep in range(2):
            e = net.getEmbedding(x)
            print(name, rep, 'cos', min_cosine(e.cpu().numpy(), g['emb']), flush=True)
for env in ('0', '1'):
    os.environ['DASV_CONV_REUSE'] = env
    ops.conv11_tc = direct
    with torch.no_grad():
        e = net.getEmbedding(x)
    print('reuse', env, 'cos', min_cosine(e.cpu().numpy(), g['emb']), flush=True)
