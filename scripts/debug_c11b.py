import os, sys, ast
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
from argparse import Namespace
from conftest import golden
from doubleattentionspeakerverification_b200 import ops, synth
g = golden('embed_k512.npz')
cfg = Namespace(**ast.literal_eval(str(g['cfg'])))
B, T, seed = [int(v) for v in g['spec']]
sd = synth.make_state_dict(cfg, seed)
x = torch.from_numpy(synth.make_logmel(B, T, seed)).cuda()
w = torch.from_numpy(sd['front_end.conv11.weight']).cuda(); b = torch.from_numpy(sd['front_end.conv11.bias']).cuda()
ref = torch.relu(torch.nn.functional.conv2d(x.double().unsqueeze(1), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)  # NHWC fp64
for name, y in (('tc', ops.conv11_tc(x, w, b).double()), ('direct', ops.conv11_direct(x, w, b, out_dtype=torch.bfloat16).double()),
                ('direct_f32', ops.conv11_direct(x, w, b).double())):
    err = y - ref
    m = ref > 0.05
    rel = (err[m] / ref[m])
    print(name, 'max abs %.4g' % float(err.abs().max()), 'rms rel %.3e' % float(rel.pow(2).mean().sqrt()), 'mean rel %.3e' % float(rel.mean()),
          'frac |rel|>2^-8: %.4f' % float((rel.abs() > 2 ** -8).double().mean()), flush=True)
    big = (err.abs() > 0.05).nonzero()
    print('   n big', len(big), big[:5].tolist())
