"""Sweep ring geometry of the DoubleMHA forward kernel on the microbench shape (tuning aid)."""
import os, subprocess, sys, json
code = r'''
import sys, os, torch
sys.path.insert(0, os.getcwd())
from doubleattentionspeakerverification_b200 import ops
B, T, D, H = 512, 200, 1024, 16
g = torch.Generator(device='cuda').manual_seed(0)
q = torch.randn(D // H, H, device='cuda', generator=g) * 0.3
a = torch.randn(D // H, device='cuda', generator=g) * 0.3
out = {}
for name, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
    xs = [torch.randn(B, T, D, device='cuda', generator=g).to(dt) for _ in range(2)]
    for i in range(3): ops.dmha_fwd(xs[i & 1], q, a, need_align=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(20): ops.dmha_fwd(xs[i & 1], q, a, need_align=False)
    e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) * 1e3 / 20, 1)
    del xs
print(out)
'''
for fps in (0, 8, 16, 32):
    for st in (0, 3, 4, 5, 6, 8):
        env = dict(os.environ)
        if fps: env['DASV_DMHA_FPS'] = str(fps)
        if st: env['DASV_DMHA_STAGES'] = str(st)
        r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True)
        print('fps', fps, 'stages', st, (r.stdout.strip() or r.stderr.strip()[-200:]), flush=True)
