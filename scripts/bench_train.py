"""Forward + backward of the VGG4L front-end + DoubleMHA pooling (exampleModel sizes) on this package's kernels vs the
torch/cuDNN autograd path: ms per step and the conv FLOP rate (3 x forward FLOPs minus conv11's input gradient).
usage: python scripts/bench_train.py [batch]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from doubleattentionspeakerverification_b200 import CNNs, poolings

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = 400
torch.manual_seed(0)
net = CNNs.VGG4L(1024, precision='bf16', train_kernels=True).cuda()
pool = poolings.DoubleMHA(5120, 32, mask_prob=0.3).cuda().train()
x = torch.randn(B, T, 80, device='cuda') * 2
C = [128, 128, 256, 256, 512, 512, 1024, 1024]
dims = [(400, 80), (400, 80), (200, 40), (200, 40), (100, 20), (100, 20), (50, 10), (50, 10)]
fwd = sum(2.0 * B * t * f * co * 9 * ci for (t, f), ci, co in zip(dims, [1] + C[:-1], C))
flops = 3.0 * fwd


def step():
    net.zero_grad(set_to_none=True); pool.zero_grad(set_to_none=True)
    out, _ = pool(net(x))
    out.square().mean().backward()


def time_ms(reps):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {'batch': B}
ms = time_ms(5)
res['kernels_bf16'] = {'ms': round(ms, 2), 'utt_per_s': round(B / ms * 1e3), 'conv_tflops': round(flops / ms / 1e9)}
net.train_kernels = False
torch.backends.cudnn.allow_tf32 = False
ms = time_ms(3)
res['torch_cudnn_fp32'] = {'ms': round(ms, 2), 'utt_per_s': round(B / ms * 1e3), 'conv_tflops': round(flops / ms / 1e9)}
torch.backends.cudnn.allow_tf32 = True
ms = time_ms(3)
res['torch_cudnn_tf32'] = {'ms': round(ms, 2), 'utt_per_s': round(B / ms * 1e3), 'conv_tflops': round(flops / ms / 1e9)}
with torch.autocast('cuda', dtype=torch.bfloat16):
    ms = time_ms(3)
res['torch_cudnn_autocast_bf16'] = {'ms': round(ms, 2), 'utt_per_s': round(B / ms * 1e3), 'conv_tflops': round(flops / ms / 1e9)}
print(json.dumps(res))
