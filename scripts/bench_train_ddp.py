"""Data-parallel training step (one process per GPU, torchrun): VGG4L(1024) + DoubleMHA forward + backward on the package's
kernels, gradients averaged with train_utils.allreduce_gradients (NCCL).  Weak scaling: 128 utterances per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_train_ddp.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from doubleattentionspeakerverification_b200 import CNNs, poolings, train_utils

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
B = 128
torch.manual_seed(0)
net = CNNs.VGG4L(1024, precision='bf16', train_kernels=True).cuda()
pool = poolings.DoubleMHA(5120, 32, mask_prob=0.3).cuda().train()
params = list(net.parameters()) + list(pool.parameters())
if world > 1:
    train_utils.broadcast_parameters(net); train_utils.broadcast_parameters(pool)
x = torch.randn(B, 400, 80, device='cuda', generator=torch.Generator(device='cuda').manual_seed(rank)) * 2


def step(sync=True):
    for p in params:
        p.grad = None
    out, _ = pool(net(x))
    out.square().mean().backward()
    if sync and world > 1:
        train_utils.allreduce_gradients(params)


def timed(sync, reps=5):
    for _ in range(2):
        step(sync)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step(sync)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


ms_sync, ms_local = timed(True), timed(False)
# all ranks hold the same averaged gradient afterwards
step(True)
g = torch.cat([p.grad.flatten() for p in params])
chk = torch.stack([g.sum(), g.abs().sum()])
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
else:
    same = True
if rank == 0:
    print(json.dumps({'n_gpus': world, 'batch_per_gpu': B, 'ms_per_step': round(ms_sync, 2), 'ms_without_allreduce': round(ms_local, 2),
                      'utterances_per_s': round(world * B / ms_sync * 1e3), 'gradient_bytes': int(g.numel() * 4),
                      'gradients_identical_on_all_ranks': same}))
if world > 1:
    dist.destroy_process_group()
