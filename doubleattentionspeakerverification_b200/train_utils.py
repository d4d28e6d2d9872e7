"""Data-parallel training plumbing: one process per GPU instead of the reference's ``nn.DataParallel``
(scripts/train.py:68-70), which scatters the batch from one Python process and gathers the outputs on GPU 0.

Utterances are independent in the loss (train.py:195-203: mean cross-entropy over the batch), so every rank runs the
model on its own slice of the batch and the parameter gradients are averaged with ONE all-reduce per flat bucket
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  BatchNorm1d statistics (b1/b2/b3, scripts/model.py:45-50)
stay per rank, as they do per replica under DataParallel.
"""
import torch
import torch.distributed as dist


def shard_batch(n, rank=None, world=None):
    """The contiguous slice of a batch of ``n`` items this rank owns (sizes differ by at most one)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def allreduce_gradients(params, group=None, bucket_bytes=64 << 20, local_weight=None):
    """Average ``p.grad`` over the ranks of ``group``, in place.  Gradients are packed into flat buckets of about
    ``bucket_bytes`` (few large collectives instead of one per tensor).  ``local_weight`` = this rank's share of the
    global batch (default 1/world): with uneven shards the weighted sum equals the gradient of the global-batch mean
    loss.  Parameters without a gradient on this rank contribute zeros (every rank must pass the same list)."""
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    if world == 1 or not params:
        return
    w = (1.0 / world) if local_weight is None else float(local_weight)
    bucket, size = [], 0

    def flush():
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket]) * w
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        o = 0
        for p in bucket:
            n = p.numel()
            g = flat[o:o + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            o += n

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
