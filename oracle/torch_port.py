"""CPU port of the reference path on the reference's own arithmetic library (torch CPU).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference is pure
Python over PyTorch and ``/root/reference`` does not travel to the GPU box, so
the timed CPU baseline (``bench.py`` ``cpu_baseline`` / ``--impl reference``,
``kind: "port"``) is this functional restatement: the same torch library calls,
in the same order, on the same layouts as ``scripts/CNNs.py:68-91``,
``scripts/poolings.py:73-80,45-51,126-129`` and ``scripts/model.py:52-59`` —
including the reference's redundant H×H score matmul + diagonal, which is what
the reference's CPU path actually pays for.  Pinned against the live reference
by ``tests/golden`` (``tests/test_oracle_golden.py``).
"""
import math

import torch
import torch.nn.functional as F

from doubleattentionspeakerverification_b200.synth import conv_names


def as_torch(sd):
    return {k: torch.from_numpy(v.copy()) if hasattr(v, 'dtype') and not torch.is_tensor(v) else v
            for k, v in sd.items()}


def front_end(x, sd, front='VGG4L'):
    # [B,T,80] -> NCHW [B,1,T,80]; (conv,relu,conv,relu,pool)* ; [B,C,T',F'] -> [B,T',C*F']
    h = x.unsqueeze(1)
    names = conv_names(front)
    for i in range(0, len(names), 2):
        for n in names[i:i + 2]:
            h = F.relu(F.conv2d(h, sd['front_end.%s.weight' % n], sd['front_end.%s.bias' % n], stride=1, padding=1))
        h = F.max_pool2d(h, 2, stride=2, ceil_mode=True)
    h = h.transpose(1, 2).contiguous()
    return h.view(h.size(0), h.size(1), -1)


def mha(x, query):
    B, T, _ = x.shape
    dh, H = query.shape
    key = x.view(B * T, H, dh)
    value = x.view(B, T, H, dh)
    full = torch.matmul(key, query) / math.sqrt(H)          # [B*T,H,H]: the reference computes all of it
    scores = torch.diagonal(full, dim1=-2, dim2=-1).view(B, T, H)
    p = F.softmax(scores, dim=-2)
    ctx = torch.sum(value * p.unsqueeze(-1), dim=1)
    return ctx, p


def heads(ctx, att, keep=None):
    u = torch.matmul(ctx, att).squeeze(-1)
    if keep is not None:
        u = u.masked_fill(~keep, -float('inf'))
    w = F.softmax(u, dim=-1).unsqueeze(-1)
    return torch.sum(ctx * w, dim=1), w


def double_mha(x, query, att, keep=None):
    ctx, p = mha(x, query)
    out, w = heads(ctx, att, keep)
    return out, p


def tail(pooled, sd, eps=1e-5):
    e1 = F.relu(F.linear(pooled, sd['fc1.weight'], sd['fc1.bias']))
    e2 = F.relu(F.linear(e1, sd['fc2.weight'], sd['fc2.bias']))
    return F.batch_norm(e2, sd['b2.running_mean'], sd['b2.running_var'], sd['b2.weight'], sd['b2.bias'],
                        training=False, eps=eps)


@torch.no_grad()
def get_embedding(x, sd, cfg):
    """eval-mode SpeakerClassifier.getEmbedding for the DoubleMHA / MHA / Attention poolings."""
    feats = front_end(x, sd, cfg.front_end)
    if cfg.pooling_method == 'DoubleMHA':
        pooled, _ = double_mha(feats, sd['poolingLayer.utteranceAttention.query'], sd['poolingLayer.headsAttention.att'])
    elif cfg.pooling_method == 'MHA':
        ctx, _ = mha(feats, sd['poolingLayer.query'])
        pooled = ctx.reshape(ctx.size(0), -1)
    else:
        a = sd['poolingLayer.att']
        p = F.softmax(torch.matmul(feats, a).squeeze(-1), dim=-1).unsqueeze(-1)
        pooled = torch.sum(feats * p, dim=1)
    return tail(pooled, sd)


def cosine(e1, e2):
    return F.cosine_similarity(e1, e2, dim=-1, eps=1e-8)
