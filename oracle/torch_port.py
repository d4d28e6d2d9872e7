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


def am_softmax(x, W, label, m, s, step=0, annealing=False):
    """scripts/loss.py:37-52 (AMSoftmax.forward): cosine logits, margin m subtracted at the label, scale s."""
    xn = x / torch.norm(x, p=2, dim=1, keepdim=True).clamp(min=1e-12)
    wn = W / torch.norm(W, p=2, dim=0, keepdim=True).clamp(min=1e-12)
    costh = torch.mm(xn, wn)
    delt = torch.zeros(costh.size()).scatter_(1, label.view(-1, 1), m)
    costh_m = costh - delt
    alpha = max(0, 1000. / (pow(1. + 0.0001 * float(step), 2.))) if annealing else 0.       # loss.py:28-35
    return costh, s * ((costh_m + alpha * costh) / (1 + alpha))


def training_step(x, label, sd, cfg, keep, step=0, eps=1e-5):
    """One train.py step (scripts/train.py:215-220) for the DoubleMHA model: SpeakerClassifier.forward in train mode
    (scripts/model.py:61-71: batch-statistics b2, preLayer, AM-Softmax; head drop-out with the given keep mask,
    poolings.py:39-51) -> CrossEntropyLoss -> backward.  Returns (loss, prediction, logits, {name: gradient})."""
    names = [k for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point() and 'running' not in k
             and not k.startswith(('b1.', 'b3.'))]
    p = dict(sd)
    for k in names:
        p[k] = sd[k].clone().requires_grad_(True)
    with torch.enable_grad():
        feats = front_end(x, p, cfg.front_end)
        pooled, _ = double_mha(feats, p['poolingLayer.utteranceAttention.query'], p['poolingLayer.headsAttention.att'], keep)
        e1 = F.relu(F.linear(pooled, p['fc1.weight'], p['fc1.bias']))
        e2 = F.batch_norm(F.relu(F.linear(e1, p['fc2.weight'], p['fc2.bias'])), None, None, p['b2.weight'], p['b2.bias'],
                          training=True, eps=eps)
        e3 = F.linear(e2, p['preLayer.weight'], p['preLayer.bias'])
        pred, logits = am_softmax(e3, p['predictionLayer.W'], label, cfg.marginFactor, cfg.scalingFactor, step, cfg.annealing)
        loss = F.cross_entropy(logits, label)
        loss.backward()
    return float(loss.detach()), pred.detach(), logits.detach(), {k: p[k].grad for k in names if p[k].grad is not None}


def cosine(e1, e2):
    return F.cosine_similarity(e1, e2, dim=-1, eps=1e-8)
