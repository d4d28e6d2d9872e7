"""numpy restatement of the reference's embedding-extraction path (the checker).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines it restates.  The arithmetic is written out explicitly (no torch),
so it is an independent check of both the reference's library calls and the CUDA
kernels.  Pinned against the live reference by ``oracle/make_golden.py`` →
``tests/golden/*.npz`` → ``tests/test_oracle_golden.py``.

Layouts: activations are NHWC ``[B, T, F, C]`` inside the front-end; everything
visible at a module boundary uses the reference's layout.
"""
import math

import numpy as np


# --------------------------------------------------------------------------- front-end
def conv3x3_same(x, w, b):
    """3x3 stride-1 pad-1 convolution with bias (torch.nn.Conv2d(…,3,stride=1,padding=1),
    scripts/CNNs.py:59-66).  x ``[B,T,F,Cin]`` NHWC, w ``[Cout,Cin,3,3]`` (reference OIHW), b ``[Cout]``."""
    B, T, F, Cin = x.shape
    Cout = w.shape[0]
    xp = np.zeros((B, T + 2, F + 2, Cin), dtype=x.dtype)
    xp[:, 1:T + 1, 1:F + 1, :] = x
    y = np.zeros((B, T, F, Cout), dtype=x.dtype)
    for dy in range(3):
        for dx in range(3):
            # cross-correlation: out[t,f] += x[t+dy-1, f+dx-1] * w[:,:,dy,dx]
            tap = xp[:, dy:dy + T, dx:dx + F, :].reshape(-1, Cin)
            y += (tap @ w[:, :, dy, dx].T.astype(x.dtype)).reshape(B, T, F, Cout)
    return y + b.astype(x.dtype)


def relu(x):
    return np.maximum(x, 0)


def maxpool2x2_ceil(x):
    """F.max_pool2d(x, 2, stride=2, ceil_mode=True) (scripts/CNNs.py:74,78,82,86) on NHWC."""
    B, T, F, C = x.shape
    T2, F2 = (T + 1) // 2, (F + 1) // 2
    xp = np.full((B, 2 * T2, 2 * F2, C), -np.inf, dtype=x.dtype)
    xp[:, :T, :F, :] = x
    return xp.reshape(B, T2, 2, F2, 2, C).max(axis=(2, 4))


def _zero_rows(x, lengths):
    if lengths is None:
        return x
    t = np.arange(x.shape[1])[None, :, None, None]
    return np.where(t < np.asarray(lengths)[:, None, None, None], x, 0).astype(x.dtype)


def vgg_forward(x, convs, lengths=None):
    """VGG3L/VGG4L.forward (scripts/CNNs.py:34-52,68-91).

    x ``[B,T,80]``; convs = [(w,b)] * 6 or 8.  Returns ``([B,T',C*F'], out_lengths)``
    with feature index ``c*F' + f`` (the transpose+view at CNNs.py:88-89).
    ``lengths`` (valid frames per utterance) applies the masking rule of SURVEY §5.7,
    which makes a padded batch reproduce per-utterance batch-1 reference runs:
    zero rows t >= L after every conv+ReLU; L halves (ceil) at each pool.
    """
    h = x[:, :, :, None]                      # CNNs.py:70 — NCHW [B,1,T,80] == NHWC with C=1
    L = None if lengths is None else np.asarray(lengths).copy()
    h = _zero_rows(h, L)
    for i in range(0, len(convs), 2):
        h = _zero_rows(relu(conv3x3_same(h, *convs[i])), L)          # CNNs.py:72
        h = _zero_rows(relu(conv3x3_same(h, *convs[i + 1])), L)      # CNNs.py:73
        h = maxpool2x2_ceil(h)                                        # CNNs.py:74
        if L is not None:
            L = (L + 1) // 2
    B, T2, F2, C = h.shape
    out = h.transpose(0, 1, 3, 2).reshape(B, T2, C * F2)             # CNNs.py:88-89
    return out, L


# --------------------------------------------------------------------------- poolings
def _softmax(s, axis):
    m = np.max(s, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0)
    e = np.exp(s - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def mha_forward(x, query, lengths=None):
    """innerKeyValueAttention + MultiHeadAttention.forward (scripts/poolings.py:73-80,100-109).

    x ``[B,T,D]``, query ``[dh,H]``.  Returns ctx ``[B,H,dh]``, alignment ``[B,T,H]``,
    lse ``[B,H]``.  NB the scale is 1/sqrt(H): d_k = query.size(-1) (poolings.py:75).
    """
    B, T, D = x.shape
    dh, H = query.shape
    xv = x.reshape(B, T, H, dh)                                      # poolings.py:102-103
    s = np.einsum('bthd,dh->bth', xv, query.astype(x.dtype)) / math.sqrt(H)   # :76 (diagonal of key@query)
    if lengths is not None:
        t = np.arange(T)[None, :, None]
        s = np.where(t < np.asarray(lengths)[:, None, None], s, -np.inf)
    m = s.max(axis=1, keepdims=True)
    e = np.exp(s - m)
    l = e.sum(axis=1, keepdims=True)
    p = e / l                                                        # :77 softmax over time
    ctx = np.einsum('bth,bthd->bhd', p, xv)                          # :78-79
    lse = (m + np.log(l))[:, 0, :]
    return ctx.astype(x.dtype), p.astype(x.dtype), lse.astype(x.dtype)


def head_attention(ctx, att, keep=None):
    """HeadAttention.forward, narrow path (scripts/poolings.py:45-51,61-71).

    ctx ``[B,H,dh]``, att ``[dh]`` or ``[dh,1]``; keep ``[B,H]`` bool = heads NOT dropped
    (training-mode head drop-out, poolings.py:39-43, with the RNG draw injected).
    Returns out ``[B,dh]``, w ``[B,H]``.
    """
    a = np.asarray(att).reshape(-1).astype(ctx.dtype)
    u = ctx @ a                                                      # :47
    if keep is not None:
        u = np.where(np.asarray(keep, bool), u, -np.inf)            # :42
    w = _softmax(u, axis=-1)                                         # :50
    out = np.einsum('bh,bhd->bd', w, ctx)                            # :68-69
    return out.astype(ctx.dtype), w.astype(ctx.dtype)


def dmha_forward(x, query, att, lengths=None, keep=None):
    """DoubleMHA.forward (scripts/poolings.py:126-129).  Returns dict with the module's
    outputs (``out [B,dh]``, ``align [B,T,H]``) and the intermediates the backward needs."""
    ctx, p, lse = mha_forward(x, query, lengths)
    out, w = head_attention(ctx, att, keep)
    return dict(out=out, align=p, ctx=ctx, lse=lse, w=w)


def dmha_backward(x, query, att, g_out, lengths=None, keep=None):
    """Closed-form gradient of DoubleMHA.forward's first output w.r.t. x, query, att
    (equals torch autograd through scripts/poolings.py:73-80,45-51,61-71)."""
    B, T, D = x.shape
    dh, H = query.shape
    f = dmha_forward(x, query, att, lengths, keep)
    c, p, w = f['ctx'], f['align'], f['w']
    a = np.asarray(att).reshape(-1).astype(x.dtype)
    xv = x.reshape(B, T, H, dh)
    inv = 1.0 / math.sqrt(H)
    dw = np.einsum('bd,bhd->bh', g_out, c)
    du = w * (dw - np.sum(w * dw, axis=1, keepdims=True))
    dc = w[:, :, None] * g_out[:, None, :] + du[:, :, None] * a[None, None, :]
    datt = np.einsum('bh,bhd->d', du, c)
    dp = np.einsum('bhd,bthd->bth', dc, xv)
    ds = p * (dp - np.einsum('bhd,bhd->bh', dc, c)[:, None, :])
    dx = p[..., None] * dc[:, None, :, :] + ds[..., None] * (query.T.astype(x.dtype) * inv)[None, None, :, :]
    dq = np.einsum('bth,bthd->dh', ds, xv) * inv
    return dict(dx=dx.reshape(B, T, D).astype(x.dtype), dquery=dq.astype(x.dtype),
                datt=datt.reshape(dh, 1).astype(x.dtype))


def attention_forward(x, att, lengths=None):
    """Attention.forward (scripts/poolings.py:22-27): one query over time, no scale."""
    a = np.asarray(att).reshape(-1).astype(x.dtype)
    s = x @ a
    if lengths is not None:
        s = np.where(np.arange(x.shape[1])[None, :] < np.asarray(lengths)[:, None], s, -np.inf)
    p = _softmax(s, axis=-1)
    return np.einsum('bt,btd->bd', p, x).astype(x.dtype), p[:, :, None].astype(x.dtype)


# --------------------------------------------------------------------------- embedding tail
def fc_tail(pooled, W1, b1, W2, b2, bn_w, bn_b, bn_mean, bn_var, eps=1e-5):
    """relu(fc1) -> b2(relu(fc2)) in eval mode (scripts/model.py:56-57).  b1 is NOT applied."""
    dt = pooled.dtype
    e1 = relu(pooled @ W1.T.astype(dt) + b1.astype(dt))
    e2 = relu(e1 @ W2.T.astype(dt) + b2.astype(dt))
    return ((e2 - bn_mean.astype(dt)) / np.sqrt(bn_var.astype(dt) + dt.type(eps)) * bn_w.astype(dt) + bn_b.astype(dt)).astype(dt)


def pool_forward(feats, sd, cfg, lengths=None):
    pm = cfg.pooling_method
    if pm == 'DoubleMHA':
        return dmha_forward(feats, sd['poolingLayer.utteranceAttention.query'],
                            sd['poolingLayer.headsAttention.att'], lengths)['out']
    if pm == 'MHA':
        ctx, _, _ = mha_forward(feats, sd['poolingLayer.query'], lengths)
        return ctx.reshape(ctx.shape[0], -1)
    if pm == 'Attention':
        return attention_forward(feats, sd['poolingLayer.att'], lengths)[0]
    raise ValueError(pm)


def get_embedding(x, sd, cfg, lengths=None, dtype=np.float32):
    """SpeakerClassifier.getEmbedding (scripts/model.py:52-59), eval mode."""
    from doubleattentionspeakerverification_b200.synth import conv_names
    x = np.asarray(x, dtype)
    convs = [(sd['front_end.%s.weight' % n].astype(dtype), sd['front_end.%s.bias' % n].astype(dtype))
             for n in conv_names(cfg.front_end)]
    feats, L = vgg_forward(x, convs, lengths)
    pooled = pool_forward(feats, sd, cfg, L)
    return fc_tail(pooled, sd['fc1.weight'], sd['fc1.bias'], sd['fc2.weight'], sd['fc2.bias'],
                   sd['b2.weight'], sd['b2.bias'], sd['b2.running_mean'], sd['b2.running_var'])


# --------------------------------------------------------------------------- scoring
def cosine_scores(e1, e2, eps=1e-8):
    """scoreCosineDistance = F.cosine_similarity(dim=-1, eps=1e-8) (scripts/utils.py:18-21), row-wise."""
    n1 = np.maximum(np.linalg.norm(e1, axis=-1), eps)
    n2 = np.maximum(np.linalg.norm(e2, axis=-1), eps)
    return np.sum(e1 * e2, axis=-1) / (n1 * n2)


def cosine_matrix(enrol, test, eps=1e-8):
    """All-pairs version of ``cosine_scores``: ``[Ne,E] x [Nt,E] -> [Ne,Nt]``."""
    a = enrol / np.maximum(np.linalg.norm(enrol, axis=-1, keepdims=True), eps)
    b = test / np.maximum(np.linalg.norm(test, axis=-1, keepdims=True), eps)
    return a @ b.T


def calculate_eer(CL, IM):
    """Trainer.__calculate_EER (scripts/train.py:135-150) over Score (scripts/utils.py:5-15), vectorised."""
    CL = np.asarray(CL, np.float64)
    IM = np.asarray(IM, np.float64)
    thresholds = np.arange(-1, 1, 0.01)
    FRR = np.array([round(float(np.sum(CL < th)) * 100 / float(len(CL)), 4) for th in thresholds])
    FAR = np.array([round(float(np.sum(IM >= th)) * 100 / float(len(IM)), 4) for th in thresholds])
    idx = np.argwhere(np.diff(np.sign(FAR - FRR)) != 0).reshape(-1)
    if len(idx) > 0:
        return round((FAR[int(idx[0])] + FRR[int(idx[0])]) / 2, 4)
    return 50.00
