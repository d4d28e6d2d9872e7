"""Deterministic synthetic weights and log-mel inputs (pure numpy, no GPU).

The reference's example weights are Git-LFS pointers and no dataset is reachable
(SURVEY.md §0), so every test, golden fixture and benchmark uses random-init
weights of the reference architecture and synthetic CMN'd log-mel input.  The
generators use ``numpy.random.RandomState`` (bit-stable by numpy policy), so a
fixture only needs to store a seed.

Keys and shapes are the reference ``state_dict`` contract
(``scripts/model.py:10-50``, ``scripts/CNNs.py:24-32,56-66``,
``scripts/poolings.py:84-91,31-37``).
"""
import math
from argparse import Namespace

import numpy as np

FEATURE_SIZE = 80  # hard-coded by the reference, scripts/model.py:13


def example_config(**over):
    """exampleModel == train.py defaults (scripts/train.py:263,269-273)."""
    cfg = dict(front_end='VGG4L', kernel_size=1024, embedding_size=400, heads_number=32,
               pooling_method='DoubleMHA', mask_prob=0.3, scalingFactor=30.0,
               marginFactor=0.4, annealing=False, num_spkrs=5994)
    cfg.update(over)
    return Namespace(**cfg)


def vgg_channels(front_end, kernel_size):
    """(Cin, Cout) of every conv, in order (scripts/CNNs.py:27-32,59-66)."""
    k = int(kernel_size)
    if front_end == 'VGG3L':
        outs = [k // 4, k // 4, k // 2, k // 2, k, k]
    elif front_end == 'VGG4L':
        outs = [k // 8, k // 8, k // 4, k // 4, k // 2, k // 2, k, k]
    else:
        raise ValueError('unknown front_end %r' % (front_end,))
    ins = [1] + outs[:-1]
    return list(zip(ins, outs))


def conv_names(front_end):
    nblocks = 3 if front_end == 'VGG3L' else 4
    return ['conv%d%d' % (b, i) for b in range(1, nblocks + 1) for i in (1, 2)]


def vgg_output_dim(front_end, kernel_size, feature_size=FEATURE_SIZE):
    """ceil(80 / 2^n) * K  (scripts/CNNs.py:7-20)."""
    f = feature_size
    for _ in range(3 if front_end == 'VGG3L' else 4):
        f = (f + 1) // 2
    return f * int(kernel_size)


def make_state_dict(cfg, seed=1234, randomize_bn=True):
    """Random-init weights keyed exactly like the reference ``state_dict``.

    He-scaled normal conv/linear weights keep activations O(1) through 8 conv
    layers; BN running stats/affine are randomised so a broken BN epilogue
    cannot hide behind the identity.
    """
    rs = np.random.RandomState(seed)
    sd = {}
    for name, (cin, cout) in zip(conv_names(cfg.front_end), vgg_channels(cfg.front_end, cfg.kernel_size)):
        std = math.sqrt(2.0 / (9 * cin))
        sd['front_end.%s.weight' % name] = (rs.standard_normal((cout, cin, 3, 3)) * std).astype(np.float32)
        sd['front_end.%s.bias' % name] = (rs.standard_normal((cout,)) * 0.05).astype(np.float32)
    D = vgg_output_dim(cfg.front_end, cfg.kernel_size)
    E = int(cfg.embedding_size)
    H = int(getattr(cfg, 'heads_number', 1))
    pm = cfg.pooling_method
    if pm == 'DoubleMHA':
        dh = D // H
        sd['poolingLayer.utteranceAttention.query'] = (rs.standard_normal((dh, H)) * math.sqrt(2.0 / (dh + H))).astype(np.float32)
        sd['poolingLayer.headsAttention.att'] = (rs.standard_normal((dh, 1)) * math.sqrt(2.0 / (dh + 1))).astype(np.float32)
        vec = dh
    elif pm == 'MHA':
        dh = D // H
        sd['poolingLayer.query'] = (rs.standard_normal((dh, H)) * math.sqrt(2.0 / (dh + H))).astype(np.float32)
        vec = D
    elif pm == 'Attention':
        sd['poolingLayer.att'] = (rs.standard_normal((D, 1)) * math.sqrt(2.0 / (D + 1))).astype(np.float32)
        vec = D
    else:
        raise ValueError('unknown pooling_method %r' % (pm,))

    def linear(prefix, fin, fout):
        sd[prefix + '.weight'] = (rs.standard_normal((fout, fin)) * math.sqrt(2.0 / fin)).astype(np.float32)
        sd[prefix + '.bias'] = (rs.standard_normal((fout,)) * 0.05).astype(np.float32)

    def bnorm(prefix, n):
        if randomize_bn:
            sd[prefix + '.weight'] = (1.0 + 0.2 * rs.standard_normal((n,))).astype(np.float32)
            sd[prefix + '.bias'] = (0.1 * rs.standard_normal((n,))).astype(np.float32)
            sd[prefix + '.running_mean'] = (0.3 * rs.standard_normal((n,))).astype(np.float32)
            sd[prefix + '.running_var'] = (0.5 + rs.uniform(0.0, 1.0, (n,))).astype(np.float32)
        else:
            sd[prefix + '.weight'] = np.ones((n,), np.float32)
            sd[prefix + '.bias'] = np.zeros((n,), np.float32)
            sd[prefix + '.running_mean'] = np.zeros((n,), np.float32)
            sd[prefix + '.running_var'] = np.ones((n,), np.float32)
        sd[prefix + '.num_batches_tracked'] = np.array(0, np.int64)

    linear('fc1', vec, E)
    bnorm('b1', E)
    linear('fc2', E, E)
    bnorm('b2', E)
    linear('preLayer', E, E)
    bnorm('b3', E)
    S = int(cfg.num_spkrs)
    sd['predictionLayer.W'] = (rs.standard_normal((E, S)) * math.sqrt(2.0 / (E + S))).astype(np.float32)
    return sd


def make_logmel(batch, frames, seed=0, feature_size=FEATURE_SIZE):
    """Synthetic CMN'd log-mel ``[B, T, 80]``: 2*randn minus the per-utterance column
    mean (the reference's CMN, scripts/featureExtractor.py:25-26)."""
    rs = np.random.RandomState(seed)
    x = 2.0 * rs.standard_normal((batch, frames, feature_size))
    x -= x.mean(axis=1, keepdims=True)
    return x.astype(np.float32)


def make_lengths(batch, lo, hi, seed=0):
    rs = np.random.RandomState(seed)
    return rs.randint(lo, hi + 1, size=(batch,)).astype(np.int32)


def make_waveform(n_samples, sfr=16000, seed=0):
    """A speech-like synthetic waveform in [-1, 1): a few drifting harmonics under a slow envelope plus noise
    (float64, like ``soundfile.read``)."""
    rs = np.random.RandomState(seed)
    t = np.arange(n_samples) / float(sfr)
    y = np.zeros(n_samples)
    f0 = 90.0 + 80.0 * rs.rand()
    for k in range(1, 12):
        y += (rs.rand() / k) * np.sin(2 * np.pi * k * f0 * t * (1.0 + 0.05 * np.sin(2 * np.pi * 1.7 * t)) + 6.28 * rs.rand())
    env = 0.5 + 0.5 * np.sin(2 * np.pi * 3.1 * t + rs.rand())
    y = 0.2 * env * y / max(np.abs(y).max(), 1e-9) + 0.01 * rs.standard_normal(n_samples)
    return np.clip(y, -1.0, 1.0 - 1.0 / 32768)


def make_pooling_case(batch, frames, dim, heads, seed=0, with_lengths=False):
    """Inputs for a DoubleMHA pooling test: x ``[B,T,D]``, query ``[dh,H]``, att ``[dh,1]``,
    upstream gradient g ``[B,dh]``, keep mask ``[B,H]`` (>=1 head kept per row), lengths."""
    rs = np.random.RandomState(seed)
    dh = dim // heads
    x = rs.standard_normal((batch, frames, dim)).astype(np.float32)
    # 4x xavier so both softmaxes are clearly non-uniform (score std ~1.4, not ~0.3)
    query = (rs.standard_normal((dh, heads)) * 4.0 * math.sqrt(2.0 / (dh + heads))).astype(np.float32)
    att = (rs.standard_normal((dh, 1)) * 4.0 * math.sqrt(2.0 / (dh + 1))).astype(np.float32)
    g = rs.standard_normal((batch, dh)).astype(np.float32)
    keep = rs.randint(0, 3, size=(batch, heads)) > 0
    keep[:, 0] = True
    lengths = None
    if with_lengths:
        lengths = rs.randint(max(1, frames // 2), frames + 1, size=(batch,)).astype(np.int32)
        lengths[0] = frames
    return dict(x=x, query=query, att=att, g=g, keep=keep, lengths=lengths)


def ragged_spec():
    """BASELINE configs[3] at test scale: ten 2-20 s utterances (T in [200, 2000]).  Returns (frame counts, input seed of
    utterance 0, weight seed); utterance i is ``make_logmel(1, T_i, seed0 + i)``."""
    rs = np.random.RandomState(3)
    return [int(T) for T in rs.randint(200, 2001, size=10)], 100, 1234


# training-step fixtures (tests/golden/grad_<name>.npz): model + batch of one train.py step; `stride` = sampling of the
# flattened gradients stored in the fixture
TRAIN_STEP_SPECS = [dict(name='small', kernel_size=64, embedding_size=32, heads_number=8, num_spkrs=6, B=4, T=48, seed=3, stride=1),
                    dict(name='k512', kernel_size=512, embedding_size=64, heads_number=16, num_spkrs=10, B=4, T=48, seed=5, stride=61)]


def grad_sample_stride(numel, stride):
    """Gradients of small tensors are stored whole in the fixtures, large ones as every `stride`-th element."""
    return stride if numel > 4096 else 1


def train_step_config(spec):
    return example_config(**{k: v for k, v in spec.items() if k not in ('name', 'stride', 'B', 'T', 'seed')})


def train_step_inputs(spec):
    """Inputs of a training-step fixture: log-mel batch, labels, injected head keep mask."""
    B, T, seed = spec['B'], spec['T'], spec['seed']
    x = make_logmel(B, T, seed)
    rs = np.random.RandomState(seed + 1000)
    label = rs.randint(0, spec['num_spkrs'], size=(B,)).astype(np.int64)
    keep = make_pooling_case(B, 3, vgg_output_dim('VGG4L', spec['kernel_size']), spec['heads_number'], seed=seed)['keep']
    return x, label, keep


def load_state_dict(module, sd, prefix=''):
    """Load a ``make_state_dict`` dictionary (numpy) into a torch module of the reference's layout."""
    import torch
    own = module.state_dict()
    module.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(sd[prefix + k])) for k in own})
    return module
