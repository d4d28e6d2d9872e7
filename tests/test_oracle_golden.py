"""Pin the CPU oracle (numpy restatement + torch port) to the reference's own outputs.

The fixtures in tests/golden were produced by oracle/make_golden.py running the live
reference (scripts/model.py, poolings.py, CNNs.py, utils.py) in the build container.
"""
import ast
from argparse import Namespace

import numpy as np
import pytest
import torch

from conftest import golden, max_rel
from doubleattentionspeakerverification_b200 import synth
from oracle import path_oracle as po
from oracle import torch_port as tp

TOL = 2e-5  # fp32 vs fp32, different summation order


@pytest.mark.parametrize('idx', range(5))
def test_pooling_forward_backward(idx):
    g = golden('pooling_%d.npz' % idx)
    B, T, D, H, seed = [int(v) for v in g['shape']]
    c = synth.make_pooling_case(B, T, D, H, seed)
    f = po.dmha_forward(c['x'], c['query'], c['att'])
    assert max_rel(f['out'], g['eval_out']) < TOL
    assert max_rel(f['align'], g['align']) < TOL
    assert max_rel(f['ctx'], g['ctx']) < TOL
    assert max_rel(f['w'], g['head_align']) < TOL
    for mode, keep in (('eval', None), ('train', c['keep'])):
        ff = po.dmha_forward(c['x'], c['query'], c['att'], keep=keep)
        assert max_rel(ff['out'], g[mode + '_out']) < TOL
        b = po.dmha_backward(c['x'], c['query'], c['att'], c['g'], keep=keep)
        assert max_rel(b['dx'][:, ::3, ::5], g[mode + '_dx_sample']) < 5e-5
        assert max_rel(b['dquery'], g[mode + '_dquery']) < 5e-5
        assert max_rel(b['datt'], g[mode + '_datt']) < 5e-5
    # torch port (the CPU baseline) against the same fixture
    out, p = tp.double_mha(torch.from_numpy(c['x']), torch.from_numpy(c['query']), torch.from_numpy(c['att']))
    assert max_rel(out.numpy(), g['eval_out']) < TOL
    assert max_rel(p.numpy(), g['align']) < TOL


def test_pooling_length_mask_equals_truncation():
    c = synth.make_pooling_case(4, 40, 256, 8, seed=5, with_lengths=True)
    f = po.dmha_forward(c['x'], c['query'], c['att'], lengths=c['lengths'])
    for b, L in enumerate(c['lengths']):
        fb = po.dmha_forward(c['x'][b:b + 1, :L], c['query'], c['att'])
        assert max_rel(f['out'][b:b + 1], fb['out']) < 1e-6
        assert np.all(f['align'][b, L:] == 0)


def test_attention_pooling():
    g = golden('attention_0.npz')
    B, T, D, H, seed = [int(v) for v in g['shape']]
    c = synth.make_pooling_case(B, T, D, H, seed)
    out, p = po.attention_forward(c['x'], c['att'])
    assert max_rel(out, g['out']) < TOL and max_rel(p, g['align']) < TOL


@pytest.mark.parametrize('idx', range(4))
def test_frontend(idx):
    g = golden('frontend_%d.npz' % idx)
    front = str(g['front'])
    K, B, T, seed = [int(v) for v in g['spec']]
    cfg = synth.example_config(front_end=front, kernel_size=K, embedding_size=32, heads_number=8, num_spkrs=4)
    sd = synth.make_state_dict(cfg, seed)
    x = synth.make_logmel(B, T, seed)
    convs = [(sd['front_end.%s.weight' % n], sd['front_end.%s.bias' % n]) for n in synth.conv_names(front)]
    y, _ = po.vgg_forward(x, convs)
    assert y.shape == g['out'].shape
    assert max_rel(y, g['out']) < TOL
    yt = tp.front_end(torch.from_numpy(x), tp.as_torch(sd), front).numpy()
    assert max_rel(yt, g['out']) < TOL


def _cfg_of(g):
    return Namespace(**ast.literal_eval(str(g['cfg'])))


@pytest.mark.parametrize('name', ['small', 'small_vgg3', 'k512', 'example_b2', 'small_mha', 'small_att'])
def test_embedding(name):
    g = golden('embed_%s.npz' % name)
    cfg = _cfg_of(g)
    B, T, seed = [int(v) for v in g['spec']]
    sd = synth.make_state_dict(cfg, seed)
    x = synth.make_logmel(B, T, seed)
    emb = po.get_embedding(x, sd, cfg)
    assert max_rel(emb, g['emb']) < 1e-4
    embt = tp.get_embedding(torch.from_numpy(x), tp.as_torch(sd), cfg).numpy()
    assert max_rel(embt, g['emb']) < 1e-4
    if 'emb_varlen' in g.files:
        # padded + length-masked batch == the reference run per utterance at batch 1
        embv = po.get_embedding(x, sd, cfg, lengths=g['lengths'])
        assert max_rel(embv, g['emb_varlen']) < 1e-4


def test_embedding_example_config():
    """configs[0]: exampleModel config, one 4 s utterance (torch port only; numpy conv of 52 GFLOP is slow)."""
    g = golden('embed_example.npz')
    cfg = _cfg_of(g)
    B, T, seed = [int(v) for v in g['spec']]
    sd = synth.make_state_dict(cfg, seed)
    x = synth.make_logmel(B, T, seed)
    emb = tp.get_embedding(torch.from_numpy(x), tp.as_torch(sd), cfg).numpy()
    assert emb.shape == (1, 400)
    assert max_rel(emb, g['emb']) < 1e-4


def test_cosine():
    g = golden('cosine_0.npz')
    rs = np.random.RandomState(int(g['seed']))
    e1 = rs.standard_normal((64, 400)).astype(np.float32)
    e2 = rs.standard_normal((64, 400)).astype(np.float32)
    assert np.max(np.abs(po.cosine_scores(e1, e2) - g['scores'])) < 1e-6
    assert np.max(np.abs(np.diag(po.cosine_matrix(e1, e2)) - g['scores'])) < 1e-6
    assert np.max(np.abs(tp.cosine(torch.from_numpy(e1), torch.from_numpy(e2)).numpy() - g['scores'])) < 1e-6


def test_eer():
    g = golden('eer_0.npz')
    for (seed, n_cl, n_im, sep), want in zip(g['specs'], g['eer']):
        rs = np.random.RandomState(int(seed))
        CL = (0.5 + sep + 0.2 * rs.standard_normal(int(n_cl))).clip(-1, 1).astype(np.float32)
        IM = (0.5 - sep + 0.2 * rs.standard_normal(int(n_im))).clip(-1, 1).astype(np.float32)
        assert po.calculate_eer(CL, IM) == want


# ------------------------------------------------------------------------------------ features
@pytest.mark.parametrize('idx', range(3))
def test_feature_oracle_matches_independent_implementation(idx):
    """oracle/feature_oracle.py (restated librosa 0.7.2 stft + filters.mel) vs vectors made with transformers.audio_utils
    (oracle/make_golden_features.py)."""
    from oracle import feature_oracle as fo
    g = golden('logmel_%d.npz' % idx)
    sfr, n, seed = [int(v) for v in g['shape']]
    y = synth.make_waveform(n, sfr, seed)
    assert np.abs(fo.mel_filterbank(sfr) - g['melw']).max() < 1e-7
    assert np.abs(fo.mfsc(y, sfr) - g['mfsc']).max() < 1e-5
    assert np.abs(fo.extract(y, sfr) - g['feat']).max() < 1e-5


def test_feature_host_tables_match_the_oracle():
    """The product's own window / mel-filterbank construction (featureExtractor._tables) against the oracle's."""
    from oracle import feature_oracle as fo
    from doubleattentionspeakerverification_b200 import featureExtractor as fe
    for sfr, n_mels, window in ((16000, 80, 'hamming'), (8000, 40, 'hann'), (16000, 24, 'hamming')):
        win_length = int(sfr * 0.025)
        taps, melw, rng = fe._tables(sfr, win_length, window, n_mels)
        lpad = (512 - win_length) // 2
        assert np.abs(fo.fft_window(window, win_length)[lpad:lpad + win_length] - taps).max() < 1e-7
        assert np.abs(melw - fo.mel_filterbank(sfr, 512, n_mels)).max() < 1e-6
        assert np.array_equal(rng, fo.mel_ranges(melw))
    assert fe.frames_for(np.array([511, 512, 671, 672, 16000]), 160).tolist() == [0, 1, 1, 2, 97]


def test_ragged_fixture_short_utterances():
    """embed_ragged.npz (2-20 s utterances through the live reference at batch 1): the torch port on the shortest ones."""
    g = golden('embed_ragged.npz')
    Ts, seed0, wseed = synth.ragged_spec()
    assert list(g['lengths']) == Ts and [int(v) for v in g['spec']] == [seed0, wseed]
    cfg = _cfg_of(g)
    sd = tp.as_torch(synth.make_state_dict(cfg, wseed))
    for i in np.argsort(Ts)[:2]:
        e = tp.get_embedding(torch.from_numpy(synth.make_logmel(1, Ts[i], seed=seed0 + int(i))), sd, cfg).numpy()
        assert max_rel(e, g['emb'][i:i + 1]) < TOL


@pytest.mark.parametrize('spec', synth.TRAIN_STEP_SPECS, ids=lambda s: s['name'])
def test_training_step_oracle(spec):
    """oracle/torch_port.training_step == one train.py step of the live reference (loss, outputs, every gradient)."""
    g = golden('grad_%s.npz' % spec['name'])
    cfg = synth.train_step_config(spec)
    x, label, keep = synth.train_step_inputs(spec)
    sd = tp.as_torch(synth.make_state_dict(cfg, spec['seed']))
    loss, pred, logits, grads = tp.training_step(torch.from_numpy(x), torch.from_numpy(label), sd, cfg, torch.from_numpy(keep))
    assert abs(loss - float(g['loss'])) < 1e-5 * abs(float(g['loss']))
    assert max_rel(pred.numpy(), g['pred']) < TOL and max_rel(logits.numpy(), g['am']) < TOL
    names = [k[5:] for k in g.files if k.startswith('grad.')]
    assert set(names) == set(grads.keys())
    for n in names:
        got = grads[n].numpy().reshape(-1)
        assert max_rel(got[::synth.grad_sample_stride(got.size, spec['stride'])], g['grad.' + n]) < 1e-4, n
        assert abs(np.linalg.norm(got.astype(np.float64)) - float(g['norm.' + n])) < 1e-4 * float(g['norm.' + n]), n
