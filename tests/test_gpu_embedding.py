"""GPU parity of SpeakerClassifier.getEmbedding (front-end + pooling + fused FC/BN tail), trial
scoring, and a train.py-style step through the fused pooling backward."""
import ast
from argparse import Namespace

import numpy as np
import pytest
import torch

from conftest import golden, max_rel, min_cosine
from doubleattentionspeakerverification_b200 import model, ops, synth, utils
from oracle import path_oracle as po

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build(g, precision):
    cfg = Namespace(**ast.literal_eval(str(g['cfg'])))
    cfg.precision = precision
    B, T, seed = [int(v) for v in g['spec']]
    net = model.SpeakerClassifier(cfg, 'cuda')
    synth.load_state_dict(net, synth.make_state_dict(cfg, seed))
    return net.cuda().eval(), synth.make_logmel(B, T, seed)


@pytest.mark.parametrize('name', ['small', 'small_vgg3', 'k512', 'example', 'example_b2', 'small_mha', 'small_att'])
def test_embedding_fp32(name):
    g = golden('embed_%s.npz' % name)
    net, x = build(g, 'fp32')
    with torch.no_grad():
        emb = net.getEmbedding(dev(x))
    assert max_rel(emb.cpu().numpy(), g['emb']) < 1e-4               # north_star fp32 bar
    if 'emb_varlen' in g.files:
        with torch.no_grad():
            ev = net.getEmbedding(dev(x), lengths=dev(g['lengths']))
        assert max_rel(ev.cpu().numpy(), g['emb_varlen']) < 1e-4     # padded batch == per-utterance batch-1 reference


@pytest.mark.parametrize('name', ['k512', 'example', 'example_b2'])
def test_embedding_bf16(name):
    g = golden('embed_%s.npz' % name)
    net, x = build(g, 'bf16')
    assert net.front_end.resolved_precision() == 'bf16'
    with torch.no_grad():
        emb = net.getEmbedding(dev(x))
    assert min_cosine(emb.cpu().numpy(), g['emb']) >= 0.9999         # north_star bf16 bar
    if 'emb_varlen' in g.files:
        with torch.no_grad():
            ev = net.getEmbedding(dev(x), lengths=dev(g['lengths']))
        assert min_cosine(ev.cpu().numpy(), g['emb_varlen']) >= 0.9999
        s_ref = po.cosine_scores(g['emb_varlen'][:1], g['emb_varlen'][1:2])
        s = utils.scoreCosineDistance(ev[:1], ev[1:2]).cpu().numpy()
        assert abs(float(s[0]) - float(s_ref[0])) < 1e-3             # north_star trial-score bar


def test_state_dict_contract():
    cfg = synth.example_config(kernel_size=64, embedding_size=32, heads_number=8, num_spkrs=5)
    net = model.SpeakerClassifier(cfg, 'cuda')
    sd = synth.make_state_dict(cfg, 1)
    own = net.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k in own:
        assert tuple(own[k].shape) == tuple(np.shape(sd[k])), k


def test_scoring():
    g = golden('cosine_0.npz')
    rs = np.random.RandomState(int(g['seed']))
    e1 = rs.standard_normal((64, 400)).astype(np.float32)
    e2 = rs.standard_normal((64, 400)).astype(np.float32)
    s = utils.scoreCosineDistance(dev(e1), dev(e2)).cpu().numpy()
    assert np.max(np.abs(s - g['scores'])) < 1e-6
    m = utils.score_matrix(dev(e1), dev(e2)).cpu().numpy()
    assert np.max(np.abs(m - po.cosine_matrix(e1, e2))) < 1e-5
    emb = dev(np.concatenate([e1, e2]))
    ia = torch.arange(64, device='cuda')
    assert np.max(np.abs(utils.score_pairs(emb, ia, ia + 64).cpu().numpy() - g['scores'])) < 1e-6
    one = utils.scoreCosineDistance(dev(e1[:1]), dev(e2[:1]))
    assert one.shape == (1,)


def test_training_step_through_fused_pooling():
    """train.py-style step (scripts/train.py:215-226): forward(x, label) -> CE -> backward; the front-end
    runs through torch autograd, the pooling through the hand-written backward.  Gradients must match
    the same model with the pooling expressed in plain torch ops."""
    torch.manual_seed(0)
    cfg = synth.example_config(kernel_size=64, embedding_size=32, heads_number=8, num_spkrs=6)
    net = model.SpeakerClassifier(cfg, 'cuda')
    synth.load_state_dict(net, synth.make_state_dict(cfg, 3))
    net = net.cuda().train()
    x = dev(synth.make_logmel(4, 48, 3))
    label = torch.tensor([0, 3, 5, 1], device='cuda')
    keep = torch.tensor(synth.make_pooling_case(4, 3, 320, 8, seed=1)['keep'], device='cuda')

    def run(fused):
        net.zero_grad()
        feats = net.front_end(x)
        if fused:
            e0, _ = net.poolingLayer(feats, keep=keep)
        else:
            q, a = net.poolingLayer.utteranceAttention.query, net.poolingLayer.headsAttention.att
            B, T, D = feats.shape
            H = 8
            xv = feats.view(B, T, H, D // H)
            p = torch.softmax(torch.einsum('bthd,dh->bth', xv, q) / np.sqrt(H), dim=1)
            ctx = torch.einsum('bth,bthd->bhd', p, xv)
            u = (ctx @ a).squeeze(-1).masked_fill(~keep, float('-inf'))
            e0 = torch.einsum('bh,bhd->bd', torch.softmax(u, -1), ctx)
        e1 = torch.relu(net.fc1(e0))
        e2 = net.b2(torch.relu(net.fc2(e1)))
        pred, logits = net.predictionLayer(net.preLayer(e2), label, 0)
        loss = torch.nn.functional.cross_entropy(logits, label)
        loss.backward()
        return float(loss), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}

    l_ref, g_ref = run(False)
    l_fused, g_fused = run(True)
    assert abs(l_ref - l_fused) < 1e-4 * max(1.0, abs(l_ref))
    for n in g_ref:
        assert max_rel(g_fused[n].cpu().numpy(), g_ref[n].cpu().numpy()) < 2e-3, n
    assert 'poolingLayer.utteranceAttention.query' in g_fused and 'front_end.conv11.weight' in g_fused
