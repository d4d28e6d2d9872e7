"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` is imported here and nowhere else;
the fixtures it writes are what pins ``oracle/path_oracle.py`` and
``oracle/torch_port.py`` (tests/test_oracle_golden.py) and what the GPU parity
tests compare against.  Inputs and weights are regenerated at test time from the
seeds recorded in each fixture (``doubleattentionspeakerverification_b200.synth``),
so only the reference's OUTPUTS are stored.

    python oracle/make_golden.py            # rewrites tests/golden/
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/scripts')

from doubleattentionspeakerverification_b200 import synth  # noqa: E402
import model as ref_model        # noqa: E402  (reference scripts/model.py)
import poolings as ref_poolings  # noqa: E402  (reference scripts/poolings.py)
import CNNs as ref_cnns          # noqa: E402  (reference scripts/CNNs.py)
import utils as ref_utils        # noqa: E402  (reference scripts/utils.py)

OUT = os.path.join(ROOT, 'tests', 'golden')
torch.set_num_threads(8)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def load_into(module, sd, prefix=''):
    own = module.state_dict()
    module.load_state_dict({k: t(sd[prefix + k]) for k in own})


def inject_keep(head_att, keep):
    """Replace ONLY the RNG draw of HeadAttention.__maskAttention (poolings.py:39-43, which
    needs torch.cuda.FloatTensor) with a supplied keep mask; the masked fill and everything
    downstream stay the reference's code."""
    def masked(score, mask_value=-float('inf')):
        score[~t(keep).view(score.size())] = mask_value
        return score
    head_att._HeadAttention__maskAttention = masked


def pooling_cases():
    shapes = [dict(batch=3, frames=37, dim=256, heads=8, seed=11),
              dict(batch=2, frames=50, dim=1024, heads=16, seed=12),
              dict(batch=2, frames=25, dim=5120, heads=32, seed=13),
              dict(batch=1, frames=9, dim=320, heads=8, seed=14),
              dict(batch=4, frames=203, dim=1024, heads=16, seed=15)]
    for i, shp in enumerate(shapes):
        c = synth.make_pooling_case(**shp)
        m = ref_poolings.DoubleMHA(shp['dim'], shp['heads'], mask_prob=0.3)
        with torch.no_grad():
            m.utteranceAttention.query.copy_(t(c['query']))
            m.headsAttention.att.copy_(t(c['att']))
        out = {}
        for mode in ('eval', 'train'):
            m.train(mode == 'train')
            if mode == 'train':
                inject_keep(m.headsAttention, c['keep'])
            x = t(c['x']).clone().requires_grad_(True)
            m.zero_grad()
            pooled, align = m(x)
            pooled = pooled.view(shp['batch'], -1)
            (pooled * t(c['g'])).sum().backward()
            out[mode + '_out'] = pooled.detach().numpy()
            out[mode + '_dquery'] = m.utteranceAttention.query.grad.numpy().copy()
            out[mode + '_datt'] = m.headsAttention.att.grad.numpy().copy()
            dx = x.grad.numpy()
            out[mode + '_dx_sample'] = dx[:, ::3, ::5].copy()
            if mode == 'eval':
                out['align'] = align.detach().numpy()
                with torch.no_grad():
                    ctx = m.utteranceAttention.getHeadsContextVectors(x)
                    _, head_align = m.getAlignments(x)
                out['ctx'] = ctx.numpy()
                out['head_align'] = head_align.numpy().reshape(shp['batch'], shp['heads'])
        np.savez_compressed(os.path.join(OUT, 'pooling_%d.npz' % i), shape=np.array(
            [shp['batch'], shp['frames'], shp['dim'], shp['heads'], shp['seed']]), **out)
        print('pooling', i, shp)
    # the single-query Attention pooling (poolings.py:14-27)
    c = synth.make_pooling_case(batch=3, frames=21, dim=320, heads=1, seed=21)
    m = ref_poolings.Attention(320)
    with torch.no_grad():
        m.att.copy_(t(c['att']))
        ct, p = m(t(c['x']))
    np.savez_compressed(os.path.join(OUT, 'attention_0.npz'), shape=np.array([3, 21, 320, 1, 21]),
                        out=ct.numpy(), align=p.numpy())


def frontend_cases():
    specs = [('VGG4L', 64, 2, 50, 31), ('VGG4L', 64, 1, 33, 32), ('VGG3L', 64, 2, 40, 33),
             ('VGG4L', 512, 1, 48, 34)]
    for i, (front, K, B, T, seed) in enumerate(specs):
        cfg = synth.example_config(front_end=front, kernel_size=K, embedding_size=32, heads_number=8, num_spkrs=4)
        sd = synth.make_state_dict(cfg, seed)
        net = (ref_cnns.VGG3L if front == 'VGG3L' else ref_cnns.VGG4L)(K).eval()
        load_into(net, sd, 'front_end.')
        x = synth.make_logmel(B, T, seed)
        with torch.no_grad():
            y = net(t(x)).numpy()
        np.savez_compressed(os.path.join(OUT, 'frontend_%d.npz' % i), front=front,
                            spec=np.array([K, B, T, seed]), out=y)
        print('frontend', i, front, K, y.shape)


def build_ref(cfg, sd):
    net = ref_model.SpeakerClassifier(Namespace(**vars(cfg)), 'cpu').eval()
    load_into(net, sd)
    return net


def embedding_cases():
    specs = [dict(name='small', kernel_size=64, embedding_size=32, heads_number=8, B=3, T=50, seed=41,
                  lengths=[50, 37, 23]),
             dict(name='small_vgg3', front_end='VGG3L', kernel_size=64, embedding_size=48, heads_number=16, B=2, T=41,
                  seed=42, lengths=[41, 30]),
             dict(name='k512', kernel_size=512, embedding_size=256, heads_number=16, B=2, T=64, seed=43,
                  lengths=[64, 45]),
             dict(name='example', kernel_size=1024, embedding_size=400, heads_number=32, B=1, T=400, seed=1234,
                  lengths=None),
             dict(name='example_b2', kernel_size=1024, embedding_size=400, heads_number=32, B=2, T=100, seed=44,
                  lengths=[100, 71]),
             dict(name='small_mha', pooling_method='MHA', kernel_size=64, embedding_size=32, heads_number=8, B=2,
                  T=50, seed=45, lengths=None),
             dict(name='small_att', pooling_method='Attention', kernel_size=64, embedding_size=32, heads_number=8,
                  B=2, T=50, seed=46, lengths=None)]
    for s in specs:
        s = dict(s)
        name, B, T, seed, lengths = s.pop('name'), s.pop('B'), s.pop('T'), s.pop('seed'), s.pop('lengths')
        cfg = synth.example_config(num_spkrs=7, **s)
        sd = synth.make_state_dict(cfg, seed)
        net = build_ref(cfg, sd)
        x = synth.make_logmel(B, T, seed)
        out = {}
        with torch.no_grad():
            out['emb'] = net.getEmbedding(t(x)).numpy()
            feats = net.front_end(t(x))
            out['feats_sample'] = feats.numpy()[:, :, ::17].copy()
            out['pooled'] = net.poolingLayer(feats)[0].reshape(B, -1).numpy()
            if lengths is not None:
                # oracle for a padded, length-masked batch = the reference run per utterance, batch 1, unpadded
                out['emb_varlen'] = np.concatenate(
                    [net.getEmbedding(t(x[b:b + 1, :L])).numpy() for b, L in enumerate(lengths)], 0)
                out['lengths'] = np.array(lengths, np.int32)
        np.savez_compressed(os.path.join(OUT, 'embed_%s.npz' % name), cfg=np.array(repr(vars(cfg))),
                            spec=np.array([B, T, seed]), **out)
        print('embedding', name, out['emb'].shape)


def ragged_case():
    """exampleModel config on 2-20 s utterances, each run through the live reference at batch 1, unpadded
    (scripts/train.py:107-131): the oracle of the padded + masked + bucketed extractor at T' up to 125."""
    Ts, seed0, wseed = synth.ragged_spec()
    cfg = synth.example_config(num_spkrs=7)
    net = build_ref(cfg, synth.make_state_dict(cfg, wseed))
    embs = []
    with torch.no_grad():
        for i, T in enumerate(Ts):
            embs.append(net.getEmbedding(t(synth.make_logmel(1, T, seed=seed0 + i))).numpy())
            print('ragged', i, T)
    np.savez_compressed(os.path.join(OUT, 'embed_ragged.npz'), cfg=np.array(repr(vars(cfg))), lengths=np.array(Ts, np.int32),
                        spec=np.array([seed0, wseed]), emb=np.concatenate(embs, 0))


def gradient_cases():
    """One train.py step (scripts/train.py:215-220): net(input, label, step) -> CrossEntropyLoss -> backward, from the
    live reference in train mode (batch-statistics BatchNorm1d, AM-Softmax margin, head drop-out with an injected keep
    mask).  Stores the loss, both outputs and every parameter gradient (strided sample + L2 norm for the large model)."""
    for spec in synth.TRAIN_STEP_SPECS:
        spec = dict(spec)
        name, stride = spec.pop('name'), spec.pop('stride')
        x, label, keep = synth.train_step_inputs(spec)
        cfg = synth.train_step_config(spec)
        net = ref_model.SpeakerClassifier(Namespace(**vars(cfg)), 'cpu')
        load_into(net, synth.make_state_dict(cfg, spec['seed']))
        net.train()
        inject_keep(net.poolingLayer.headsAttention, keep)
        pred, am = net(t(x), label=t(label), step=0)
        loss = torch.nn.CrossEntropyLoss()(am, t(label))
        loss.backward()
        out = dict(loss=np.array(float(loss.detach())), pred=pred.detach().numpy(), am=am.detach().numpy(),
                   b2_running_mean=net.b2.running_mean.numpy().copy(), b2_running_var=net.b2.running_var.numpy().copy())
        for n, p in net.named_parameters():
            if p.grad is None:
                continue                                            # b1, b3: unused by forward (model.py:61-71)
            g = p.grad.numpy().reshape(-1)
            out['grad.' + n] = g[::synth.grad_sample_stride(g.size, stride)].copy()
            out['norm.' + n] = np.array(np.linalg.norm(g.astype(np.float64)))
        np.savez_compressed(os.path.join(OUT, 'grad_%s.npz' % name), cfg=np.array(repr(vars(cfg))),
                            spec=np.array([spec['B'], spec['T'], spec['seed'], stride]), **out)
        print('gradient', name, float(loss.detach()))


def scoring_case():
    rs = np.random.RandomState(51)
    e1 = rs.standard_normal((64, 400)).astype(np.float32)
    e2 = rs.standard_normal((64, 400)).astype(np.float32)
    s = ref_utils.scoreCosineDistance(t(e1), t(e2)).numpy()
    np.savez_compressed(os.path.join(OUT, 'cosine_0.npz'), seed=np.array(51), scores=s)


def eer_synthetic_scores(seed, n_cl, n_im, sep):
    rs = np.random.RandomState(seed)
    CL = (0.5 + sep + 0.2 * rs.standard_normal(n_cl)).clip(-1, 1).astype(np.float32)
    IM = (0.5 - sep + 0.2 * rs.standard_normal(n_im)).clip(-1, 1).astype(np.float32)
    return CL, IM


def eer_case():
    """Trainer.__calculate_EER (scripts/train.py:135-150) from the live reference.  train.py imports data.py, which
    imports soundfile (not installed, unused by the EER code): an empty stub module lets the import through."""
    import types
    sys.modules.setdefault('soundfile', types.ModuleType('soundfile'))
    import train as ref_train    # reference scripts/train.py
    specs = [(61, 500, 700, 0.2), (62, 1000, 3000, 0.05), (63, 64, 64, 0.6), (64, 200, 100, -0.1)]
    eers = []
    for seed, n_cl, n_im, sep in specs:
        CL, IM = eer_synthetic_scores(seed, n_cl, n_im, sep)
        eers.append(ref_train.Trainer._Trainer__calculate_EER(None, [float(v) for v in CL], [float(v) for v in IM]))
    np.savez_compressed(os.path.join(OUT, 'eer_0.npz'), specs=np.array(specs, np.float64), eer=np.array(eers, np.float64))
    print('eer', eers)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:                     # regenerate single groups: eer | ragged | grad
        for what in sys.argv[1:]:
            {'eer': eer_case, 'ragged': ragged_case, 'grad': gradient_cases}[what]()
        sys.exit(0)
    pooling_cases()
    frontend_cases()
    embedding_cases()
    ragged_case()
    gradient_cases()
    scoring_case()
    eer_case()
    print('golden fixtures written to', OUT)
