"""The import swap INTEGRATION.md §2 describes, exercised on the reference's OWN ``scripts/model.py``: its three operator
imports (``from poolings import *``, ``from CNNs import *``, ``from loss import *``, scripts/model.py:4-6) are pointed at
this package through ``sys.modules`` and the reference's unmodified ``SpeakerClassifier`` is assembled from them.

Runs only where the reference checkout exists (this container; the GPU box has no /root/reference, and there is no CPU
path to run the operators here), so it checks what the swap must guarantee without a GPU: the reference's constructor
finds every name it uses with the signature it uses, the assembled module has the reference's parameters (names, shapes,
dtypes), a reference state_dict loads into it, and the reference's own ``getEmbedding`` / ``forward`` code reaches this
package's operators (they refuse a CPU tensor loudly instead of computing anything).
"""
import importlib.util
import os
import sys

import pytest
import torch

from doubleattentionspeakerverification_b200 import CNNs, loss, poolings, synth

REF_SCRIPTS = '/root/reference/scripts'
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF_SCRIPTS, 'model.py')),
                                reason='needs the reference checkout (not present on the GPU box)')

OPERATOR_MODULES = ('poolings', 'CNNs', 'loss')


def _load_reference_model(tag, swapped):
    """Execute the reference's scripts/model.py as module ``tag`` with its operator imports resolved either to this
    package (``swapped``) or to the reference's own files."""
    saved = {n: sys.modules.get(n) for n in OPERATOR_MODULES}
    path_added = False
    try:
        for n in OPERATOR_MODULES:
            sys.modules.pop(n, None)
        if swapped:
            sys.modules.update(poolings=poolings, CNNs=CNNs, loss=loss)
        else:
            sys.path.insert(0, REF_SCRIPTS)
            path_added = True
        spec = importlib.util.spec_from_file_location(tag, os.path.join(REF_SCRIPTS, 'model.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        if path_added:
            sys.path.remove(REF_SCRIPTS)
        for n, m in saved.items():
            sys.modules.pop(n, None)
            if m is not None:
                sys.modules[n] = m


@pytest.mark.parametrize('front_end,pooling', [('VGG4L', 'DoubleMHA'), ('VGG3L', 'DoubleMHA'), ('VGG4L', 'MHA'), ('VGG4L', 'Attention')])
def test_reference_model_file_assembles_from_this_package(front_end, pooling):
    cfg = synth.example_config(kernel_size=64, embedding_size=32, heads_number=8, num_spkrs=5)
    cfg.front_end, cfg.pooling_method = front_end, pooling
    swapped = _load_reference_model('dasv_swapped_model', True)
    stock = _load_reference_model('dasv_stock_model', False)
    torch.manual_seed(0)
    ours = swapped.SpeakerClassifier(cfg, 'cpu')
    torch.manual_seed(0)
    ref = stock.SpeakerClassifier(cfg, 'cpu')
    # the operators inside the reference's classifier are this package's
    assert type(ours.front_end).__module__ == CNNs.__name__
    assert type(ours.poolingLayer).__module__ == poolings.__name__
    assert type(ours.predictionLayer).__module__ == loss.__name__
    assert type(ref.front_end).__module__ == 'CNNs'
    sd_ours, sd_ref = ours.state_dict(), ref.state_dict()
    assert list(sd_ours.keys()) == list(sd_ref.keys())
    for k in sd_ref:
        assert sd_ours[k].shape == sd_ref[k].shape and sd_ours[k].dtype == sd_ref[k].dtype, k
    assert ours.vector_size == ref.vector_size
    ours.load_state_dict(sd_ref)                      # a reference checkpoint loads as is (train.py:107-115)
    for k in sd_ref:
        assert torch.equal(ours.state_dict()[k], sd_ref[k]), k
    # the reference's getEmbedding / forward bodies (scripts/model.py:52-71) run into this package's operators
    x = torch.randn(2, 24, 80)
    with pytest.raises(Exception, match='CUDA'):
        with torch.no_grad():
            ours.eval().getEmbedding(x)
    with pytest.raises(Exception, match='CUDA'):
        ours.train()(x, torch.tensor([0, 1]), 0)
    with torch.no_grad():
        assert tuple(ref.eval().getEmbedding(x).shape) == (2, 32)       # the stock file itself is intact


def test_star_import_exports_cover_what_the_reference_uses():
    """``from X import *`` brings in every public name: the ones scripts/model.py:21-50 and scripts/train.py use must exist."""
    for mod, names in ((CNNs, ('VGG3L', 'VGG4L', 'getVGG3LOutputDimension', 'getVGG4LOutputDimension')),
                       (poolings, ('Attention', 'MultiHeadAttention', 'DoubleMHA')),
                       (loss, ('AMSoftmax',))):
        public = getattr(mod, '__all__', [n for n in vars(mod) if not n.startswith('_')])
        for n in names:
            assert n in public, (mod.__name__, n)
