"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports exactly the
symbols include/dasv_b200.h declares, and the Python modules keep the reference's state_dict
contract and refuse CPU tensors (no fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from doubleattentionspeakerverification_b200 import _lib, build, synth


@pytest.fixture(scope='module')
def library():
    build.build()
    return _lib.lib()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'dasv_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return set(re.findall(r'\b(dasv_[a-z0-9_]+)\s*\(', text))


def test_header_library_and_binding_agree(library):
    hdr = header_symbols()
    assert hdr == set(_lib.SIGNATURES.keys())
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r' T (dasv_[a-z0-9_]+)', out))
    assert hdr <= exported, hdr - exported
    assert exported - hdr == set(), 'exported but undeclared: %s' % (exported - hdr)
    assert library.dasv_abi_version() == 1


def test_library_is_sm100a_native():
    """The shipped SASS must contain the Blackwell tensor/TMA instructions (tcgen05.mma, tcgen05.ld, TMA)."""
    build.build()
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip('cuobjdump not available')
    assert 'sm_100a' in sass
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP'):
        assert mnemonic in sass, mnemonic


def test_error_reporting_without_gpu(library):
    # argument validation happens before any CUDA call, so it is testable on the CPU box
    rc = library.dasv_dmha_fwd(None, 0, None, None, None, None, None, None, None, None, None, None, 1, 1, 64, 8, None)
    assert rc != 0 and b'null' in library.dasv_last_error()
    assert library.dasv_dmha_fwd(None, 0, None, None, None, None, None, None, None, None, None, None, 0, 1, 64, 8, None) == 0   # empty batch
    rc = library.dasv_conv3x3_igemm_bf16(1, 1, 1, None, 1, 1, 1, 1, 8, 80, 60, 64, None, None)
    assert rc != 0 and b'multiple of 64' in library.dasv_last_error()
    assert library.dasv_dmha_bwd_workspace_bytes(4, 10, 256, 8) == 4 * (256 + 32) * 4
    assert library.dasv_packed_conv_weight_bf16_elems(64, 128) == 128 * 9 * 128


def test_modules_keep_reference_contract_and_refuse_cpu():
    from doubleattentionspeakerverification_b200 import model, poolings, CNNs
    cfg = synth.example_config(kernel_size=64, embedding_size=32, heads_number=8, num_spkrs=5)
    net = model.SpeakerClassifier(cfg, 'cpu')
    sd = synth.make_state_dict(cfg, 1)
    assert set(net.state_dict().keys()) == set(sd.keys())
    synth.load_state_dict(net, sd)
    assert net.vector_size == 5 * 64 // 8 and net.fc1.in_features == 40
    assert CNNs.getVGG4LOutputDimension(80, outputChannel=1024) == 5120
    assert CNNs.getVGG3LOutputDimension(80, outputChannel=1024) == 10240
    m = poolings.DoubleMHA(256, 8, mask_prob=0.3)
    assert tuple(m.utteranceAttention.query.shape) == (32, 8) and tuple(m.headsAttention.att.shape) == (32, 1)
    assert m.headsAttention.mask_prob == 3
    with pytest.raises(Exception, match='CUDA'):
        m.eval()(torch.randn(2, 5, 256))
    with pytest.raises(Exception, match='CUDA'):
        with torch.no_grad():
            net.eval().getEmbedding(torch.randn(1, 20, 80))
