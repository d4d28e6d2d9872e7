"""GPU parity of the log-mel feature kernels (csrc/features.cu, called through the C ABI) against the numpy oracle
(oracle/feature_oracle.py) and the committed golden vectors (tests/golden/logmel_*.npz)."""
import numpy as np
import pytest
import torch

from conftest import golden
from doubleattentionspeakerverification_b200 import _lib, featureExtractor as fe, model, synth
from oracle import feature_oracle as fo

pytestmark = pytest.mark.gpu
# fp32 FFT on the device vs float64 FFT stored as complex64 in the reference: the error of a bin is relative to the
# frame's energy, not to the bin, so the bar is absolute on the log scale
TOL_LOG = 2e-3


@pytest.mark.parametrize('idx', range(3))
def test_golden_mfsc_and_cmn(idx):
    g = golden('logmel_%d.npz' % idx)
    sfr, n, seed = [int(v) for v in g['shape']]
    y = synth.make_waveform(n, sfr, seed)
    y0 = y.copy()
    mf = fe.mfsc(y, sfr)
    assert np.array_equal(y, y0)                       # unlike the reference, the caller's array is not scaled in place
    assert mf.shape == g['mfsc'].shape and mf.dtype == np.float32
    assert np.abs(mf - g['mfsc']).max() < TOL_LOG              # measured: 4e-5 max on these waveforms
    assert np.abs(mf - g['mfsc']).mean() < 2e-5
    feat, frames = fe.logmel_batch(y[None, :], [n], sfr)
    assert int(frames[0]) == g['feat'].shape[0]
    assert np.abs(feat[0].cpu().numpy() - g['feat']).max() < TOL_LOG
    assert np.abs(fe.normalize(mf.T) - g['feat']).max() < TOL_LOG


def test_ragged_batch_matches_per_utterance_oracle():
    sfr = 16000
    ns = [16000, 512, 671, 672, 40000, 300, 8123]
    N = max(ns)
    wave = np.zeros((len(ns), N), np.float32)
    ys = []
    for i, n in enumerate(ns):
        y = synth.make_waveform(n, sfr, seed=10 + i)
        ys.append(y)
        wave[i, :n] = y
        wave[i, n:] = 7.0                               # junk after the utterance must not be read into any frame
    feat, frames = fe.logmel_batch(wave, ns, sfr)
    assert frames.tolist() == [fo.num_frames(n, 160) for n in ns] == [97, 1, 1, 2, 247, 0, 48]
    feat = feat.cpu().numpy()
    for i, n in enumerate(ns):
        T = int(frames[i])
        if T:
            want = fo.extract(ys[i].astype(np.float32).astype(np.float64), sfr)
            assert np.abs(feat[i, :T] - want).max() < TOL_LOG
        assert np.all(feat[i, T:] == 0)


def test_pure_tone_lands_in_the_right_filter_and_cmn_is_zero_mean():
    sfr = 16000
    t = np.arange(sfr) / sfr
    y = 0.5 * np.sin(2 * np.pi * 2000.0 * t)
    mf = fe.mfsc(y, sfr)
    melw = fo.mel_filterbank(sfr)
    want_bin = int(np.argmax(melw[:, int(round(2000.0 / (sfr / 512)))]))
    assert np.all(np.argmax(mf, axis=0) == want_bin)
    feat, _ = fe.logmel_batch(synth.make_waveform(24000, sfr, 3)[None, :], [24000], sfr)
    assert float(feat[0].mean(0).abs().max()) < 1e-5


def test_other_sample_rate_window_and_errors():
    y = synth.make_waveform(12000, 8000, seed=5)
    mf = fe.mfsc(y, 8000, window='hann', n_mels=40)
    want = fo.mfsc(y, 8000, window='hann', n_mels=40)
    assert mf.shape == want.shape == (40, 1 + (12000 - 512) // 80)
    assert np.abs(mf - want).max() < TOL_LOG
    with pytest.raises(_lib.DasvError):
        fe.mfsc(y[:400], 8000)                           # shorter than one frame (librosa raises as well)
    with pytest.raises(_lib.DasvError):
        fe.mfsc(y, 44100)                                # 25 ms window > n_fft = 512


def test_waveform_to_embedding_pipeline():
    """getEmbedding on GPU features == getEmbedding on the oracle's features (fp32 path)."""
    cfg = synth.example_config(kernel_size=256, embedding_size=128, heads_number=8, num_spkrs=5)
    cfg.precision = 'fp32'
    sd = synth.make_state_dict(cfg, seed=7)
    net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), sd).cuda().eval()
    ns = [16000, 11000]
    wave = np.zeros((2, 16000), np.float32)
    for i, n in enumerate(ns):
        wave[i, :n] = synth.make_waveform(n, 16000, seed=20 + i)
    feat, frames = fe.logmel_batch(wave, ns, 16000)
    with torch.no_grad():
        emb = net.getEmbedding(feat, lengths=frames).cpu().numpy()
        for i, n in enumerate(ns):
            want = fo.extract(wave[i, :n].astype(np.float64), 16000).astype(np.float32)
            e1 = net.getEmbedding(torch.from_numpy(want)[None].cuda()).cpu().numpy()[0]
            assert np.abs(emb[i] - e1).max() / np.abs(e1).max() < 1e-3


def test_extract_local_audio_buckets_and_order():
    from doubleattentionspeakerverification_b200 import extract
    cfg = synth.example_config(kernel_size=256, embedding_size=128, heads_number=8, num_spkrs=5)
    cfg.precision = 'fp32'
    net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, seed=9)).cuda().eval()
    ns = [9000, 30000, 16000, 31000, 9500, 700]
    waves = [synth.make_waveform(n, 16000, seed=40 + i) for i, n in enumerate(ns)]
    embed = lambda x, L: net.getEmbedding(x, lengths=L)
    with torch.no_grad():
        got = extract.extract_local_audio(embed, waves, [5, 0, 1, 2, 3, 4], 16000, 'cuda', max_samples=70000).cpu().numpy()
        for row, i in enumerate([5, 0, 1, 2, 3, 4]):
            one = extract.extract_local_audio(embed, waves, [i], 16000, 'cuda').cpu().numpy()[0]
            assert np.abs(got[row] - one).max() / np.abs(one).max() < 1e-4, i       # batch composition does not matter
    with pytest.raises(ValueError):
        extract.extract_local_audio(embed, [waves[0][:100]], [0], 16000, 'cuda')


def test_cmvn_matches_data_py_normalisation():
    """data.py:21-30 'cmvn' (population std, floor 0.01) and 'cmn', batched with padding and as the drop-in function."""
    ns = [16000, 9000]
    wave = np.zeros((2, 16000), np.float32)
    for i, n in enumerate(ns):
        wave[i, :n] = synth.make_waveform(n, 16000, seed=60 + i)
    wave[1, :9000] *= 1e-6                                        # near-silent: log(max(1, .)) = 0 everywhere -> std = 0 -> left unscaled
    raw, frames = fe.logmel_batch(wave, ns, 16000, cmn=False)
    got, _ = fe.logmel_batch(wave, ns, 16000, cmn='cmvn')
    for i in range(2):
        T = int(frames[i])
        want = fo.normalize_features(raw[i, :T].cpu().numpy(), 'cmvn')
        assert np.abs(got[i, :T].cpu().numpy() - want).max() < 1e-4
        assert np.all(got[i, T:].cpu().numpy() == 0)
    f0 = raw[0, :int(frames[0])].cpu().numpy()
    for mode in ('cmn', 'cmvn'):
        assert np.abs(fe.normalizeFeatures(f0, mode) - fo.normalize_features(f0, mode)).max() < 1e-4


def test_get_embedding_example_from_a_wav_file(tmp_path):
    """scripts/getEmbeddingExample.py end to end: a 16-bit PCM WAV file + a checkpoint in the reference's format ->
    the embedding, equal to getEmbedding on the oracle's features of the same samples."""
    import pickle
    import wave as wavmod
    from argparse import Namespace
    from doubleattentionspeakerverification_b200 import getEmbeddingExample as ex
    y = synth.make_waveform(20000, 16000, seed=77)
    pcm = np.round(y * 32768).astype('<i2')
    wav = str(tmp_path / 'utt.wav')
    with wavmod.open(wav, 'wb') as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    cfg = synth.example_config(kernel_size=256, embedding_size=64, heads_number=8, num_spkrs=3)
    cfg.precision = 'fp32'
    net = synth.load_state_dict(model.SpeakerClassifier(cfg, 'cuda'), synth.make_state_dict(cfg, seed=3))
    ckpt = str(tmp_path / 'model.chkpt')
    torch.save({'settings': cfg, 'model': net.state_dict()}, ckpt)
    cfgp = str(tmp_path / 'config.pkl')
    with open(cfgp, 'wb') as h:
        pickle.dump(cfg, h)
    emb = ex.main(cfg, Namespace(audioPath=wav, modelConfig=cfgp, modelCheckpoint=ckpt, device='cuda')).cpu().numpy()
    feats = fo.extract(pcm.astype(np.float64) / 32768, 16000).astype(np.float32)
    with torch.no_grad():
        want = net.cuda().eval().getEmbedding(torch.from_numpy(feats)[None].cuda()).cpu().numpy()
    assert emb.shape == (1, 64)
    assert np.abs(emb - want).max() / np.abs(want).max() < 1e-3
