"""GPU drop-in for the reference's ``scripts/featureExtractor.py`` (log mel-filterbank features + CMN).

Same function names and argument meaning (``mfsc``, ``normalize``, ``extractFeatures``); the arithmetic runs in
``csrc/features.cu`` through the C ABI (``dasv_logmel_f32`` / ``dasv_cmn_f32``).  There is no CPU path: the host
side only builds the two small tables the kernel needs (window taps and the mel filterbank, both functions of the
sample rate) and moves the waveform to the device.  ``logmel_batch`` is the batched entry the extraction pipeline
uses: padded waveforms in, ``[B, T, n_mels]`` features + frame counts out, ready for ``VGG4L``.

Differences from the reference that a caller can observe: ``mfsc`` does not scale the caller's array in place
(featureExtractor.py:16 does ``y *= 32768``), and the arithmetic is fp32 on the device (librosa computes the FFT in
float64 and stores complex64): features agree to ~1e-5 absolute on the log scale.
"""
import functools
import math

import numpy as np
import torch

from . import _lib, ops

N_FFT = 512          # scripts/featureExtractor.py:11


def _slaney_hz(m):
    m = np.asarray(m, np.float64)
    lin = m * (200.0 / 3)
    return np.where(m >= 15.0, 1000.0 * np.exp((math.log(6.4) / 27.0) * (m - 15.0)), lin)


def _slaney_mel(f):
    return f / (200.0 / 3) if f < 1000.0 else 15.0 + math.log(f / 1000.0) / (math.log(6.4) / 27.0)


@functools.lru_cache(maxsize=16)
def _tables(sfr, win_length, window, n_mels):
    """(window taps [win_length] f32, mel weights [n_mels, 257] f32, non-zero bin range per filter [n_mels, 2] i32).

    Window: ``scipy.signal.get_window(window, win_length, fftbins=True)`` as librosa.stft does.  Mel weights:
    librosa.filters.mel(htk=False, norm=None) = triangles of peak 1 whose corners are equally spaced on the Slaney mel
    scale between 0 and sfr/2 (scripts/featureExtractor.py:12-13,21)."""
    import scipy.signal
    taps = scipy.signal.get_window(window, win_length, fftbins=True).astype(np.float32)
    corners = _slaney_hz(np.linspace(_slaney_mel(0.0), _slaney_mel(sfr / 2.0), n_mels + 2))
    bins = np.linspace(0.0, sfr / 2.0, N_FFT // 2 + 1)
    left, centre, right = corners[:-2, None], corners[1:-1, None], corners[2:, None]
    up = (bins[None, :] - left) / (centre - left)
    down = (right - bins[None, :]) / (right - centre)
    melw = np.maximum(0.0, np.minimum(up, down)).astype(np.float32)
    rng = np.zeros((n_mels, 2), np.int32)
    for m in range(n_mels):
        nz = np.flatnonzero(melw[m])
        if nz.size:
            rng[m] = (nz[0], nz[-1] + 1)
    return taps, melw, rng


_dev_tables = {}


def _device_tables(device, sfr, win_length, window, n_mels):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), sfr, win_length, window, n_mels)
    t = _dev_tables.get(key)
    if t is None:
        t = tuple(torch.from_numpy(a).to(device) for a in _tables(sfr, win_length, window, n_mels))
        _dev_tables[key] = t
    return t


def frames_for(n_samples, hop):
    """Frames librosa.stft(center=False) produces for ``n_samples`` samples (0 if shorter than one n_fft frame)."""
    n_samples = np.asarray(n_samples)
    return np.where(n_samples >= N_FFT, 1 + (n_samples - N_FFT) // hop, 0)


def logmel_batch(wave, n_samples, sfr, window_size=0.025, window_stride=0.010, window='hamming', n_mels=80,
                 preemCoef=0.97, cmn=True, device=None):
    """Batched ``normalize(mfsc(y).T)``.  wave ``[B, N]`` float (numpy or torch, zero-padded), n_samples ``[B]``.

    ``cmn``: True / 'cmn' = cepstral mean normalisation, 'cmvn' = mean and variance (``data.py:21-30``), False = none.
    Returns (features ``[B, Tmax, n_mels]`` float32 CUDA tensor, frames ``[B]`` int32 CUDA tensor); rows past an
    utterance's frame count are zero (the masked front-end ignores them)."""
    win_length, hop = int(sfr * window_size), int(sfr * window_stride)
    if win_length > N_FFT:
        raise _lib.DasvError('window of %d samples exceeds n_fft = %d' % (win_length, N_FFT))
    if device is None:
        device = wave.device if isinstance(wave, torch.Tensor) and wave.is_cuda else torch.device('cuda', torch.cuda.current_device())
    wave_d = torch.as_tensor(wave).to(device=device, dtype=torch.float32, non_blocking=True)
    if wave_d.dim() != 2:
        raise _lib.DasvError('wave must be [B, N]')
    n_host = np.asarray(n_samples.cpu() if isinstance(n_samples, torch.Tensor) else n_samples).astype(np.int64)
    if n_host.shape != (wave_d.shape[0],) or (n_host > wave_d.shape[1]).any():
        raise _lib.DasvError('n_samples must be [B] and <= wave.shape[1]')
    frames = frames_for(n_host, hop).astype(np.int32)
    taps, melw, rng = _device_tables(torch.device(device), int(sfr), win_length, window, int(n_mels))
    counts = torch.from_numpy(np.stack([n_host.astype(np.int32), frames])).to(device, non_blocking=True)   # one small H2D copy
    feat = ops.logmel(wave_d, counts[0], counts[1], int(frames.max()) if frames.size else 0, taps, hop, melw, rng,
                      float(preemCoef), 32768.0, cmn)
    return feat, counts[1]


def mfsc(y, sfr, window_size=0.025, window_stride=0.010, window='hamming', n_mels=80, preemCoef=0.97):
    """scripts/featureExtractor.py:8-23: float32 numpy ``[n_mels, T]`` log mel-filterbank energies of one waveform."""
    y = np.asarray(y)
    if y.ndim != 1:
        raise _lib.DasvError('mfsc expects a mono waveform')
    if len(y) < N_FFT:
        raise _lib.DasvError('input of %d samples is shorter than one frame of %d' % (len(y), N_FFT))   # librosa.util.frame raises too
    feat, _ = logmel_batch(y[None, :], [len(y)], sfr, window_size, window_stride, window, n_mels, preemCoef, cmn=False)
    return feat[0].t().contiguous().cpu().numpy()


def normalize(features):
    """scripts/featureExtractor.py:25-26."""
    return features - np.mean(features, axis=0)


def normalizeFeatures(features, normalization='cmn'):
    """scripts/data.py:21-30 on the GPU: ``[T, n_mels]`` numpy features -> 'cmn' or 'cmvn' normalised copy."""
    if normalization not in ('cmn', 'cmvn'):
        raise _lib.DasvError("normalization must be 'cmn' or 'cmvn'")
    f = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32)).cuda()[None].contiguous()
    frames = torch.tensor([f.shape[1]], dtype=torch.int32, device=f.device)
    with torch.cuda.device(f.device):
        rc = _lib.lib().dasv_cmn_f32(f.data_ptr(), frames.data_ptr(), 1, f.shape[1], f.shape[2], 1 if normalization == 'cmvn' else 0,
                                     torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, 'dasv_cmn_f32')
    return f[0].cpu().numpy()


def read_wav(path):
    """(float64 samples in [-1, 1), sample rate) like ``soundfile.read``; falls back to the standard library's
    ``wave`` module (16/32-bit PCM) when soundfile is not installed."""
    try:
        import soundfile as sf
        return sf.read(path)
    except ImportError:
        import wave
        with wave.open(path, 'rb') as w:
            width, ch, sfr, n = w.getsampwidth(), w.getnchannels(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
        if width not in (2, 4):
            raise _lib.DasvError('%s: only 16- and 32-bit PCM WAV can be read without soundfile' % path)
        a = np.frombuffer(raw, dtype='<i2' if width == 2 else '<i4').astype(np.float64) / float(1 << (8 * width - 1))
        return (a.reshape(-1, ch) if ch > 1 else a), sfr


def extractFeatures(audioPath):
    """scripts/featureExtractor.py:29-33: ``[T, 80]`` CMN'd features of one audio file."""
    y, sfreq = read_wav(audioPath)
    feat, _ = logmel_batch(np.asarray(y)[None, :], [len(y)], sfreq)
    return feat[0].cpu().numpy()
