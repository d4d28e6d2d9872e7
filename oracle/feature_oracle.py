"""numpy restatement of the reference's feature extraction (the checker for csrc/features.cu).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference (``scripts/featureExtractor.py:8-26``) delegates the arithmetic to **librosa 0.7.2**
(``requirements.txt:3``), which is NOT in the reference tree and NOT installed in this image, so the
library's published algorithm is restated here function by function:

* ``librosa.core.stft(y, n_fft, hop_length, win_length, window, center=False)``: the window comes from
  ``scipy.signal.get_window(window, win_length, fftbins=True)`` (scipy IS installed: same call), is zero-padded
  to ``n_fft`` around its centre (``util.pad_center``), frames are ``y[t*hop : t*hop + n_fft]`` for
  ``t < 1 + (len(y) - n_fft) // hop`` (``util.frame``), the transform is ``numpy.fft.rfft`` of window * frame,
  stored as complex64.
* ``librosa.feature.melspectrogram(S=D, sr, n_mels, fmin, fmax, norm=None)``: with ``S`` given the input is used
  as is (MAGNITUDE here, not power) and multiplied by ``filters.mel(sr, n_fft=2*(rows-1), n_mels, fmin, fmax,
  htk=False, norm=None)``: triangular filters on the Slaney mel scale (linear below 1 kHz, 200/3 Hz per mel;
  logarithmic above with step log(6.4)/27), peak 1, float32.

Pin: **parity unpinned by the reference itself** (it has no tests and librosa cannot be imported here).  The
restatement is pinned instead against an independent implementation of the same published algorithm,
``transformers.audio_utils`` (``mel_filter_bank(..., mel_scale='slaney', norm=None)`` and ``spectrogram(...,
center=False, power=1.0)``, written to reproduce librosa), by ``oracle/make_golden_features.py`` ->
``tests/golden/logmel_*.npz`` -> ``tests/test_oracle_golden.py``.
"""
import numpy as np

N_FFT = 512      # scripts/featureExtractor.py:11


def hz_to_mel(f):
    """librosa.core.hz_to_mel(htk=False): Slaney's Auditory Toolbox scale."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(m):
    """librosa.core.mel_to_hz(htk=False)."""
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft=N_FFT, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm=None) -> float32 [n_mels, 1 + n_fft//2]."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    return w.astype(np.float32)


def fft_window(window, win_length, n_fft=N_FFT):
    """scipy.signal.get_window(window, win_length, fftbins=True) padded to n_fft (librosa.util.pad_center)."""
    import scipy.signal
    w = scipy.signal.get_window(window, win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    return np.pad(w, (lpad, n_fft - win_length - lpad), mode='constant')


def preemphasis(y, coef=0.97):
    """scripts/featureExtractor.py:16-18 (without mutating the caller's array)."""
    y = np.asarray(y, dtype=np.float64) * 32768
    out = y.copy()
    out[1:] = y[1:] - coef * y[:-1]
    out[0] = y[0] * (1 - coef)
    return out


def num_frames(n_samples, hop, n_fft=N_FFT):
    return 1 + (n_samples - n_fft) // hop if n_samples >= n_fft else 0


def mfsc(y, sfr, window_size=0.025, window_stride=0.010, window='hamming', n_mels=80, preemCoef=0.97):
    """scripts/featureExtractor.py:8-23: float32 ``[n_mels, T]`` log mel-filterbank energies."""
    win_length = int(sfr * window_size)
    hop = int(sfr * window_stride)
    y = preemphasis(y, preemCoef)
    T = num_frames(len(y), hop)
    if T <= 0:
        raise ValueError('input of %d samples is shorter than one frame of %d' % (len(y), N_FFT))
    win = fft_window(window, win_length)
    idx = np.arange(N_FFT)[:, None] + hop * np.arange(T)[None, :]
    S = np.fft.rfft(win[:, None] * y[idx], axis=0).astype(np.complex64)       # librosa.stft dtype
    D = np.abs(S)
    param = mel_filterbank(sfr, N_FFT, n_mels, 0.0, sfr / 2.0).dot(D)          # float32 . float32
    return np.log(np.maximum(1, param))


def normalize(features):
    """scripts/featureExtractor.py:25-26 (cepstral mean normalisation over time, features ``[T, n_mels]``)."""
    return features - np.mean(features, axis=0)


def normalize_features(features, normalization='cmn'):
    """scripts/data.py:21-30 (``normalizeFeatures``; ``Dataset.__normalize`` :47-54 is the same), without mutating the input."""
    features = np.array(features, copy=True)
    mean = np.mean(features, axis=0)
    features -= mean
    if normalization == 'cmn':
        return features
    if normalization == 'cmvn':
        std = np.std(features, axis=0)
        std = np.where(std > 0.01, std, 1.0)
        return features / std
    raise ValueError(normalization)


def extract(y, sfr, **kw):
    """scripts/featureExtractor.py:29-33 minus the file read: ``[T, n_mels]`` CMN'd features."""
    return normalize(np.transpose(mfsc(y, sfr, **kw)))


def mel_ranges(melw):
    """First and one-past-last bin with a non-zero weight for every filter (the kernel's sparse loop bounds)."""
    r = np.zeros((melw.shape[0], 2), dtype=np.int32)
    for m in range(melw.shape[0]):
        nz = np.nonzero(melw[m])[0]
        if nz.size:
            r[m] = (nz[0], nz[-1] + 1)
    return r
