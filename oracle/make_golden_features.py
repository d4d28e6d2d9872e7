"""Golden vectors for the feature-extraction oracle (``oracle/feature_oracle.py``).

librosa (the library the reference calls, scripts/featureExtractor.py:19-21) is not installed here, so the
vectors come from an INDEPENDENT implementation of the same published algorithm: ``transformers.audio_utils``
(``mel_filter_bank(mel_scale='slaney', norm=None)``, ``window_function('hamming', periodic)``,
``spectrogram(center=False, power=1.0)``), run on seeded synthetic waveforms.  The reference's own steps
around the library calls (x32768, whole-signal pre-emphasis, log(max(1, .)), CMN) are applied here literally.

    python oracle/make_golden_features.py      # writes tests/golden/logmel_*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubleattentionspeakerverification_b200 import synth  # noqa: E402


def reference_like(y, sfr, n_mels=80, coef=0.97):
    from transformers import audio_utils as au
    win_length, hop = int(sfr * 0.025), int(sfr * 0.010)
    y = np.asarray(y, np.float64) * 32768                       # featureExtractor.py:16
    e = y.copy()
    e[1:] = y[1:] - coef * y[:-1]                               # :17
    e[0] = y[0] * (1 - coef)                                    # :18
    window = au.window_function(win_length, 'hamming', periodic=True, frame_length=512, center=True)
    mag = au.spectrogram(e, window, frame_length=512, hop_length=hop, fft_length=512, power=1.0, center=False,
                         dtype=np.float64)                      # [257, T]  (|stft|, :19-20)
    melw = au.mel_filter_bank(257, n_mels, 0.0, sfr / 2.0, sfr, norm=None, mel_scale='slaney')   # [257, n_mels]
    param = melw.T.astype(np.float32).dot(mag.astype(np.float32))                                # :21
    mf = np.log(np.maximum(1, param))                           # :22
    feat = mf.T
    return mf, feat - feat.mean(axis=0), melw.T                 # :25-26


def main():
    out = os.path.join(ROOT, 'tests', 'golden')
    for i, (sfr, seconds, seed) in enumerate([(16000, 1.0, 0), (16000, 2.37, 1), (8000, 1.5, 2)]):
        y = synth.make_waveform(int(sfr * seconds), sfr, seed)
        mf, feat, melw = reference_like(y, sfr)
        np.savez_compressed(os.path.join(out, 'logmel_%d.npz' % i), shape=np.array([sfr, len(y), seed]),
                            mfsc=mf.astype(np.float32), feat=feat.astype(np.float32), melw=melw.astype(np.float32))
        print(i, sfr, len(y), mf.shape, float(mf.min()), float(mf.max()))


if __name__ == '__main__':
    main()
