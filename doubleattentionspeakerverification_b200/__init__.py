"""B200-native (sm_100a) speaker-embedding extraction path of the DoubleMHA speaker-verification model.

Drop-in for the reference's ``scripts/CNNs.py``, ``scripts/poolings.py``, ``scripts/model.py`` and the
scoring helper of ``scripts/utils.py``; everything is computed by the hand-written CUDA kernels in
``csrc/`` through the C ABI of ``include/dasv_b200.h``.  There is no CPU or PyTorch fallback.
"""
__all__ = ['CNNs', 'poolings', 'model', 'utils', 'loss', 'ops', 'synth', 'extract']
