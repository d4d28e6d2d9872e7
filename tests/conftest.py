import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import time
        import torch
        has_gpu = torch.cuda.is_available()
        # a device node without a working CUDA context is a box that is still coming up: wait for it rather than skip
        tries = 0
        while not has_gpu and os.path.exists('/dev/nvidiactl') and tries < 10:
            time.sleep(3.0)
            has_gpu = torch.cuda.is_available()
            tries += 1
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b):
    """max |a-b| / max|b| — the 'relative error' used for the fp32 1e-4 bar."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def min_cosine(a, b):
    a = np.asarray(a, np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, np.float64).reshape(b.shape[0], -1)
    c = np.sum(a * b, -1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))
    return float(c.min())


def report(name, **values):
    """Append measured parity values to gpurun_out/parity_values.txt (scratch; summarised under profiles/ by hand), so the
    bars written in the tests can be stated next to what was actually measured."""
    try:
        d = os.path.join(ROOT, 'gpurun_out')
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, 'parity_values.txt'), 'a') as f:
            f.write(name + ' ' + ' '.join('%s=%.6g' % kv for kv in sorted(values.items())) + '\n')
    except OSError:
        pass
