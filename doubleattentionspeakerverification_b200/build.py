"""In-tree build of the C-ABI CUDA library (``libdasv_b200.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting
``.so`` sits next to this file (git-ignored, but shipped to the GPU box with the tree).

    python -m doubleattentionspeakerverification_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB = os.path.join(PKG, 'libdasv_b200.so')
ROOT = os.path.dirname(PKG)

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'), '-I', CSRC]


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path, headers_digest):
    h = hashlib.sha1(headers_digest)
    with open(path, 'rb') as f:
        h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers_digest():
    h = hashlib.sha1()
    for d in (CSRC, os.path.join(ROOT, 'include')):
        for f in sorted(os.listdir(d)):
            if f.endswith(('.cuh', '.h')):
                with open(os.path.join(d, f), 'rb') as fh:
                    h.update(fh.read())
    return h.digest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libdasv_b200.so.  Returns the library path."""
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hd = _headers_digest()
    jobs, objs = [], []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src[:-3] + '.o')
        stamp = op + '.sha1'
        dg = _digest(sp, hd)
        objs.append(op)
        if not force and os.path.exists(op) and os.path.exists(stamp) and open(stamp).read() == dg:
            continue
        jobs.append((sp, op, stamp, dg))

    def compile_one(job):
        sp, op, stamp, dg = job
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', sp, '-o', op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (sp, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, 'w') as f:
            f.write(dg)
        return op

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB) or force:
        # the CUDA runtime is linked dynamically: the process already holds one (torch's libcudart.so.12), and a static copy would
        # carry the runtime's whole entry-point table into the shipped library
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'shared']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
