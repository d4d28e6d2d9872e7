"""Trial scoring (reference ``scripts/utils.py:18-21``) and the EER sweep (``scripts/train.py:135-150``) on the sm_100a kernels.

``scoreCosineDistance`` keeps the reference signature; ``score_pairs`` / ``score_matrix`` are the
batched forms the validation loop (scripts/train.py:117-133) is replaced with.
"""
import numpy as np
import torch

from . import ops


def scoreCosineDistance(emb1, emb2):
    """F.cosine_similarity(emb1, emb2, dim=-1, eps=1e-8) for ``[N,E]`` (or ``[E]``) embeddings."""
    a = emb1.reshape(-1, emb1.size(-1)).float()
    b = emb2.reshape(-1, emb2.size(-1)).float()
    if a.size(0) != b.size(0):
        a, b = torch.broadcast_tensors(a, b)
    n = a.size(0)
    idx = torch.arange(n, device=a.device, dtype=torch.int32)
    scores = ops.cosine_pairs(torch.cat([a, b], 0).contiguous(), idx, idx + n)
    return scores.view(emb1.shape[:-1]) if emb1.dim() > 1 else scores.view(())


def score_pairs(emb, idx_a, idx_b):
    """scores[i] = cos(emb[idx_a[i]], emb[idx_b[i]]) — the "uttA uttB" trial-list form."""
    return ops.cosine_pairs(emb, idx_a, idx_b)


def score_matrix(enrol, test):
    """scores[i,j] = cos(enrol[i], test[j]) — the cross-product form."""
    return ops.cosine_matrix(enrol, test)


EER_THRESHOLDS = np.arange(-1, 1, 0.01)          # scripts/train.py:137


def threshold_ge_counts(scores):
    """int64 [200]: how many of the CUDA tensor ``scores`` are >= each EER threshold (one kernel; no CPU path)."""
    return ops.threshold_counts(scores.reshape(-1).float(), EER_THRESHOLDS)


def eer_from_counts(ge_cl, n_cl, ge_im, n_im):
    """The FAR/FRR sweep of Trainer.__calculate_EER from the >=-threshold counts of the client and impostor scores."""
    FRR = np.array([round((n_cl - g) * 100 / float(n_cl), 4) for g in ge_cl])
    FAR = np.array([round(g * 100 / float(n_im), 4) for g in ge_im])
    idx = np.argwhere(np.diff(np.sign(FAR - FRR)) != 0).reshape(-1)
    if len(idx) > 0:
        return round((FAR[int(idx[0])] + FRR[int(idx[0])]) / 2, 4)
    return 50.00


def calculate_EER(CL, IM):
    """Trainer.__calculate_EER (scripts/train.py:135-150) with Score (scripts/utils.py:5-15): 200 thresholds
    np.arange(-1, 1, 0.01), FRR = % of client scores < th, FAR = % of impostor scores >= th (both rounded to 4
    decimals), EER at the first sign change of FAR - FRR, else 50.  ``CL`` / ``IM`` are CUDA score tensors; the
    2 x 200 x N comparisons run in one kernel each instead of 400 Python loops."""
    ge_cl = threshold_ge_counts(CL).cpu().numpy()
    ge_im = threshold_ge_counts(IM).cpu().numpy()
    return eer_from_counts(ge_cl, CL.numel(), ge_im, IM.numel())
