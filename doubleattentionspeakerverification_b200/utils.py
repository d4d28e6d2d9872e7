"""Trial scoring (reference ``scripts/utils.py:18-21``) on the sm_100a kernels.

``scoreCosineDistance`` keeps the reference signature; ``score_pairs`` / ``score_matrix`` are the
batched forms the validation loop (scripts/train.py:117-133) is replaced with.
"""
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
