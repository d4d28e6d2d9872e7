"""Drop-in for the reference's ``scripts/poolings.py``: same classes, constructor signatures,
parameter names/shapes and return tuples, computed by the fused sm_100a kernels.

New capability (not in the reference): every ``forward`` takes an optional ``lengths`` (valid
frames per utterance) so a padded batch reproduces per-utterance batch-1 results.
"""
import torch
from torch import nn

from . import ops


def new_parameter(*size):
    """scripts/poolings.py:9-12 — xavier-normal initialised fp32 parameter."""
    out = nn.Parameter(torch.empty(*size, dtype=torch.float32))
    nn.init.xavier_normal_(out)
    return out


class _DoubleMHAFn(torch.autograd.Function):
    """Fused DoubleMHA (att given) / MultiHeadAttention (att None) with the closed-form backward
    of SURVEY.md §3.4.  The alignment output is not differentiable (train.py discards it,
    scripts/model.py:65)."""

    @staticmethod
    def forward(ctx, x, query, att, lengths, keep):
        r = ops.dmha_fwd(x, query, att, lengths=lengths, keep=keep, need_align=True)
        ctx.has_head = att is not None
        ctx.lengths = lengths
        ctx.save_for_backward(x, query, att if att is not None else query.new_empty(0),
                              r['ctx'], r['lse'], r['headw'] if att is not None else query.new_empty(0))
        first = r['out'] if att is not None else r['ctx']
        third = r['headw'] if att is not None else r['lse']
        ctx.mark_non_differentiable(r['align'], third)
        return first, r['align'], third

    @staticmethod
    def backward(ctx, g_first, _g_align, _g_third):
        x, query, att, cvec, lse, headw = ctx.saved_tensors
        g_first = g_first.contiguous().float()
        if ctx.has_head:
            dx, dq, da = ops.dmha_bwd(x, query, att, g_first, None, cvec, lse, headw, lengths=ctx.lengths)
            return dx, dq, da.view_as(att), None, None
        dx, dq, _ = ops.dmha_bwd(x, query, None, None, g_first.view_as(cvec), cvec, lse, None, lengths=ctx.lengths)
        return dx, dq, None, None, None


def innerKeyValueAttention(query, key, value, lengths=None):
    """scripts/poolings.py:73-80.  key ``[B*T,H,dh]`` and value ``[B,T,H,dh]`` are views of the same
    frames; returns ``(ct [B,H,dh], p_attn [B,T,H])``.  Scale is 1/sqrt(query.size(-1)) = 1/sqrt(H)."""
    B, T, H, dh = value.shape
    ct, p_attn, _ = _DoubleMHAFn.apply(value.reshape(B, T, H * dh), query, None, lengths, None)
    return ct, p_attn


class Attention(nn.Module):
    """scripts/poolings.py:14-27 — one learned query over time, no scale."""

    def __init__(self, embedding_size):
        super().__init__()
        self.embedding_size = embedding_size
        self.att = new_parameter(self.embedding_size, 1)

    def forward(self, ht, lengths=None):
        if torch.is_grad_enabled() and (ht.requires_grad or self.att.requires_grad):
            # training path: stock torch ops + autograd (like the conv stack; only DoubleMHA / MHA have hand-written
            # backward kernels).  Same math as the reference, without its batch-collapsing squeeze().
            if lengths is not None:
                raise NotImplementedError('length masking is an inference-only capability')
            score = torch.softmax(torch.matmul(ht, self.att).squeeze(-1), dim=-1).unsqueeze(-1)
            return torch.sum(ht * score, dim=1), score
        ct, p = ops.attention_fwd(ht, self.att, lengths=lengths)
        return ct, p.view(ht.size(0), ht.size(1), 1)


class HeadAttention(nn.Module):
    """scripts/poolings.py:29-71 (narrow path; ``attentionSmoothing=True`` is dead code in the
    reference — it raises there and is rejected here)."""

    def __init__(self, encoder_size, heads_number, mask_prob=0.25, attentionSmoothing=False):
        super().__init__()
        if attentionSmoothing:
            raise NotImplementedError('attentionSmoothing=True is broken in the reference (poolings.py:53-59)')
        self.embedding_size = encoder_size // heads_number
        self.att = new_parameter(self.embedding_size, 1)
        self.mask_prob = int(1 / mask_prob)
        self.attentionSmoothing = attentionSmoothing

    def draw_keep_mask(self, batch, heads, device):
        """The reference's head drop-out draw (poolings.py:41): keep where random_(n) > 0."""
        return torch.randint(0, self.mask_prob, (batch, heads), device=device) > 0

    def forward(self, ht, keep=None):
        if self.training and keep is None:
            keep = self.draw_keep_mask(ht.size(0), ht.size(1), ht.device)
        if torch.is_grad_enabled() and (ht.requires_grad or self.att.requires_grad):
            # stand-alone use under autograd: stock torch ops (DoubleMHA fuses this stage with its own backward kernel)
            score = torch.matmul(ht, self.att).squeeze(-1)
            if keep is not None:
                score = score.masked_fill(~keep.bool(), float('-inf'))
            score = torch.softmax(score, dim=-1).unsqueeze(-1)
            return torch.sum(ht * score, dim=1), score
        ct, w = ops.attention_fwd(ht, self.att, keep=keep)
        return ct, w.view(ht.size(0), ht.size(1), 1)


class MultiHeadAttention(nn.Module):
    """scripts/poolings.py:83-109."""

    def __init__(self, encoder_size, heads_number):
        super().__init__()
        self.encoder_size = encoder_size
        assert self.encoder_size % heads_number == 0
        self.head_size = self.encoder_size // heads_number
        self.heads_number = heads_number
        self.query = new_parameter(self.head_size, self.heads_number)
        self.alignment = None

    def _run(self, ht, lengths=None):
        ct, self.alignment, _ = _DoubleMHAFn.apply(ht, self.query, None, lengths, None)
        return ct

    def getAlignments(self, ht, lengths=None):
        self._run(ht, lengths)
        return self.alignment

    def getHeadsContextVectors(self, ht, lengths=None):
        return self._run(ht, lengths)

    def forward(self, ht, lengths=None):
        ct = self._run(ht, lengths)
        return ct.reshape(ct.size(0), -1), self.alignment


class DoubleMHA(nn.Module):
    """scripts/poolings.py:112-129 — both attentions in one kernel launch."""

    def __init__(self, encoder_size, heads_number, mask_prob=0.2):
        super().__init__()
        self.heads_number = heads_number
        self.utteranceAttention = MultiHeadAttention(encoder_size, heads_number)
        self.heads_size = encoder_size // heads_number
        self.headsAttention = HeadAttention(encoder_size, heads_number, mask_prob=mask_prob, attentionSmoothing=False)

    def _run(self, x, lengths=None, keep=None):
        if self.training and keep is None:
            keep = self.headsAttention.draw_keep_mask(x.size(0), self.heads_number, x.device)
        out, align, headw = _DoubleMHAFn.apply(x, self.utteranceAttention.query, self.headsAttention.att, lengths, keep)
        self.utteranceAttention.alignment = align
        return out, align, headw

    def getAlignments(self, x, lengths=None):
        _, align, headw = self._run(x, lengths)
        return align, headw.view(x.size(0), self.heads_number, 1)

    def pooled(self, x, lengths=None, keep=None):
        """The pooled vector alone, for callers that discard the alignment (SpeakerClassifier.getEmbedding,
        scripts/model.py:55): without autograd the kernel then skips the [B,T,H] alignment store altogether."""
        if torch.is_grad_enabled() and (x.requires_grad or self.utteranceAttention.query.requires_grad):
            return self._run(x, lengths, keep)[0]
        if self.training and keep is None:
            keep = self.headsAttention.draw_keep_mask(x.size(0), self.heads_number, x.device)
        return ops.dmha_fwd(x, self.utteranceAttention.query, self.headsAttention.att, lengths=lengths, keep=keep,
                            need_align=False)['out']

    def forward(self, x, lengths=None, keep=None):
        """``keep`` ([B,H] bool) injects the training-mode head drop-out draw (for reproducible tests)."""
        out, align, _ = self._run(x, lengths, keep)
        return out, align
