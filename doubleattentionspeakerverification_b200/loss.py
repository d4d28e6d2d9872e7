"""AM-Softmax classifier head (reference ``scripts/loss.py:14-52``): ``predictionLayer`` of ``SpeakerClassifier``.

On a CUDA device the forward (both L2 normalisations, the cosine GEMM, the margin at the label -- scattered on the
device, the reference round-trips through the CPU, loss.py:45-48 -- annealing and scale) and the backward through both
normalisations run on this package's kernels (``csrc/train_tail.cu``).  There is no CPU path: a non-CUDA input raises.
"""
import torch
from torch import nn

from . import ops


class _AMSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, label, s, margin_scaled):
        ctx.set_materialize_grads(False)                  # train.py only differentiates the logits: no zero tensor for costh
        costh, logits, ix, iw = ops.amsoftmax_fwd(x, W, label, s, margin_scaled)
        ctx.save_for_backward(x, W, costh, ix, iw)
        ctx.s = s
        return costh, logits

    @staticmethod
    def backward(ctx, dcosth, dlogits):
        x, W, costh, ix, iw = ctx.saved_tensors
        dx, dW = ops.amsoftmax_bwd(dcosth, dlogits, x, W, costh, ix, iw, ctx.s)
        return dx, dW, None, None, None


class AMSoftmax(nn.Module):

    def __init__(self, in_feats, n_classes, m=0.3, s=15, annealing=False):
        super().__init__()
        self.in_feats = in_feats
        self.m = m
        self.s = s
        self.annealing = annealing
        self.W = nn.Parameter(torch.randn(in_feats, n_classes), requires_grad=True)
        nn.init.xavier_normal_(self.W, gain=1)

    def _alpha(self, step):
        return max(0, 1000. / (pow(1. + 0.0001 * float(step), 2.))) if self.annealing else 0.

    def getAnnealedFactor(self, step):
        return 1 / (1 + self._alpha(step))

    def forward(self, x, label=None, step=0):
        assert x.size(0) == label.size(0)
        assert x.size(1) == self.in_feats
        alpha = self._alpha(step)
        # s * ((costh - m*onehot) + alpha*costh) / (1 + alpha) = s*costh - s*m/(1+alpha) at the label  (loss.py:37-52)
        return _AMSoftmaxFn.apply(x.float().contiguous(), self.W, label.to(x.device), float(self.s),
                                  float(self.s) * float(self.m) / (1 + alpha))
