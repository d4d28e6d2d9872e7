"""AM-Softmax classifier head (reference ``scripts/loss.py:14-52``).

Training-only and OUT OF SCOPE for the extraction path: kept in stock
PyTorch so ``SpeakerClassifier`` has the reference's ``predictionLayer.W`` parameter and
``train.py``-style steps run.  The only change is that the margin is scattered on the label's own
device (the reference round-trips through the CPU, loss.py:45-48).
"""
import torch
from torch import nn


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
        xn = x / torch.norm(x, p=2, dim=1, keepdim=True).clamp(min=1e-12)
        wn = self.W / torch.norm(self.W, p=2, dim=0, keepdim=True).clamp(min=1e-12)
        costh = torch.mm(xn, wn)
        margin = torch.zeros_like(costh).scatter_(1, label.view(-1, 1).to(costh.device), self.m)
        alpha = self._alpha(step)
        combined = ((costh - margin) + alpha * costh) / (1 + alpha)
        return costh, self.s * combined
