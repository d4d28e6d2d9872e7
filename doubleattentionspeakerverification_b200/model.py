"""Drop-in for the reference's ``scripts/model.py``: ``SpeakerClassifier(parameters, device)`` with the
same sub-module names (``front_end``, ``poolingLayer``, ``fc1``, ``b1``, ``fc2``, ``b2``, ``preLayer``,
``b3``, ``predictionLayer``), hence the same ``state_dict`` keys and shapes, and the same
``getEmbedding`` / ``forward`` contracts (scripts/model.py:52-71).
"""
import torch
from torch import nn
from torch.nn import functional as F

from . import ops
from .CNNs import VGG3L, VGG4L, getVGG3LOutputDimension, getVGG4LOutputDimension
from .loss import AMSoftmax
from .poolings import Attention, DoubleMHA, MultiHeadAttention


class SpeakerClassifier(nn.Module):

    def __init__(self, parameters, device):
        super().__init__()
        parameters.feature_size = 80                                  # scripts/model.py:13
        self.device = device
        self._init_front_end(parameters)
        self._init_pooling(parameters)
        self._init_fc(parameters)
        self.predictionLayer = AMSoftmax(parameters.embedding_size, parameters.num_spkrs, s=parameters.scalingFactor,
                                         m=parameters.marginFactor, annealing=parameters.annealing)
        self._tail_cache = None

    def _init_front_end(self, parameters):                            # scripts/model.py:21-29
        precision = getattr(parameters, 'precision', 'auto')
        tk = bool(getattr(parameters, 'train_kernels', False))       # training on this package's conv kernels (bf16)
        if parameters.front_end == 'VGG3L':
            self.vector_size = getVGG3LOutputDimension(parameters.feature_size, outputChannel=parameters.kernel_size)
            self.front_end = VGG3L(parameters.kernel_size, precision=precision, train_kernels=tk)
        if parameters.front_end == 'VGG4L':
            self.vector_size = getVGG4LOutputDimension(parameters.feature_size, outputChannel=parameters.kernel_size)
            self.front_end = VGG4L(parameters.kernel_size, precision=precision, train_kernels=tk)

    def _init_pooling(self, parameters):                              # scripts/model.py:31-41
        self.pooling_method = parameters.pooling_method
        if self.pooling_method == 'Attention':
            self.poolingLayer = Attention(self.vector_size)
        elif self.pooling_method == 'MHA':
            self.poolingLayer = MultiHeadAttention(self.vector_size, parameters.heads_number)
        elif self.pooling_method == 'DoubleMHA':
            self.poolingLayer = DoubleMHA(self.vector_size, parameters.heads_number, mask_prob=parameters.mask_prob)
            self.vector_size = self.vector_size // parameters.heads_number

    def _init_fc(self, parameters):                                   # scripts/model.py:43-50
        E = parameters.embedding_size
        self.fc1 = nn.Linear(self.vector_size, E)
        self.b1 = nn.BatchNorm1d(E)
        self.fc2 = nn.Linear(E, E)
        self.b2 = nn.BatchNorm1d(E)
        self.preLayer = nn.Linear(E, E)
        self.b3 = nn.BatchNorm1d(E)

    # ------------------------------------------------------------------ fused eval-mode tail
    def _tail_params(self):
        """fc1/fc2 transposed and b2 folded to scale/shift (eval mode), cached per parameter version."""
        ps = (self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.b2.weight, self.b2.bias,
              self.b2.running_mean, self.b2.running_var)
        tag = tuple((p.data_ptr(), p._version) for p in ps)
        if self._tail_cache is None or self._tail_cache[0] != tag:
            with torch.no_grad():
                scale = (self.b2.weight / torch.sqrt(self.b2.running_var + self.b2.eps)).float().contiguous()
                shift = (self.b2.bias - self.b2.running_mean * scale).float().contiguous()
                packed = (self.fc1.weight.t().contiguous().float(), self.fc1.bias.float().contiguous(),
                          self.fc2.weight.t().contiguous().float(), self.fc2.bias.float().contiguous(), scale, shift)
            self._tail_cache = (tag, packed)
        return self._tail_cache[1]

    def _tail(self, embedding0):
        fused = (not self.training) and not (torch.is_grad_enabled() and
                                             (embedding0.requires_grad or self.fc1.weight.requires_grad))
        if fused:
            return ops.fc_tail(embedding0.float().contiguous(), *self._tail_params())
        embedding1 = F.relu(self.fc1(embedding0))
        return self.b2(F.relu(self.fc2(embedding1)))                  # batch statistics / autograd: stock torch

    def getEmbedding(self, x, lengths=None):
        """scripts/model.py:52-59.  ``lengths`` (valid input frames per utterance) enables padded batches."""
        out_len = None
        if lengths is None:
            encoder_output = self.front_end(x)
        else:
            encoder_output = self.front_end(x, lengths=lengths)
            out_len = self.front_end.output_lengths(torch.as_tensor(lengths))
        if isinstance(self.poolingLayer, DoubleMHA):
            embedding0 = self.poolingLayer.pooled(encoder_output, lengths=out_len)     # the alignment is discarded here (model.py:55)
        elif out_len is None:
            embedding0, alignment = self.poolingLayer(encoder_output)
        else:
            embedding0, alignment = self.poolingLayer(encoder_output, lengths=out_len)
        return self._tail(embedding0)

    def forward(self, x, label=None, step=0):
        """scripts/model.py:61-71."""
        encoder_output = self.front_end(x)
        embedding0, alignment = self.poolingLayer(encoder_output)
        embedding2 = self._tail(embedding0)
        embedding3 = self.preLayer(embedding2)
        prediction, ouputTensor = self.predictionLayer(embedding3, label, step)
        return prediction, ouputTensor
