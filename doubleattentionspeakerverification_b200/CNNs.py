"""Drop-in for the reference's ``scripts/CNNs.py``: ``VGG3L`` / ``VGG4L`` with the same constructor,
``nn.Conv2d`` sub-modules (hence the same ``state_dict`` keys and OIHW weight shapes) and the same
``forward(x [B,T,80]) -> [B,T',C*F']`` contract (feature index ``c*F' + f``, CNNs.py:88-89).

Inference (no autograd) runs on this package's sm_100a kernels, NHWC end to end:
  * ``precision='bf16'``: conv11 direct kernel -> bf16, then tcgen05 implicit-GEMM convs with
    bias + ReLU + 2x2 ceil max-pool (+ length mask, + the final [B,T',C*F'] re-layout) fused in the
    epilogue.  Needs every tensor-core conv to have Cin % 64 == 0 (kernel_size >= 512 for VGG4L).
  * ``precision='fp32'``: CUDA-core fp32 implicit GEMM + separate pool kernel (the 1e-4 parity path).
  * ``precision='auto'`` (default): bf16 when the channel counts allow it, else fp32.
Under autograd the convolutions run through torch (cuDNN): conv backward is out of scope
(SURVEY.md §7 "Conv backward is NOT required"), only the pooling has a hand-written backward.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import ops


def _out_dim(inputDimension, n, outputChannel):
    d = np.array(inputDimension, dtype=np.float32)
    for _ in range(n):
        d = np.ceil(d / 2)
    return int(d) * outputChannel


def getVGG3LOutputDimension(inputDimension, outputChannel=128):
    """scripts/CNNs.py:7-12."""
    return _out_dim(inputDimension, 3, outputChannel)


def getVGG4LOutputDimension(inputDimension, outputChannel=128):
    """scripts/CNNs.py:14-20."""
    return _out_dim(inputDimension, 4, outputChannel)


class _VGG(nn.Module):
    _divisors = ()   # kernel_size / d = channels of each block

    def __init__(self, kernel_size, precision='auto'):
        super().__init__()
        cin = 1
        self._names = []
        for blk, d in enumerate(self._divisors, start=1):
            cout = int(kernel_size / d)
            for i in (1, 2):
                name = 'conv%d%d' % (blk, i)
                setattr(self, name, nn.Conv2d(cin, cout, 3, stride=1, padding=1))
                self._names.append(name)
                cin = cout
        self.precision = precision
        self._packed = {}

    # ---------------------------------------------------------------- weight packing cache
    def _pack(self, name, kind):
        conv = getattr(self, name)
        w = conv.weight
        key = (name, kind)
        tag = (w.data_ptr(), w._version, str(w.device))
        hit = self._packed.get(key)
        if hit is None or hit[0] != tag:
            with torch.no_grad():
                packed = ops.pack_conv_weight_bf16(w.detach()) if kind == 'bf16' else ops.pack_conv_weight_f32(w.detach())
            hit = (tag, packed)
            self._packed[key] = hit
        return hit[1]

    def resolved_precision(self):
        if self.precision in ('bf16', 'fp32'):
            return self.precision
        ok = all(getattr(self, n).in_channels % 64 == 0 and getattr(self, n).out_channels % 8 == 0 for n in self._names[1:])
        return 'bf16' if ok else 'fp32'

    # ---------------------------------------------------------------- forward
    def forward(self, paddedInputTensor, lengths=None):
        x = paddedInputTensor
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            if lengths is not None:
                raise NotImplementedError('length masking is an inference-only capability')
            return self._forward_autograd(x)
        return self._forward_kernels(x, lengths)

    def _forward_autograd(self, x):
        # training path: torch/cuDNN convolutions in the reference's NCHW layout (CNNs.py:68-91)
        h = x.view(x.size(0), x.size(1), 1, x.size(2)).transpose(1, 2)
        for i in range(0, len(self._names), 2):
            h = F.relu(getattr(self, self._names[i])(h))
            h = F.relu(getattr(self, self._names[i + 1])(h))
            h = F.max_pool2d(h, 2, stride=2, ceil_mode=True)
        h = h.transpose(1, 2)
        return h.contiguous().view(h.size(0), h.size(1), h.size(2) * h.size(3))

    @torch.no_grad()
    def _forward_kernels(self, x, lengths):
        x = x.float().contiguous()
        B = x.size(0)
        prec = self.resolved_precision()
        L = None if lengths is None else torch.as_tensor(lengths, device=x.device).to(torch.int32)
        c11 = getattr(self, self._names[0])
        nblocks = len(self._names) // 2
        if prec == 'bf16':
            # conv11 is bound by its NHWC bf16 write (2.1 GB per 256 x 4 s batch at the 3.95 TB/s pure-write bandwidth):
            # the CUDA-core kernel reaches 79 % of that; the tensor-core variant (ops.conv11_tc) measured slower.
            h = ops.conv11_direct(x, c11.weight, c11.bias, L, out_dtype=torch.bfloat16)
            for blk in range(nblocks):
                if blk > 0:
                    c = getattr(self, self._names[2 * blk])
                    h = ops.conv3x3_igemm_bf16(h, self._pack(self._names[2 * blk], 'bf16'), c.bias, c.out_channels, L)
                c = getattr(self, self._names[2 * blk + 1])
                last = blk == nblocks - 1
                # CTA pairs (cta_group::2) measured faster on the pooled layers with >= 256 input channels (conv22 +3 %,
                # conv32 +9 %, conv42 +1 %) and slower elsewhere; results are bit-identical either way
                pair = c.in_channels >= 256 and c.out_channels % 256 == 0
                h = ops.conv3x3_igemm_bf16(h, self._pack(self._names[2 * blk + 1], 'bf16'), c.bias, c.out_channels, L,
                                           pool=True, ref_layout=last, out_dtype=torch.float32, pair=pair)
                if L is not None:
                    L = (L + 1) // 2
            return h
        h = ops.conv11_direct(x, c11.weight, c11.bias, L, out_dtype=torch.float32)
        for blk in range(nblocks):
            if blk > 0:
                c = getattr(self, self._names[2 * blk])
                h = ops.conv3x3_f32(h, self._pack(self._names[2 * blk], 'f32'), c.bias, L)
            c = getattr(self, self._names[2 * blk + 1])
            h = ops.conv3x3_f32(h, self._pack(self._names[2 * blk + 1], 'f32'), c.bias, L)
            h = ops.maxpool2x2(h, ref_layout=(blk == nblocks - 1))
            if L is not None:
                L = (L + 1) // 2
        return h

    def output_lengths(self, lengths):
        """Valid frames after the front-end's ceil-mode pools (ceil(L / 2^n))."""
        L = torch.as_tensor(lengths)
        for _ in range(len(self._names) // 2):
            L = (L + 1) // 2
        return L


class VGG3L(_VGG):
    """scripts/CNNs.py:22-52."""
    _divisors = (4, 2, 1)


class VGG4L(_VGG):
    """scripts/CNNs.py:54-91."""
    _divisors = (8, 4, 2, 1)
