"""Drop-in for the reference's ``scripts/CNNs.py``: ``VGG3L`` / ``VGG4L`` with the same constructor,
``nn.Conv2d`` sub-modules (hence the same ``state_dict`` keys and OIHW weight shapes) and the same
``forward(x [B,T,80]) -> [B,T',C*F']`` contract (feature index ``c*F' + f``, CNNs.py:88-89).

Inference (no autograd) runs on this package's sm_100a kernels, NHWC end to end:
  * ``precision='bf16'``: conv11 direct kernel -> bf16, then tcgen05 implicit-GEMM convs with
    bias + ReLU + 2x2 ceil max-pool (+ length mask, + the final [B,T',C*F'] re-layout) fused in the
    epilogue.  Needs every tensor-core conv to have Cin % 64 == 0 (kernel_size >= 512 for VGG4L).
  * ``precision='fp16'``: the same kernels with fp16 activations and weights (tcgen05 kind::f16 takes either 16-bit
    format, but both operands must have the same one: a mixed pair is an illegal instruction on B200): three more
    mantissa bits than bf16 at the same speed (embedding cosine 0.99991 -> 0.99999, trial-score error 1e-3 -> 2e-4 on the
    exampleModel config), for models whose activations stay below 65504; stores saturate instead of overflowing.
  * ``precision='fp32x3'``: fp32 parity ON the tensor cores: every fp32 activation and weight travels as two bf16 numbers
    (hi = bf16(v), lo = bf16(v - hi)) and a product is three MMAs, hi*hi + hi*lo + lo*hi, accumulated in fp32 (error ~2^-16
    per product; embeddings within 1e-4 of the reference, measured ~1e-5) -- the same kernels at a third of the bf16 rate.
  * ``precision='fp32'``: CUDA-core fp32 implicit GEMM + separate pool kernel (plain fp32 FMAs; the slowest path).
  * ``precision='auto'`` (default): bf16 when the channel counts allow it, else fp32.
Under autograd the default is torch (cuDNN) convolutions in fp32, numerically the reference's own training path.
``train_kernels=True`` (with bf16 precision and channel counts that are multiples of 64) switches training to this
package's kernels as well: the forward keeps every layer's ReLU output, the backward is tcgen05 implicit GEMMs for the
input gradients (the forward kernel with rotated weights and a linear epilogue) and the weight gradients
(``csrc/conv_wgrad.cu``) plus streaming kernels for the ReLU / max-pool backward and the bias gradients.
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

    def __init__(self, kernel_size, precision='auto', train_kernels=False):
        super().__init__()
        self.train_kernels = train_kernels
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
        # conv11 inside conv12's kernel (ops.conv12_fused).  Off by default: it removes the 2 x 2.1 GB round trip of conv11's
        # output (ncu: DRAM read of the layer 2.1 GB -> 33 MB) but its FMA warps take issue slots from the MMA and epilogue
        # warps (tensor pipe 71 % -> 54 % active): 2.38 ms against 0.77 + 1.78 ms for the two kernels, and no gain per step
        # under the power cap (profiles/r2_fused_conv12_summary.txt).
        self.fuse_first = False
        self.use_pairs = False           # pooled layers on CTA pairs (cta_group::2); see _forward_kernels
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
                if kind == 'f32':
                    packed = ops.pack_conv_weight_f32(w.detach())
                elif kind == 'x3':
                    packed = ops.pack_conv_weight_x3(w.detach())
                else:
                    packed = ops.pack_conv_weight_bf16(w.detach(), torch.float16 if kind == 'f16' else torch.bfloat16)
            hit = (tag, packed)
            self._packed[key] = hit
        return hit[1]

    def resolved_precision(self):
        if self.precision in ('bf16', 'fp16', 'fp32', 'fp32x3'):
            return self.precision
        ok = all(getattr(self, n).in_channels % 64 == 0 and getattr(self, n).out_channels % 8 == 0 for n in self._names[1:])
        return 'bf16' if ok else 'fp32'

    def _precision_for(self, feature_size):
        """The precision one call runs in: the tensor-core kernels pool in the epilogue and need an even number of bins
        in front of every pool, so 'auto' falls back to fp32 for other feature sizes (80 -> 40 -> 20 -> 10 is fine)."""
        prec = self.resolved_precision()
        f, even = int(feature_size), True
        for _ in range(len(self._names) // 2):
            even = even and f % 2 == 0
            f = (f + 1) // 2
        if prec in ('bf16', 'fp16', 'fp32x3') and not even:
            if self.precision in ('bf16', 'fp16', 'fp32x3'):
                raise ValueError('the tensor-core path needs an even number of frequency bins in front of every pool (got %d)' % feature_size)
            return 'fp32'
        return prec

    # ---------------------------------------------------------------- forward
    def forward(self, paddedInputTensor, lengths=None):
        x = paddedInputTensor
        # decided from the tensors actually used: nn.DataParallel replicas have no registered parameters
        # (self.parameters() is empty there), but their conv sub-modules carry the broadcast weights
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(
            getattr(self, n).weight.requires_grad or getattr(self, n).bias.requires_grad for n in self._names))
        if needs_grad:
            if self.train_kernels and self._train_kernels_ok() and not x.requires_grad:
                params = []
                for n in self._names:
                    c = getattr(self, n)
                    params += [c.weight, c.bias]
                L = None if lengths is None else torch.as_tensor(lengths, device=x.device).to(torch.int32)
                return _VGGTrainFn.apply(self, x, L, *params)
            if lengths is not None:
                raise NotImplementedError('length masking under autograd needs train_kernels=True')
            return self._forward_autograd(x)
        return self._forward_kernels(x, lengths)

    def _train_kernels_ok(self):
        return self.resolved_precision() == 'bf16' and all(
            getattr(self, n).in_channels % 64 == 0 and getattr(self, n).out_channels % 64 == 0 for n in self._names[1:])

    def _pack_dgrad(self, name):
        """Packed weights of the input-gradient pass: dx = conv3x3(g, W') with W'[ci][co][ky][kx] = W[co][ci][2-ky][2-kx]."""
        conv = getattr(self, name)
        w = conv.weight
        key = (name, 'dgrad')
        tag = (w.data_ptr(), w._version, str(w.device))
        hit = self._packed.get(key)
        if hit is None or hit[0] != tag:
            with torch.no_grad():
                packed = ops.pack_conv_weight_bf16(w.detach().flip(2, 3).transpose(0, 1).contiguous())
            hit = (tag, packed)
            self._packed[key] = hit
        return hit[1]

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
        prec = self._precision_for(x.size(2))
        L = None if lengths is None else torch.as_tensor(lengths, device=x.device).to(torch.int32)
        c11 = getattr(self, self._names[0])
        nblocks = len(self._names) // 2
        if prec in ('bf16', 'fp16', 'fp32x3'):
            act = torch.float16 if prec == 'fp16' else torch.bfloat16
            x3 = prec == 'fp32x3'
            wk = 'x3' if x3 else ('f16' if prec == 'fp16' else 'bf16')       # both MMA operands must have the same 16-bit format
            # conv11 is bound by its NHWC 16-bit write (2.1 GB per 256 x 4 s batch at the 3.95 TB/s pure-write bandwidth);
            # fuse_first computes it inside conv12's kernel instead (see __init__)
            c12 = getattr(self, self._names[1])
            fuse = self.fuse_first and not x3 and c11.out_channels in (64, 128) and x.size(2) % 2 == 0
            if fuse:
                h = ops.conv12_fused(x, c11.weight, c11.bias, self._pack(self._names[1], wk), c12.bias, c12.out_channels, L,
                                     pool=True, act_dtype=act)
                if L is not None:
                    L = (L + 1) // 2
            else:
                # lazy masking: every consumer below takes the lengths along, so rows beyond the one the next 3x3 kernel reads
                # are neither computed nor zero-filled (3-4 % of a 2-20 s batch was spent writing zeros)
                h = ops.conv11_direct(x, c11.weight, c11.bias, L, out_dtype=act, split=x3, lazy_mask=True)
            for blk in range(1 if fuse else 0, nblocks):
                if blk > 0:
                    c = getattr(self, self._names[2 * blk])
                    h = ops.conv3x3_igemm_bf16(h, self._pack(self._names[2 * blk], wk), c.bias, c.out_channels, L, x3=x3, lazy_mask=True)
                c = getattr(self, self._names[2 * blk + 1])
                last = blk == nblocks - 1
                # CTA pairs (cta_group::2: 256 channels x N pixels per pair, the patch shared) were faster on the pooled layers with
                # >= 256 input channels in round 1; with tap-row reuse, the wave-aware plans and the balanced schedule the
                # single-CTA tiles now win at every batch size (B = 256: conv32 1844 -> 1770 us, conv42 2048 -> 1986 us, step
                # 10.78 -> 10.62 ms; B = 64: conv42 536 -> 476 us; profiles/r2_pair_vs_single.txt).  Results are bit-identical
                # either way; `use_pairs` keeps the pair kernels reachable.
                pair = self.use_pairs and c.in_channels >= 256 and c.out_channels % 256 == 0
                h = ops.conv3x3_igemm_bf16(h, self._pack(self._names[2 * blk + 1], wk), c.bias, c.out_channels, L,
                                           pool=True, ref_layout=last, out_dtype=torch.float32, pair=pair, x3=x3, lazy_mask=True)
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


class _VGGTrainFn(torch.autograd.Function):
    """The whole front-end as one autograd node on this package's kernels (bf16 activations, f32 parameter gradients).
    Forward = CNNs.py:68-91 with every ReLU output kept; backward = the chain rule through it, layer by layer."""

    @staticmethod
    def run_forward(mod, x, lengths):
        """(features, ReLU outputs of every conv in layer order, pooled outputs per block, lengths per block)."""
        names = mod._names
        nblocks = len(names) // 2
        L = lengths
        c11 = getattr(mod, names[0])
        acts = [ops.conv11_direct(x, c11.weight.detach(), c11.bias.detach(), L, out_dtype=torch.bfloat16)]
        pooled, Ls = [], []
        h = acts[0]
        for blk in range(nblocks):
            Ls.append(L)
            if blk > 0:
                c = getattr(mod, names[2 * blk])
                h = ops.conv3x3_igemm_bf16(h, mod._pack(names[2 * blk], 'bf16'), c.bias.detach(), c.out_channels, L)
                acts.append(h)
            c = getattr(mod, names[2 * blk + 1])
            h = ops.conv3x3_igemm_bf16(h, mod._pack(names[2 * blk + 1], 'bf16'), c.bias.detach(), c.out_channels, L)
            acts.append(h)
            last = blk == nblocks - 1
            h = ops.maxpool2x2(h, ref_layout=last, out_dtype=torch.float32 if last else torch.bfloat16)
            pooled.append(h)
            if L is not None:
                L = (L + 1) // 2
        return h, acts, pooled, Ls

    @staticmethod
    def forward(ctx, mod, x, lengths, *params):
        x = x.detach().float().contiguous()
        h, acts, pooled, Ls = _VGGTrainFn.run_forward(mod, x, lengths)
        ctx.mod, ctx.x, ctx.acts, ctx.pooled, ctx.Ls = mod, x, acts, pooled, Ls
        return h

    @staticmethod
    def backward(ctx, dfeat):
        mod, x, acts, pooled, Ls = ctx.mod, ctx.x, ctx.acts, ctx.pooled, ctx.Ls
        names = mod._names
        nblocks = len(names) // 2
        grads = [None] * (2 * len(names))
        gp = dfeat.float().contiguous()                           # gradient at the pooled output of the current block
        for blk in range(nblocks - 1, -1, -1):
            L = Ls[blk]
            y2 = acts[2 * blk + 1]
            g = ops.unpool_relu_bwd(gp, y2)                       # gradient at conv_k2's output (before ReLU)
            x2 = acts[2 * blk]                                    # conv_k2's input = conv_k1's ReLU output
            i2 = 2 * blk + 1
            grads[2 * i2], grads[2 * i2 + 1] = ops.conv3x3_wgrad(x2, g, with_bias=True)
            c2 = getattr(mod, names[i2])
            # gradient at conv_k1's output (before its ReLU): the ReLU backward is fused into the input-gradient store
            g = ops.conv3x3_dgrad(g, mod._pack_dgrad(names[i2]), c2.in_channels, L, relu_mask=x2)
            i1 = 2 * blk
            if blk == 0:
                dw, db = ops.conv11_bwd(x, g, L)
                grads[0], grads[1] = dw, db
            else:
                x1 = pooled[blk - 1]
                grads[2 * i1], grads[2 * i1 + 1] = ops.conv3x3_wgrad(x1, g, with_bias=True)
                c1 = getattr(mod, names[i1])
                gp = ops.conv3x3_dgrad(g, mod._pack_dgrad(names[i1]), c1.in_channels, L)
        ctx.acts = ctx.pooled = None
        return (None, None, None) + tuple(grads)


class VGG3L(_VGG):
    """scripts/CNNs.py:22-52."""
    _divisors = (4, 2, 1)


class VGG4L(_VGG):
    """scripts/CNNs.py:54-91."""
    _divisors = (8, 4, 2, 1)
