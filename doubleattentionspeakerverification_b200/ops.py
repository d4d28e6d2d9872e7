"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw device pointers and the
current CUDA stream, check the status code.  No arithmetic happens here and nothing falls back to
PyTorch or the CPU: a tensor that is not on a CUDA device is an error.
"""
import os

import numpy as np
import torch

from . import _lib

F32, BF16, F16 = 0, 1, 2
CONV_RELU, CONV_POOL, CONV_REF_LAYOUT, CONV_PAIR, CONV_W_F16, CONV_X_F16, CONV_X3 = 1, 2, 4, 8, 16, 32, 64
SPLIT_BF16 = 3
CONV_LAZY_MASK = 256


def _dev(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.DasvError('%s must be a CUDA tensor (this package has no CPU path)' % name)
    return t


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t, name):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise _lib.DasvError('%s: dtype %s not supported (float32, bfloat16 or float16)' % (name, t.dtype))


def _f32(t, name):
    _dev(t, name)
    if t.dtype != torch.float32:
        raise _lib.DasvError('%s must be float32, got %s' % (name, t.dtype))
    return t.contiguous()


def _lengths(lengths, B, device):
    if lengths is None:
        return None
    lengths = torch.as_tensor(lengths, device=device).to(torch.int32).contiguous()
    if lengths.numel() != B:
        raise _lib.DasvError('lengths has %d entries for a batch of %d' % (lengths.numel(), B))
    return lengths


# ------------------------------------------------------------------------------------ pooling
def dmha_fwd(x, query, att=None, lengths=None, keep=None, need_align=True, need_out=True):
    """Fused DoubleMHA forward (att given) or MultiHeadAttention forward (att None).

    Returns dict(out [B,dh]|None, ctx [B,H,dh], lse [B,H], headw [B,H]|None, align [B,T,H]|None)."""
    _dev(x, 'x')
    x = x.contiguous()
    B, T, D = x.shape
    query = _f32(query, 'query')
    dh, H = query.shape
    if dh * H != D:
        raise _lib.DasvError('query [%d,%d] does not match feature size %d' % (dh, H, D))
    with torch.cuda.device(x.device):
        dev = x.device
        att_c = None if att is None else _f32(att, 'att').reshape(-1)
        lengths = _lengths(lengths, B, dev)
        keep_c = None if keep is None else _dev(keep, 'keep').to(torch.uint8).contiguous()
        f = dict(device=dev, dtype=torch.float32)
        out = torch.empty((B, dh), **f) if (att is not None and need_out) else None
        ctx = torch.empty((B, H, dh), **f)
        lse = torch.empty((B, H), **f)
        headw = torch.empty((B, H), **f) if att is not None else None
        align = torch.empty((B, T, H), **f) if need_align else None
        L = _lib.lib()
        ws = _zeroed_workspace(dev, max(int(L.dasv_dmha_fwd_workspace_bytes(B, T, D, H)), 4))
        rc = L.dasv_dmha_fwd(_p(x), _dtype_code(x, 'x'), _p(lengths), _p(query), _p(att_c), _p(keep_c),
                             _p(out), _p(ctx), _p(lse), _p(headw), _p(align), _p(ws), B, T, D, H, _stream())
        if rc != 0:
            _ws_cache.clear()                          # a failed launch may leave the counter dirty
        _lib.check(rc, 'dasv_dmha_fwd')
    return dict(out=out, ctx=ctx, lse=lse, headw=headw, align=align)


_ws_cache = {}


def _zeroed_workspace(dev, nbytes):
    """Per (device, stream) workspace for dasv_dmha_fwd: zeroed once, the kernel hands it back zeroed
    (include/dasv_b200.h), so calls on the same stream can share it without a memset per call."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros((max(nbytes, 256),), device=dev, dtype=torch.uint8)
        _ws_cache[key] = ws
    return ws


def dmha_bwd(x, query, att, g_out, g_ctx, ctx, lse, headw, lengths=None):
    """Closed-form backward of dmha_fwd.  Returns (dx, dquery [dh,H], datt [dh]|None)."""
    _dev(x, 'x')
    x = x.contiguous()
    B, T, D = x.shape
    query = _f32(query, 'query')
    dh, H = query.shape
    with torch.cuda.device(x.device):
        dev = x.device
        att_c = None if att is None else _f32(att, 'att').reshape(-1)
        g_out = None if g_out is None else _f32(g_out, 'g_out')
        g_ctx = None if g_ctx is None else _f32(g_ctx, 'g_ctx')
        lengths = _lengths(lengths, B, dev)
        dx = torch.empty_like(x)
        dquery = torch.empty((dh, H), device=dev, dtype=torch.float32)
        datt = torch.empty((dh,), device=dev, dtype=torch.float32) if att is not None else None
        L = _lib.lib()
        ws = torch.empty((max(int(L.dasv_dmha_bwd_workspace_bytes(B, T, D, H)), 4),), device=dev, dtype=torch.uint8)
        rc = L.dasv_dmha_bwd(_p(x), _dtype_code(x, 'x'), _p(lengths), _p(query), _p(att_c), _p(g_out), _p(g_ctx),
                             _p(ctx), _p(lse), _p(headw), _p(dx), _p(dquery), _p(datt), _p(ws), B, T, D, H, _stream())
        _lib.check(rc, 'dasv_dmha_bwd')
    return dx, dquery, datt


def attention_fwd(x, att, lengths=None, keep=None):
    """Single-query attention over dim 1 (Attention / stand-alone HeadAttention).  Returns (out [B,D], align [B,T])."""
    _dev(x, 'x')
    x = x.contiguous()
    B, T, D = x.shape
    with torch.cuda.device(x.device):
        att_c = _f32(att, 'att').reshape(-1)
        lengths = _lengths(lengths, B, x.device)
        keep_c = None if keep is None else _dev(keep, 'keep').to(torch.uint8).contiguous()
        out = torch.empty((B, D), device=x.device, dtype=torch.float32)
        align = torch.empty((B, T), device=x.device, dtype=torch.float32)
        rc = _lib.lib().dasv_attention_fwd(_p(x), _dtype_code(x, 'x'), _p(lengths), _p(keep_c), _p(att_c), _p(out), _p(align),
                                           B, T, D, _stream())
        _lib.check(rc, 'dasv_attention_fwd')
    return out, align


# ------------------------------------------------------------------------------------ front-end
def pack_conv_weight_f32(w):
    w = _f32(w, 'w')
    Cout, Cin = w.shape[0], w.shape[1]
    with torch.cuda.device(w.device):
        p = torch.empty((9, Cin, Cout), device=w.device, dtype=torch.float32)
        _lib.check(_lib.lib().dasv_pack_conv_weight_f32(_p(w), _p(p), Cout, Cin, _stream()), 'dasv_pack_conv_weight_f32')
    return p


def pack_conv_weight_bf16(w, dtype=torch.bfloat16):
    """w [Cout,Cin,3,3] f32 -> the tensor-core A operand [Cout_pad][9][Cin] in ``dtype`` (bfloat16, or float16 for three
    more mantissa bits: weights are far inside fp16's range)."""
    w = _f32(w, 'w')
    Cout, Cin = w.shape[0], w.shape[1]
    with torch.cuda.device(w.device):
        n = int(_lib.lib().dasv_packed_conv_weight_bf16_elems(Cout, Cin))
        p = torch.empty((n,), device=w.device, dtype=dtype)
        _lib.check(_lib.lib().dasv_pack_conv_weight_16(_p(w), _p(p), Cout, Cin, _dtype_code(p, 'packed'), _stream()), 'dasv_pack_conv_weight_16')
    return p


def conv11_direct(x, w, bias, lengths=None, out_dtype=torch.float32, split=False, lazy_mask=False):
    """x [B,T,F] f32 -> relu(conv3x3(x) + bias) as NHWC [B,T,F,Cout]; ``split=True``: [B,T,F,2*Cout] bf16 holding every fp32
    value as hi = bf16(v) (channel c) and lo = bf16(v - hi) (channel Cout + c), the activation format of the fp32x3 mode.
    ``lazy_mask`` (with lengths): rows beyond row ``lengths[b]`` may stay unwritten (see ``conv3x3_igemm_bf16``)."""
    x = _f32(x, 'x')
    B, T, Fq = x.shape
    w, bias = _f32(w, 'w'), _f32(bias, 'bias')
    Cout = w.shape[0]
    with torch.cuda.device(x.device):
        lengths = _lengths(lengths, B, x.device)
        y = torch.empty((B, T, Fq, 2 * Cout if split else Cout), device=x.device, dtype=torch.bfloat16 if split else out_dtype)
        fn = _lib.lib().dasv_conv11_direct_lazy if (lazy_mask and lengths is not None) else _lib.lib().dasv_conv11_direct
        rc = fn(_p(x), _p(w), _p(bias), _p(lengths), _p(y), SPLIT_BF16 if split else _dtype_code(y, 'y'), B, T, Fq, Cout, _stream())
        _lib.check(rc, 'dasv_conv11_direct')
    return y


def pack_conv_weight_x3(w):
    """w [Cout,Cin,3,3] f32 -> [Cout_pad][9][3*Cin] bf16 = per tap [hi | hi | lo]: the A operand of the fp32x3 mode."""
    w = _f32(w, 'w')
    Cout, Cin = w.shape[0], w.shape[1]
    with torch.cuda.device(w.device):
        n = 3 * int(_lib.lib().dasv_packed_conv_weight_bf16_elems(Cout, Cin))
        p = torch.empty((n,), device=w.device, dtype=torch.bfloat16)
        _lib.check(_lib.lib().dasv_pack_conv_weight_x3(_p(w), _p(p), Cout, Cin, _stream()), 'dasv_pack_conv_weight_x3')
    return p


def conv3x3_f32(x, wp, bias, lengths=None):
    """fp32 CUDA-core conv3x3 + bias + ReLU on NHWC; wp from pack_conv_weight_f32."""
    x = _f32(x, 'x')
    B, T, Fq, Cin = x.shape
    Cout = wp.shape[2]
    with torch.cuda.device(x.device):
        lengths = _lengths(lengths, B, x.device)
        y = torch.empty((B, T, Fq, Cout), device=x.device, dtype=torch.float32)
        rc = _lib.lib().dasv_conv3x3_f32(_p(x), _p(wp), _p(_f32(bias, 'bias')), _p(lengths), _p(y), B, T, Fq, Cin, Cout, _stream())
        _lib.check(rc, 'dasv_conv3x3_f32')
    return y


def maxpool2x2(x, ref_layout=False, out_dtype=None):
    """2x2 stride-2 ceil-mode max-pool on NHWC; ref_layout=True writes [B,T2,C*F2] (feature = c*F2+f)."""
    _dev(x, 'x')
    x = x.contiguous()
    B, T, Fq, C = x.shape
    out_dtype = out_dtype or x.dtype
    T2, F2 = (T + 1) // 2, (Fq + 1) // 2
    with torch.cuda.device(x.device):
        shape = (B, T2, C * F2) if ref_layout else (B, T2, F2, C)
        y = torch.empty(shape, device=x.device, dtype=out_dtype)
        rc = _lib.lib().dasv_maxpool2x2(_p(x), _dtype_code(x, 'x'), _p(y), _dtype_code(y, 'y'), int(ref_layout), B, T, Fq, C, _stream())
        _lib.check(rc, 'dasv_maxpool2x2')
    return y


def conv3x3_igemm_bf16(x, wp, bias, Cout, lengths=None, pool=False, ref_layout=False, out_dtype=torch.bfloat16, pair=False, relu=True, x3=False,
                       lazy_mask=False):
    """tcgen05 implicit-GEMM conv3x3 + bias + ReLU (+ fused 2x2 ceil max-pool) on NHWC 16-bit activations (bf16 or fp16; the
    output has the input's format) with bf16 or fp16 packed weights, fp32 accumulation.

    ``lazy_mask`` (with lengths, for pipelines that carry the lengths through every layer): of the output rows at or beyond
    an utterance's length only the one the next 3x3 layer reads is guaranteed zero; tiles that lie wholly further down are
    skipped and their memory stays unwritten (valid rows are unaffected: a 3x3 kernel never looks further than one row)."""
    _dev(x, 'x')
    if x.dtype not in (torch.bfloat16, torch.float16) or wp.dtype not in (torch.bfloat16, torch.float16):
        raise _lib.DasvError('conv3x3_igemm_bf16: x and wp must be bfloat16 or float16')
    x = x.contiguous()
    B, T, Fq, Cin = x.shape
    flags = (CONV_RELU if relu else 0) | (CONV_POOL if pool else 0) | (CONV_REF_LAYOUT if ref_layout else 0) | (CONV_PAIR if pair else 0)
    flags |= (CONV_W_F16 if wp.dtype == torch.float16 else 0) | (CONV_X_F16 if x.dtype == torch.float16 else 0)
    flags |= CONV_LAZY_MASK if (lazy_mask and lengths is not None) else 0
    if x3:                                                   # fp32x3 mode: x is [hi | lo] split bf16 (2 * Cin channels), so is an NHWC y
        if x.dtype != torch.bfloat16 or Cin % 2:
            raise _lib.DasvError('conv3x3_igemm_bf16: the fp32x3 mode takes split bf16 activations')
        flags |= CONV_X3
        Cin //= 2
    with torch.cuda.device(x.device):
        lengths = _lengths(lengths, B, x.device)
        if pool:
            if Fq % 2:
                raise _lib.DasvError('conv3x3_igemm_bf16: the pooled epilogue needs an even number of bins (F=%d)' % Fq)
            T2, F2 = (T + 1) // 2, Fq // 2
            shape = (B, T2, Cout * F2) if ref_layout else (B, T2, F2, 2 * Cout if x3 else Cout)
        else:
            shape = (B, T, Fq, 2 * Cout if x3 else Cout)
        if ref_layout and out_dtype != torch.float32:
            out_dtype = x.dtype
        y = torch.empty(shape, device=x.device, dtype=out_dtype if ref_layout else x.dtype)
        L = _lib.lib()
        key = (_dtype_code(y, 'y'), flags, lengths is not None, B, T, Fq, Cin, Cout, x.device.index, os.environ.get('DASV_CONV_NOSPLITK'), os.environ.get('DASV_CONV_SPLITK'))
        nws = _conv_ws_bytes.get(key)
        if nws is None:                                      # > 0 only for small batches, which run split along K
            nws = _conv_ws_bytes[key] = int(L.dasv_conv3x3_igemm_workspace_bytes(key[0], flags, int(key[2]), B, T, Fq, Cin, Cout))
        ws = torch.empty((nws,), device=x.device, dtype=torch.uint8) if nws else None
        rc = L.dasv_conv3x3_igemm_bf16(_p(x), _p(wp), _p(_f32(bias, 'bias')), _p(lengths), _p(y), key[0],
                                       flags, B, T, Fq, Cin, Cout, _p(ws), _stream())
        _lib.check(rc, 'dasv_conv3x3_igemm_bf16+splitk' if nws else 'dasv_conv3x3_igemm_bf16')
    return y


_scratch_cache = {}
_conv_ws_bytes = {}


def conv12_fused(x, w11, b11, wp, bias, Cout, lengths=None, pool=True, act_dtype=torch.bfloat16):
    """conv11 + conv12 in one kernel: x [B,T,F] f32 -> relu(conv12(relu(conv11(x)))) (+ pool) as NHWC ``act_dtype``; the
    C1-channel tensor in between never goes to HBM.  Bit-identical to conv11_direct + conv3x3_igemm_bf16."""
    x = _f32(x, 'x')
    B, T, Fq = x.shape
    w11, b11 = _f32(w11, 'w11'), _f32(b11, 'b11')
    C1 = w11.shape[0]
    if wp.dtype != act_dtype:
        raise _lib.DasvError('conv12_fused: packed weights and activations must share their 16-bit format')
    flags = CONV_RELU | (CONV_POOL if pool else 0) | ((CONV_W_F16 | CONV_X_F16) if act_dtype == torch.float16 else 0)
    with torch.cuda.device(x.device):
        lengths = _lengths(lengths, B, x.device)
        L = _lib.lib()
        key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream, C1)
        ws = _scratch_cache.get(key)                         # per (device, stream): kernels of one stream run one after another
        if ws is None:
            ws = torch.empty((int(L.dasv_conv12_fused_workspace_bytes(C1)),), device=x.device, dtype=torch.uint8)
            _scratch_cache[key] = ws
        shape = (B, (T + 1) // 2, Fq // 2, Cout) if pool else (B, T, Fq, Cout)
        y = torch.empty(shape, device=x.device, dtype=act_dtype)
        rc = L.dasv_conv12_fused_bf16(_p(x), _p(w11), _p(b11), _p(wp), _p(_f32(bias, 'bias')), _p(lengths), _p(y), _p(ws),
                                      _dtype_code(y, 'y'), flags, B, T, Fq, C1, Cout, _stream())
        _lib.check(rc, 'dasv_conv12_fused_bf16')
    return y


# ------------------------------------------------------------------------------------ tail / scoring
def fc_tail(pooled, w1t, b1, w2t, b2, bn_scale, bn_shift):
    pooled = _f32(pooled, 'pooled')
    B, Din = pooled.shape
    E = w1t.shape[1]
    with torch.cuda.device(pooled.device):
        emb = torch.empty((B, E), device=pooled.device, dtype=torch.float32)
        rc = _lib.lib().dasv_fc_tail_f32(_p(pooled), _p(w1t), _p(b1), _p(w2t), _p(b2), _p(bn_scale), _p(bn_shift), _p(emb),
                                         B, Din, E, _stream())
        _lib.check(rc, 'dasv_fc_tail_f32')
    return emb


def cosine_pairs(emb, ia, ib):
    emb = _f32(emb, 'emb')
    _dev(ia, 'ia'); _dev(ib, 'ib')
    n = ia.numel()
    if ib.numel() != n:
        raise _lib.DasvError('cosine_pairs: %d and %d indices' % (n, ib.numel()))
    if n:
        # the kernel reads emb[ia], emb[ib] unchecked: validate on the host side (one small reduction)
        both = torch.stack([ia.reshape(-1), ib.reshape(-1)])
        lo, hi = int(both.min()), int(both.max())
        if lo < 0 or hi >= emb.shape[0]:
            raise _lib.DasvError('cosine_pairs: trial index out of range [0, %d): min %d, max %d' % (emb.shape[0], lo, hi))
    ia = ia.to(torch.int32).contiguous()
    ib = ib.to(torch.int32).contiguous()
    with torch.cuda.device(emb.device):
        scores = torch.empty((n,), device=emb.device, dtype=torch.float32)
        _lib.check(_lib.lib().dasv_cosine_pairs(_p(emb), _p(ia), _p(ib), _p(scores), n, emb.shape[1], _stream()), 'dasv_cosine_pairs')
    return scores


def cosine_matrix(enrol, test):
    enrol, test = _f32(enrol, 'enrol'), _f32(test, 'test')
    Ne, E = enrol.shape
    Nt = test.shape[0]
    with torch.cuda.device(enrol.device):
        scores = torch.empty((Ne, Nt), device=enrol.device, dtype=torch.float32)
        ws = torch.empty((Ne + Nt,), device=enrol.device, dtype=torch.float32)
        _lib.check(_lib.lib().dasv_cosine_matrix(_p(enrol), _p(test), _p(scores), _p(ws), Ne, Nt, E, _stream()), 'dasv_cosine_matrix')
    return scores


def h2d_segments(dst, src_host, src_off, dst_off, nbytes):
    """``len(nbytes)`` host->device copies on the current stream (one cudaMemcpyAsync each, issued from C): segment i =
    ``nbytes[i]`` bytes from byte ``src_off[i]`` of the (pinned) host tensor ``src_host`` to byte ``dst_off[i]`` of ``dst``."""
    _dev(dst, 'dst')
    if src_host.is_cuda or not src_host.is_contiguous() or not dst.is_contiguous():
        raise _lib.DasvError('h2d_segments: src_host must be a contiguous host tensor, dst a contiguous device tensor')
    so, do, nb = (np.ascontiguousarray(a, dtype=np.int64) for a in (src_off, dst_off, nbytes))
    if not (so.shape == do.shape == nb.shape) or so.ndim != 1:
        raise _lib.DasvError('h2d_segments: the offset / size arrays must be 1-D and equally long')
    if len(nb) and (int((so + nb).max()) > src_host.numel() * src_host.element_size() or int((do + nb).max()) > dst.numel() * dst.element_size()
                    or int(min(so.min(), do.min(), nb.min())) < 0):
        raise _lib.DasvError('h2d_segments: a segment lies outside its buffer')
    with torch.cuda.device(dst.device):
        rc = _lib.lib().dasv_h2d_segments(_p(dst), src_host.data_ptr(), so.ctypes.data, do.ctypes.data, nb.ctypes.data, len(nb), _stream())
        _lib.check(rc, 'dasv_h2d_segments')


def threshold_counts(scores, thresholds):
    """counts[k] = #{scores >= thresholds[k]} (double comparison).  scores f32 [n], thresholds f64 [n_th] -> int64 [n_th]."""
    scores = _f32(scores.reshape(-1), 'scores')
    with torch.cuda.device(scores.device):
        th = torch.as_tensor(thresholds, device=scores.device, dtype=torch.float64).contiguous()
        out = torch.empty((th.numel(),), device=scores.device, dtype=torch.int64)
        rc = _lib.lib().dasv_threshold_counts(_p(scores) if scores.numel() else None, scores.numel(), _p(th), th.numel(), _p(out), _stream())
        _lib.check(rc, 'dasv_threshold_counts')
    return out


# ------------------------------------------------------------------------------------ features
def logmel(wave, n_samples, frames, Tmax, window, hop, melw, mel_range, preem, scale, cmn=True):
    # cmn: False / None = raw log-mel, True or 'cmn' = mean normalisation, 'cmvn' = mean and variance (data.py:21-30)
    """Log mel-filterbank features (+ CMN) of a padded waveform batch (csrc/features.cu).  All tensors on the device:
    wave [B,N] f32, n_samples/frames [B] i32, window [win_length] f32, melw [n_mels,257] f32, mel_range [n_mels,2] i32.
    Returns [B,Tmax,n_mels] f32; rows t >= frames[b] are zero."""
    _dev(wave, 'wave')
    wave = wave.contiguous()
    B, N = wave.shape
    n_mels = melw.shape[0]
    with torch.cuda.device(wave.device):
        out = torch.zeros((B, Tmax, n_mels), device=wave.device, dtype=torch.float32)
        if B == 0 or Tmax == 0:
            return out
        L = _lib.lib()
        rc = L.dasv_logmel_f32(_p(wave), _p(n_samples), B, N, _p(window), window.numel(), hop, _p(melw), _p(mel_range), n_mels,
                               preem, scale, _p(out), Tmax, _stream())
        _lib.check(rc, 'dasv_logmel_f32')
        if cmn:
            rc = L.dasv_cmn_f32(_p(out), _p(frames), B, Tmax, n_mels, 1 if cmn == 'cmvn' else 0, _stream())
            _lib.check(rc, 'dasv_cmn_f32')
    return out


# ------------------------------------------------------------------------------------ training side of the front-end
def conv3x3_dgrad(g, wp_rot, Cx, lengths=None, relu_mask=None):
    """Input gradient of a 3x3 pad-1 convolution: g [B,T,F,Cg] bf16, wp_rot = pack_conv_weight_bf16 of the rotated,
    transposed weights.  relu_mask [B,T,F,Cx] bf16 (optional): zero the result where it is <= 0.  Returns dx bf16."""
    _dev(g, 'g')
    if g.dtype != torch.bfloat16:
        raise _lib.DasvError('conv3x3_dgrad: g must be bfloat16')
    g = g.contiguous()
    B, T, Fq, Cg = g.shape
    if relu_mask is not None and (relu_mask.dtype != torch.bfloat16 or tuple(relu_mask.shape) != (B, T, Fq, Cx) or not relu_mask.is_contiguous()):
        raise _lib.DasvError('conv3x3_dgrad: relu_mask must be a contiguous bf16 [B,T,F,Cx] tensor')
    with torch.cuda.device(g.device):
        lengths = _lengths(lengths, B, g.device)
        dx = torch.empty((B, T, Fq, Cx), device=g.device, dtype=torch.bfloat16)
        rc = _lib.lib().dasv_conv3x3_dgrad_bf16(_p(g), _p(wp_rot), _p(relu_mask), _p(lengths), _p(dx), B, T, Fq, Cg, Cx, _stream())
        _lib.check(rc, 'dasv_conv3x3_dgrad_bf16')
    return dx


def conv3x3_wgrad(x, g, dw=None, with_bias=False):
    """Weight gradient of a 3x3 pad-1 convolution (csrc/conv_wgrad.cu).  x [B,T,F,Cin] bf16 NHWC (layer input),
    g [B,T,F,Cout] bf16 NHWC (gradient at the conv output).  Returns dw [Cout,Cin,3,3] f32 (added to ``dw`` if given), or
    (dw, db [Cout]) with ``with_bias`` (the bias gradient falls out of the same kernel)."""
    _dev(x, 'x'); _dev(g, 'g')
    if x.dtype != torch.bfloat16 or g.dtype != torch.bfloat16:
        raise _lib.DasvError('conv3x3_wgrad needs bf16 activations')
    x, g = x.contiguous(), g.contiguous()
    B, T, F, Cin = x.shape
    Cout = g.shape[3]
    if g.shape[:3] != x.shape[:3]:
        raise _lib.DasvError('x and g disagree on [B,T,F]')
    with torch.cuda.device(x.device):
        L = _lib.lib()
        acc = dw is not None
        if acc and with_bias:
            raise _lib.DasvError('conv3x3_wgrad: with_bias cannot be combined with accumulation into dw')
        if dw is None:
            dw = torch.empty((Cout, Cin, 3, 3), device=x.device, dtype=torch.float32)
        elif dw.shape != (Cout, Cin, 3, 3) or dw.dtype != torch.float32 or not dw.is_contiguous():
            raise _lib.DasvError('dw must be a contiguous f32 [Cout,Cin,3,3] tensor')
        db = torch.empty((Cout,), device=x.device, dtype=torch.float32) if with_bias else None
        ws = torch.empty((max(int(L.dasv_conv3x3_wgrad_workspace_bytes(B, T, F, Cin, Cout)), 16),), device=x.device, dtype=torch.uint8)
        rc = L.dasv_conv3x3_wgrad_bf16(_p(x), _p(g), _p(dw), _p(db), _p(ws), 1 if acc else 0, B, T, F, Cin, Cout, _stream())
        _lib.check(rc, 'dasv_conv3x3_wgrad_bf16')
    return (dw, db) if with_bias else dw


def relu_bwd_(g, y):
    """In place g = y > 0 ? g : 0 (bf16, same shape)."""
    _dev(g, 'g'); _dev(y, 'y')
    if g.dtype != torch.bfloat16 or y.dtype != torch.bfloat16 or g.shape != y.shape or not g.is_contiguous() or not y.is_contiguous():
        raise _lib.DasvError('relu_bwd_ needs contiguous bf16 tensors of one shape')
    with torch.cuda.device(g.device):
        _lib.check(_lib.lib().dasv_relu_bwd_bf16(_p(g), _p(y), g.numel(), _stream()), 'dasv_relu_bwd_bf16')
    return g


def unpool_relu_bwd(gp, y):
    """Backward of relu + 2x2 ceil-mode max-pool.  y [B,T,F,C] bf16 (pre-pool ReLU output); gp bf16 [B,T2,F2,C] or, for the
    front-end's output, f32 [B,T2,C*F2].  Returns g [B,T,F,C] bf16."""
    _dev(gp, 'gp'); _dev(y, 'y')
    B, T, Fq, C = y.shape
    ref = gp.dim() == 3
    if y.dtype != torch.bfloat16 or gp.dtype != (torch.float32 if ref else torch.bfloat16):
        raise _lib.DasvError('unpool_relu_bwd: y must be bf16 and gp bf16 NHWC or f32 [B,T2,C*F2]')
    gp, y = gp.contiguous(), y.contiguous()
    with torch.cuda.device(y.device):
        g = torch.empty_like(y)
        _lib.check(_lib.lib().dasv_unpool_relu_bwd_bf16(_p(gp), int(ref), _p(y), _p(g), B, T, Fq, C, _stream()), 'dasv_unpool_relu_bwd_bf16')
    return g


def bias_grad(g):
    """Column sums of g [..., C] bf16 -> f32 [C]."""
    _dev(g, 'g')
    g = g.contiguous()
    C = g.shape[-1]
    with torch.cuda.device(g.device):
        db = torch.empty((C,), device=g.device, dtype=torch.float32)
        ws = torch.empty((int(_lib.lib().dasv_bias_grad_workspace_bytes(C)),), device=g.device, dtype=torch.uint8)
        _lib.check(_lib.lib().dasv_bias_grad_bf16(_p(g), _p(db), _p(ws), 0, g.numel() // C, C, _stream()), 'dasv_bias_grad_bf16')
    return db


def conv11_bwd(x, g, lengths=None):
    """conv11 (Cin = 1) parameter gradients: x [B,T,F] f32, g [B,T,F,C] bf16 -> (dw [C,1,3,3], db [C]) f32."""
    x = _f32(x, 'x')
    _dev(g, 'g')
    g = g.contiguous()
    B, T, Fq, C = g.shape
    with torch.cuda.device(g.device):
        lengths = _lengths(lengths, B, g.device)
        dw = torch.empty((C, 1, 3, 3), device=g.device, dtype=torch.float32)
        db = torch.empty((C,), device=g.device, dtype=torch.float32)
        ws = torch.empty((max(int(_lib.lib().dasv_conv11_bwd_workspace_bytes(B, T, C)), 16),), device=g.device, dtype=torch.uint8)
        _lib.check(_lib.lib().dasv_conv11_bwd(_p(x), _p(g), _p(lengths), _p(dw), _p(db), _p(ws), 0, B, T, Fq, C, _stream()), 'dasv_conv11_bwd')
    return dw, db


# ------------------------------------------------------------------------------------ training-mode tail
def bn1d_train_fwd(x, gamma, beta, running_mean, running_var, eps, momentum):
    """BatchNorm1d with batch statistics on x [B,E] f32; running_mean / running_var (or None) are updated in place.
    Returns (y, save_mean, save_invstd)."""
    x = _f32(x, 'x')
    B, E = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty_like(x)
        sm = torch.empty((E,), device=x.device, dtype=torch.float32)
        si = torch.empty((E,), device=x.device, dtype=torch.float32)
        rc = _lib.lib().dasv_bn1d_train_fwd(_p(x), _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(y), _p(sm), _p(si),
                                            B, E, float(eps), float(momentum), _stream())
        _lib.check(rc, 'dasv_bn1d_train_fwd')
    return y, sm, si


def bn1d_train_bwd(dy, x, gamma, save_mean, save_invstd):
    dy, x = _f32(dy, 'dy'), _f32(x, 'x')
    B, E = x.shape
    with torch.cuda.device(x.device):
        dx = torch.empty_like(x)
        dg = torch.empty((E,), device=x.device, dtype=torch.float32)
        db = torch.empty((E,), device=x.device, dtype=torch.float32)
        rc = _lib.lib().dasv_bn1d_train_bwd(_p(dy), _p(x), _p(gamma), _p(save_mean), _p(save_invstd), _p(dx), _p(dg), _p(db), B, E, _stream())
        _lib.check(rc, 'dasv_bn1d_train_bwd')
    return dx, dg, db


def amsoftmax_fwd(x, W, label, s, margin_scaled):
    """x [B,E], W [E,S] f32, label [B] int64 (device) -> (costh, logits, inv_x, inv_w)."""
    x, W = _f32(x, 'x'), _f32(W, 'W')
    B, E = x.shape
    S = W.shape[1]
    label = _dev(label, 'label').to(torch.int64).contiguous()
    with torch.cuda.device(x.device):
        f = dict(device=x.device, dtype=torch.float32)
        costh, logits = torch.empty((B, S), **f), torch.empty((B, S), **f)
        ix, iw = torch.empty((B,), **f), torch.empty((S,), **f)
        rc = _lib.lib().dasv_amsoftmax_fwd(_p(x), _p(W), _p(label), _p(costh), _p(logits), _p(ix), _p(iw), B, E, S,
                                           float(s), float(margin_scaled), _stream())
        _lib.check(rc, 'dasv_amsoftmax_fwd')
    return costh, logits, ix, iw


def amsoftmax_bwd(dcosth, dlogits, x, W, costh, inv_x, inv_w, s):
    x, W = _f32(x, 'x'), _f32(W, 'W')
    B, E = x.shape
    S = W.shape[1]
    with torch.cuda.device(x.device):
        dcosth = None if dcosth is None else _f32(dcosth, 'dcosth')
        dlogits = None if dlogits is None else _f32(dlogits, 'dlogits')
        dx, dW = torch.empty_like(x), torch.empty_like(W)
        L = _lib.lib()
        ws = torch.empty((max(int(L.dasv_amsoftmax_bwd_workspace_bytes(B, S)), 16),), device=x.device, dtype=torch.uint8)
        rc = L.dasv_amsoftmax_bwd(_p(dcosth), _p(dlogits), _p(x), _p(W), _p(costh), _p(inv_x), _p(inv_w), _p(dx), _p(dW), _p(ws),
                                  B, E, S, float(s), _stream())
        _lib.check(rc, 'dasv_amsoftmax_bwd')
    return dx, dW
