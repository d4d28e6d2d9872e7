"""ctypes binding of the C-ABI CUDA library (include/dasv_b200.h).

The product has NO CPU path: ``lib()`` raises if ``libdasv_b200.so`` has not been built, and every
op wrapper in ``ops.py`` raises on non-CUDA tensors.
"""
import ctypes
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DASV_LIB_PATH') or os.path.join(PKG, 'libdasv_b200.so')    # override: A/B runs of two builds

_c = ctypes
_vp, _i, _sz = _c.c_void_p, _c.c_int, _c.c_size_t

# name -> (restype, argtypes); mirrors include/dasv_b200.h one to one
SIGNATURES = {
    'dasv_abi_version': (_i, []),
    'dasv_last_error': (_c.c_char_p, []),
    'dasv_dmha_fwd_workspace_bytes': (_sz, [_i, _i, _i, _i]),
    'dasv_dmha_fwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dasv_dmha_bwd_workspace_bytes': (_sz, [_i, _i, _i, _i]),
    'dasv_dmha_bwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dasv_attention_fwd': (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'dasv_conv11_direct': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'dasv_conv11_direct_lazy': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'dasv_pack_conv_weight_f32': (_i, [_vp, _vp, _i, _i, _vp]),
    'dasv_packed_conv_weight_bf16_elems': (_sz, [_i, _i]),
    'dasv_pack_conv_weight_bf16': (_i, [_vp, _vp, _i, _i, _vp]),
    'dasv_pack_conv_weight_16': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'dasv_pack_conv_weight_x3': (_i, [_vp, _vp, _i, _i, _vp]),
    'dasv_conv3x3_f32': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'dasv_maxpool2x2': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'dasv_conv3x3_igemm_bf16': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'dasv_conv3x3_igemm_workspace_bytes': (_sz, [_i, _i, _i, _i, _i, _i, _i, _i]),
    'dasv_conv12_fused_workspace_bytes': (_sz, [_i]),
    'dasv_conv12_fused_bf16': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'dasv_conv3x3_dgrad_bf16': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'dasv_conv3x3_wgrad_workspace_bytes': (_sz, [_i, _i, _i, _i, _i]),
    'dasv_conv3x3_wgrad_bf16': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'dasv_relu_bwd_bf16': (_i, [_vp, _vp, _sz, _vp]),
    'dasv_unpool_relu_bwd_bf16': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    'dasv_bias_grad_workspace_bytes': (_sz, [_i]),
    'dasv_conv11_bwd_workspace_bytes': (_sz, [_i, _i, _i]),
    'dasv_bias_grad_bf16': (_i, [_vp, _vp, _vp, _i, _sz, _i, _vp]),
    'dasv_conv11_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'dasv_fc_tail_f32': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'dasv_cosine_pairs': (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    'dasv_cosine_matrix_workspace_bytes': (_sz, [_i, _i]),
    'dasv_threshold_counts': (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    'dasv_cosine_matrix': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    'dasv_bn1d_train_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _c.c_float, _c.c_float, _vp]),
    'dasv_bn1d_train_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    'dasv_amsoftmax_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _c.c_float, _c.c_float, _vp]),
    'dasv_amsoftmax_bwd_workspace_bytes': (_sz, [_i, _i]),
    'dasv_amsoftmax_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _c.c_float, _vp]),
    'dasv_h2d_segments': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'dasv_debug_conv_trace': (_i, [_vp]),
    'dasv_logmel_f32': (_i, [_vp, _vp, _i, _c.c_longlong, _vp, _i, _i, _vp, _vp, _i, _c.c_float, _c.c_float, _vp, _i, _vp]),
    'dasv_cmn_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
}

_LIB = None

# kernels launched per successful C call, and the running count bench.py reports as gpu_launches
KERNELS_PER_CALL = {'dasv_conv3x3_igemm_bf16+splitk': 2, 'dasv_amsoftmax_fwd': 3, 'dasv_amsoftmax_bwd': 4, 'dasv_dmha_bwd': 2, 'dasv_conv3x3_wgrad_bf16': 2, 'dasv_bias_grad_bf16': 2, 'dasv_conv11_bwd': 2, 'dasv_attention_fwd': 3, 'dasv_cosine_matrix': 3,
                    'dasv_h2d_segments': 0}     # copies, not kernels
LAUNCHES = {}


class DasvError(RuntimeError):
    pass


def lib():
    """The loaded library.  Fails loudly when it is missing: there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DasvError('%s is missing: build it with `python -m doubleattentionspeakerverification_b200.build` '
                            '(this package has no CPU or PyTorch fallback)' % LIB_PATH)
        try:
            h = ctypes.CDLL(LIB_PATH)
        except OSError:
            import torch  # noqa: F401  (loads libcudart.so.12, which the library links dynamically)
            h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)      # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _LIB = h
    return _LIB


def check(rc, what):
    if rc != 0:
        msg = lib().dasv_last_error()
        raise DasvError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else '?'))
    LAUNCHES[what] = LAUNCHES.get(what, 0) + KERNELS_PER_CALL.get(what, 1)
