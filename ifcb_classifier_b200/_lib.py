"""ctypes binding of the C ABI in include/ifcb_b200.h (libifcb_b200.so).

There is NO CPU fallback: if the shared library is missing or a call fails this
module raises -- the product path never routes through ``oracle/`` or PyTorch ops.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libifcb_b200.so')

IFCB_OUT_F32_NCHW, IFCB_OUT_BF16_NCHW, IFCB_OUT_U8_GRAY = 0, 1, 2
IFCB_PASS_PILLOW12, IFCB_PASS_HV = 0, 1
IFCB_PRE_BAD_TABLE, IFCB_PRE_TOO_LARGE = 1, 2
IFCB_STEM_IN_U8_GRAY, IFCB_STEM_IN_F32_NCHW = 0, 1
IFCB_POOL_MAX, IFCB_POOL_AVG_AFFINE = 0, 1
IFCB_MAX_SEGMENTS = 4
IFCB_ACT_BF16, IFCB_ACT_FP16 = 0, 1
IFCB_CONV_AUTO, IFCB_CONV_IM2COL, IFCB_CONV_WINDOW, IFCB_CONV_IM2COL_PAIR = 0, 1, 2, 3


class ConvSegment(C.Structure):
    _fields_ = [('n_begin', C.c_int32), ('n_end', C.c_int32), ('d_out', C.c_void_p),
                ('ld', C.c_int32), ('relu', C.c_int32), ('pad_h', C.c_int32), ('pad_w', C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [('d_in', C.c_void_p), ('in_ld', C.c_int32), ('Cin', C.c_int32),
                ('batch_cap', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('in_pad_h', C.c_int32), ('in_pad_w', C.c_int32),
                ('kh', C.c_int32), ('kw', C.c_int32), ('stride_h', C.c_int32), ('stride_w', C.c_int32),
                ('pad_h', C.c_int32), ('pad_w', C.c_int32), ('Cout', C.c_int32),
                ('d_weight', C.c_void_p), ('d_scale', C.c_void_p), ('d_shift', C.c_void_p),
                ('n_seg', C.c_int32), ('seg', ConvSegment * IFCB_MAX_SEGMENTS),
                ('d_residual', C.c_void_p), ('res_ld', C.c_int32), ('res_pad_h', C.c_int32), ('res_pad_w', C.c_int32),
                ('tile_n', C.c_int32), ('algo', C.c_int32), ('dtype', C.c_int32), ('d_stats', C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [('d_in', C.c_void_p), ('in_ld', C.c_int32), ('Cin', C.c_int32),
                ('batch', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('kh', C.c_int32), ('kw', C.c_int32), ('stride_h', C.c_int32), ('stride_w', C.c_int32),
                ('pad_h', C.c_int32), ('pad_w', C.c_int32),
                ('d_dout', C.c_void_p), ('dout_ld', C.c_int32), ('Cout', C.c_int32),
                ('d_dweight', C.c_void_p), ('dtype', C.c_int32), ('in_pad_h', C.c_int32), ('in_pad_w', C.c_int32),
                ('dout_pad_h', C.c_int32), ('dout_pad_w', C.c_int32)]


class ViewDesc(C.Structure):
    _fields_ = [('d', C.c_void_p), ('ld', C.c_int32), ('C', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('pad_h', C.c_int32), ('pad_w', C.c_int32)]


class RepackItem(C.Structure):
    _fields_ = [('d_master', C.c_void_p), ('d_wfwd', C.c_void_p), ('d_wdgrad', C.c_void_p), ('Cout', C.c_int32), ('taps', C.c_int32),
                ('Cin', C.c_int32), ('Cin_pad', C.c_int32), ('Cout_padk', C.c_int32), ('reserved', C.c_int32)]


class StemDesc(C.Structure):
    _fields_ = [('d_in', C.c_void_p), ('in_kind', C.c_int32),
                ('batch_cap', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('kh', C.c_int32), ('kw', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32),
                ('Cout', C.c_int32), ('d_weight', C.c_void_p), ('d_scale', C.c_void_p),
                ('d_shift', C.c_void_p), ('d_wgray', C.c_void_p), ('d_wconst', C.c_void_p),
                ('in_scale', C.c_float * 3), ('in_shift', C.c_float * 3),
                ('d_out', C.c_void_p), ('out_ld', C.c_int32), ('relu', C.c_int32), ('dtype', C.c_int32),
                ('out_pad_h', C.c_int32), ('out_pad_w', C.c_int32)]


class PoolDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('d_in', C.c_void_p), ('in_ld', C.c_int32), ('C', C.c_int32),
                ('batch_cap', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('k', C.c_int32), ('stride', C.c_int32), ('pad', C.c_int32),
                ('d_out', C.c_void_p), ('out_ld', C.c_int32),
                ('d_scale', C.c_void_p), ('d_shift', C.c_void_p), ('relu', C.c_int32), ('dtype', C.c_int32),
                ('in_pad_h', C.c_int32), ('in_pad_w', C.c_int32), ('out_pad_h', C.c_int32), ('out_pad_w', C.c_int32),
                ('ceil_mode', C.c_int32)]


class HeadDesc(C.Structure):
    _fields_ = [('d_in', C.c_void_p), ('in_ld', C.c_int32), ('C', C.c_int32), ('HW', C.c_int32),
                ('batch_cap', C.c_int32), ('n_classes', C.c_int32),
                ('d_weight', C.c_void_p), ('d_bias', C.c_void_p), ('d_scores', C.c_void_p),
                ('d_logits', C.c_void_p), ('d_top1', C.c_void_p), ('d_top1_score', C.c_void_p),
                ('dtype', C.c_int32)]


# name -> (restype, argtypes); mirrors include/ifcb_b200.h one to one
_V = C.POINTER(ViewDesc)
_SIGNATURES = {
    'ifcb_abi_version': (C.c_int, []),
    'ifcb_last_error': (C.c_char_p, []),
    'ifcb_sm_count': (C.c_int, []),
    'ifcb_parse_adc': (C.c_int64, [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    'ifcb_format_scores_json': (C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]),
    'ifcb_preprocess': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_float),
                                  C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'ifcb_plan_create': (C.c_int, [C.POINTER(C.c_void_p)]),
    'ifcb_plan_destroy': (C.c_int, [C.c_void_p]),
    'ifcb_plan_run': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    'ifcb_plan_run_at': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_plan_run_range': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_plan_refresh': (C.c_int, [C.c_void_p]),
    'ifcb_plan_num_layers': (C.c_int, [C.c_void_p]),
    'ifcb_plan_num_launches': (C.c_int, [C.c_void_p]),
    'ifcb_plan_add_conv': (C.c_int, [C.c_void_p, C.POINTER(ConvDesc)]),
    'ifcb_conv_geometry': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'ifcb_conv_auto_config': (C.c_int, [C.c_int] * 10 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'ifcb_conv_auto_tile_n': (C.c_int, [C.c_int, C.c_int]),
    'ifcb_conv_wgrad': (C.c_int, [C.POINTER(WgradDesc), C.c_void_p]),
    'ifcb_train_deterministic': (C.c_int, [C.c_void_p, C.c_int64]),
    'ifcb_conv_wgrad_workspace_bytes': (C.c_int64, [C.POINTER(WgradDesc)]),
    'ifcb_memset_zero': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    'ifcb_bn_stats': (C.c_int, [_V, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_bn_apply': (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_void_p]),
    'ifcb_bn_apply_sums': (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    'ifcb_bn_backward': (C.c_int, [_V, _V, _V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_bn_backward_accumulate': (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_bias_relu_backward': (C.c_int, [_V, _V, _V, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_scale_elems': (C.c_int, [_V, _V, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_maxpool_fwd_train': (C.c_int, [_V, _V, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_maxpool_bwd': (C.c_int, [_V, C.c_void_p, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_avgpool_fwd': (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_avgpool_bwd': (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_dilate': (C.c_int, [_V, _V, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_nchw_to_nhwc': (C.c_int, [C.c_void_p, C.c_int, _V, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_stem_im2col': (C.c_int, [C.c_void_p, C.c_int, C.c_int, _V, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    'ifcb_head_train_fwd': (C.c_int, [_V, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_head_bwd': (C.c_int, [_V, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    'ifcb_dropout_scale': (C.c_int, [C.c_void_p, C.c_int64, C.c_float, C.c_uint64, C.c_void_p]),
    'ifcb_adam_step': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    'ifcb_conv_repack': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                   C.c_int, C.c_void_p]),
    'ifcb_conv_repack_batch': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'ifcb_plan_add_stem': (C.c_int, [C.c_void_p, C.POINTER(StemDesc)]),
    'ifcb_plan_add_pool': (C.c_int, [C.c_void_p, C.POINTER(PoolDesc)]),
    'ifcb_plan_add_head': (C.c_int, [C.c_void_p, C.POINTER(HeadDesc)]),
    'ifcb_debug_im2col_probe': (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Loads libifcb_b200.so (built in-tree by ifcb_classifier_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'ifcb_classifier_b200: %s is missing -- run `python -m ifcb_classifier_b200.build` '
                '(there is no CPU fallback)' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.ifcb_abi_version() != 1:
            raise RuntimeError('ifcb_classifier_b200: ABI version mismatch')
        _lib = l
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().ifcb_last_error().decode('utf-8', 'replace')
        raise RuntimeError('ifcb_b200 %s failed (status %d): %s' % (what, rc, msg))


def conv_geometry(Cin, Cout, kh, kw, tile_n=0):
    a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    check(lib().ifcb_conv_geometry(Cin, Cout, kh, kw, tile_n, C.byref(a), C.byref(b), C.byref(c), C.byref(d)),
          'conv_geometry')
    return dict(Cin_pad=a.value, K_pad=b.value, tile_n=c.value, Cout_pad=d.value)


def conv_auto_config(H, W, Cin, Cout, kh, kw, stride=(1, 1), pad=(0, 0)):
    """(algo, tile_n) the library prefers for this conv shape (IFCB_CONV_IM2COL / _WINDOW)."""
    a, t = C.c_int32(), C.c_int32()
    check(lib().ifcb_conv_auto_config(H, W, Cin, Cout, kh, kw, stride[0], stride[1], pad[0], pad[1], C.byref(a), C.byref(t)),
          'conv_auto_config')
    return a.value, t.value
