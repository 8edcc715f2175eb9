"""ctypes binding of libaivc_b200.so (the C ABI declared in include/aivc_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C aivc_b200/csrc``).
If it is missing the product path fails loudly: there is no fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AIVC_B200_LIB: alternative build of the same ABI (A/B timing of kernel changes on one box)
LIB_PATH = os.environ.get('AIVC_B200_LIB') or os.path.join(_HERE, 'libaivc_b200.so')

F32, BF16, F16, BF16X2 = 0, 1, 2, 3     # BF16X2: split bf16 (hi | lo halves of a pixel), see aivc_b200.h
ACT = {'no': 0, 'none': 0, 'leaky_relu': 1, 'relu': 2, 'sigmoid': 3, 'gdn': 4, 'gdn_inverse': 5}
POST = {'none': 0, 'leaky_relu': 1, 'relu': 2, 'round_clamp': 3}
ENGINE_SIMT, ENGINE_TC, ENGINE_TC_X3 = 0, 1, 2


class FMap(C.Structure):
    _fields_ = [('data', C.c_void_p), ('h', C.c_int32), ('w', C.c_int32), ('c', C.c_int32),
                ('c_off', C.c_int32), ('c_stride', C.c_int32), ('pad', C.c_int32),
                ('pitch', C.c_int32), ('rows', C.c_int32), ('dtype', C.c_int32), ('_r', C.c_int32)]


class ConvOp(C.Structure):
    _fields_ = [('kind', C.c_int32), ('k', C.c_int32), ('stride', C.c_int32), ('engine', C.c_int32),
                ('inp', FMap), ('out', FMap),
                ('weight', C.c_void_p), ('bias', C.c_void_p),
                ('act', C.c_int32), ('post', C.c_int32),
                ('gdn_beta', C.c_void_p), ('gdn_gamma', C.c_void_p),
                ('residual', FMap), ('gate', FMap),
                ('out_scale', C.c_void_p), ('scratch', C.c_void_p),
                ('act_channels', C.c_int32), ('flags', C.c_int32),
                ('alg_flops', C.c_double)]


OP_LANE1, OP_FORK, OP_JOIN, OP_IN_EXACT = 1, 2, 4, 8
KERNEL_CLASSES = ('conv_simt_kernel', 'conv_tc_kernel', 'conv3x3_tc_kernel', 'conv3x3_tc_gdn_kernel',
                  'conv1x1_tc_kernel', 'col2im_tconv_kernel', 'space_to_depth_kernel', 'tconv3x3_tc_kernel')


_lib = None

_SIGS = {
    # name: (restype, argtypes)
    'aivc_abi_version': (C.c_int, []),
    'aivc_last_error': (C.c_char_p, []),
    'aivc_launch_count': (C.c_ulonglong, []),
    'aivc_profile_enable': (C.c_int, [C.c_int]),
    'aivc_profile_read': (C.c_int, [C.POINTER(C.c_double)]),
    'aivc_profile_dump': (C.c_int, [C.c_char_p]),
    'aivc_profile_read_classes': (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    'aivc_debug_tc3_tiling': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    'aivc_pack_conv_weight': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    'aivc_packed_weight_bytes': (C.c_size_t, [C.c_int] * 4),
    'aivc_conv2d_fused': (C.c_int, [C.POINTER(ConvOp), C.c_void_p]),
    'aivc_conv2d_fused_seq': (C.c_int, [C.POINTER(ConvOp), C.c_int, C.c_void_p]),
    'aivc_plan_graph_create': (C.c_int, [C.POINTER(ConvOp), C.c_int, C.POINTER(C.c_void_p)]),
    'aivc_plan_graph_launch': (C.c_int, [C.c_void_p, C.c_void_p]),
    'aivc_plan_graph_destroy': (C.c_int, [C.c_void_p]),
    'aivc_nchw_to_fmap': (C.c_int, [C.c_void_p, C.POINTER(FMap), C.c_void_p]),
    'aivc_fmap_to_nchw': (C.c_int, [C.POINTER(FMap), C.c_void_p, C.c_void_p]),
    'aivc_fill_border': (C.c_int, [C.POINTER(FMap), C.c_void_p]),
    'aivc_fmap_copy': (C.c_int, [C.POINTER(FMap), C.POINTER(FMap), C.c_void_p]),
    'aivc_yuv420_to_fmap': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.POINTER(FMap), C.c_void_p]),
    'aivc_yuv420_pack16': (C.c_int, [C.c_void_p] * 9 + [C.POINTER(FMap), C.POINTER(FMap), C.c_void_p]),
    'aivc_warp_blend': (C.c_int, [C.POINTER(FMap)] * 3 + [C.c_int, C.c_int] + [C.POINTER(FMap)] * 2
                        + [C.c_void_p, C.c_void_p]),
    'aivc_warp_blend_nchw': (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]),
    'aivc_finalize_frame': (C.c_int, [C.POINTER(FMap), C.POINTER(FMap), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.POINTER(FMap), C.c_void_p]),
    'aivc_mu_sigma_nchw': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    'aivc_quantize_latent': (C.c_int, [C.POINTER(FMap), C.POINTER(FMap), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(FMap), C.c_void_p, C.c_void_p]),
    'aivc_pdf_prob': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    'aivc_laplace_scale': (C.c_int, [C.POINTER(FMap), C.c_int, C.c_void_p, C.c_void_p]),
    'aivc_laplace_window': (C.c_int, [C.POINTER(FMap), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'aivc_dequantize_latent': (C.c_int, [C.c_void_p, C.POINTER(FMap), C.c_void_p, C.POINTER(FMap),
                                         C.c_void_p]),
    'aivc_fmap_to_i16': (C.c_int, [C.POINTER(FMap), C.c_void_p, C.c_void_p]),
    'aivc_i16_to_fmap': (C.c_int, [C.c_void_p, C.POINTER(FMap), C.c_void_p]),
    'aivc_frame_metrics_scratch_bytes': (C.c_size_t, [C.c_int, C.c_int]),
    'aivc_frame_metrics': (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    'aivc_rc_bound': (C.c_size_t, [C.c_size_t]),
    'aivc_rc_encode_bounds': (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.POINTER(C.c_size_t)]),
    'aivc_rc_encode_table': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_size_t)]),
    'aivc_rc_decode_table': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_size_t,
                                       C.c_void_p]),
    'aivc_rc_decode_laplace': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    'aivc_rc_decode_laplace_win': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                           C.c_void_p]),
    'aivc_laplace_cdf_int_host': (C.c_uint32, [C.c_float, C.c_int]),
    'aivc_sigma_from_logvar_host': (C.c_float, [C.c_float]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load (once) and return the library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'aivc_b200: %s is missing -- build it with `python -c "import __graft_entry__ as g; '
                'g.build()"` or `make -C aivc_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.aivc_abi_version() != 1:
            raise RuntimeError('aivc_b200: ABI version mismatch')
        _lib = L
    return _lib


PROFILING = False


def set_profiling(on):
    """Per-stage CUDA-event timing (aivc_profile_enable).  While it is on, plans run stage by stage instead of as
    CUDA graphs."""
    global PROFILING
    PROFILING = bool(on)
    lib().aivc_profile_enable(1 if on else 0)


def check(rc):
    if rc != 0:
        raise RuntimeError('aivc_b200: ' + lib().aivc_last_error().decode())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
