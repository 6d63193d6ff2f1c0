"""ctypes binding of libhm_b200.so (C ABI in include/hm_b200.h).

There is deliberately NO fallback: if the shared library is missing or an entry point is absent the
import raises, so a GPU box can never silently run a PyTorch/CPU path in place of the CUDA kernels.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhm_b200.so")

HM_ACT_NONE, HM_ACT_RELU, HM_ACT_LRELU, HM_ACT_TANH = 0, 1, 2, 3


class HmError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("c", C.c_int), ("cs", C.c_int), ("lo_c0", C.c_int)]


class OutF32(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int), ("h_off", C.c_int),
                ("w_off", C.c_int), ("c_off", C.c_int)]


class OutBF16(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("h_off", C.c_int), ("w_off", C.c_int), ("c_off", C.c_int)]


class SnLayer(C.Structure):
    _fields_ = [("W", C.c_void_p), ("grad", C.c_void_p), ("u", C.c_void_p), ("stash", C.c_void_p), ("n", C.c_int),
                ("m", C.c_int)]


_vp, _i, _f, _l, _sz = C.c_void_p, C.c_int, C.c_float, C.c_long, C.c_size_t
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/hm_b200.h declares must be listed here
# (tests/test_abi.py checks the header against this table and against the .so).
PROTOTYPES = {
    "hm_version": (C.c_char_p, []),
    "hm_last_cuda_error": (_i, []),
    "hm_scratch_bytes": (_sz, []),
    "hm_set_scratch": (_i, [_vp, _sz]),
    "hm_set_streamk": (_i, [_i]),
    "hm_set_sm_limit": (_i, [_i]),
    "hm_pick_bn": (_i, [_i]),
    "hm_rows_pad": (_i, [_i]),
    "hm_k_pad": (_i, [_i]),
    "hm_pack_weight": (_i, [_vp, _i, _i, _i, _l, _l, _l, _vp, _vp, _vp]),
    "hm_pack_weight_pair": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hm_conv_fprop": (_i, [_P(Operand), _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _P(OutF32),
                           _P(OutBF16), _vp, _vp]),
    "hm_conv_dgrad": (_i, [_P(Operand), _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _f, _P(OutF32),
                           _P(OutBF16), _vp, _vp]),
    "hm_conv_fprop_dil": (_i, [_P(Operand), _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _P(OutF32),
                               _P(OutBF16), _vp, _vp]),
    "hm_conv_dgrad_dil": (_i, [_P(Operand), _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _P(OutF32),
                               _P(OutBF16), _vp, _vp]),
    "hm_conv_wgrad_dil": (_i, [_P(Operand), _P(Operand), _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hm_wgrad_ws_bytes": (_sz, [_i, _i, _i, _i]),
    "hm_conv_wgrad": (_i, [_P(Operand), _P(Operand), _i, _i, _i, _i, _vp, _vp, _vp]),
    "hm_wgrad_unpack": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "hm_pack_weight_ex": (_i, [_vp, _i, _i, _l, _l, _i, _i, _l, _l, _i, _l, _vp, _vp, _vp, _vp]),
    "hm_wgrad_unpack_cols": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "hm_tap_unroll": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp]),
    "hm_tap_combine": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _f, _vp, _i, _i, _i, _vp]),
    "hm_encode_input": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _vp,
                             _vp]),
    "hm_in_ws_bytes": (_sz, [_i, _i, _i]),
    "hm_in_stats": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp]),
    "hm_in_apply": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "hm_in_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _i, _i, _i, _i, _i, _f, _vp, _vp,
                       _vp, _i, _vp, _vp]),
    "hm_fold_add": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hm_avgpool3s2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "hm_avgpool3s2_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "hm_maxpool2": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hm_maxpool2_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "hm_l1_sum": (_i, [_vp, _vp, _l, C.c_double, _vp, _vp]),
    "hm_mse_sum": (_i, [_vp, _l, _f, C.c_double, _vp, _vp]),
    "hm_mse_grad": (_i, [_vp, _l, _i, _f, _f, _vp, _vp, _i, _vp]),
    "hm_bce_sum": (_i, [_vp, _l, _f, C.c_double, _vp, _vp]),
    "hm_bce_grad": (_i, [_vp, _l, _i, _f, _f, _vp, _vp, _i, _vp]),
    "hm_finish_fake": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "hm_fake_bwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _i, _vp, _f, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "hm_f32_to_operand": (_i, [_vp, _l, _i, _i, _i, _f, _vp, _vp, _i, _vp]),
    "hm_colsum": (_i, [_vp, _l, _i, _vp, _i, _vp]),
    "hm_colsum_operand": (_i, [_vp, _vp, _l, _i, _i, _vp, _i, _vp]),
    "hm_mask_maxpool": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "hm_mask_blend": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "hm_mask_blend_bwd": (_i, [_vp, _vp, _l, _i, _vp, _vp, _vp]),
    "hm_pool_exchange": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _l, _vp]),
    "hm_mask_concat": (_i, [_vp, _vp, _vp, _l, _i, _vp, _vp, _i, _vp]),
    "hm_mask_concat_bwd": (_i, [_vp, _i, _vp, _vp, _vp, _l, _i, _vp, _vp, _vp]),
    "hm_concat_operands": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _l, _vp]),
    "hm_cond_image_operand": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "hm_sn_stash_floats": (_sz, [_i, _i]),
    "hm_sn_power_iteration": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "hm_sn_weight_grad": (_i, [_vp, _i, _i, _i, _vp]),
    "hm_box2mask_encode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "hm_bn_fold": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _i, _vp]),
    "hm_bn_stats": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp]),
    "hm_upsample2_add": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "hm_box2mask_head": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hm_bn_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _vp, _vp,
                       _vp, _vp]),
    "hm_upsample2_bwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "hm_box2mask_head_bwd": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _f, _f, _vp, _i, _vp, _vp, _i, _vp,
                                  _vp, _i, _vp]),
    "hm_box2mask_d_input": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "hm_adam_step": (_i, [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _i, _f, _vp]),
    "hm_adam_step_dev": (_i, [_vp, _vp, _vp, _vp, _l, _f, _f, _f, _f, _vp, _f, _vp]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library handle; raises HmError when it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HmError("libhm_b200.so is not built (%s). Run `python -m neurips18_hierchical_image_manipulation_b200.build`; "
                      "there is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise HmError("libhm_b200.so lacks symbol %s" % name) from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        lib = load()
        raise HmError("%s failed: hm_status=%d (cudaError=%d)" % (what, rc, lib.hm_last_cuda_error()))
