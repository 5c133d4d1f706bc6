"""ctypes binding of libhilcodec_b200.so (the C ABI in include/hilcodec_b200.h).

There is no CPU fallback: if the shared library is missing it is built on the spot with
nvcc, and if that is impossible the import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhilcodec_b200.so")

HIL_MAX_STRIDES = 8
HIL_ENCODER, HIL_DECODER = 0, 1
HIL_GRAPH_DEPLOY, HIL_GRAPH_TRAIN = 0, 1
PRE_NONE, PRE_ELU, PRE_SCALE_ELU = 0, 1, 2


class HilConfig(C.Structure):
    _fields_ = [
        ("channels_enc", C.c_int32), ("channels_dec", C.c_int32), ("n_fft_base", C.c_int32),
        ("n_residual_enc", C.c_int32), ("n_residual_dec", C.c_int32),
        ("res_scale_enc", C.c_double), ("res_scale_dec", C.c_double),
        ("n_strides", C.c_int32), ("strides", C.c_int32 * HIL_MAX_STRIDES),
        ("kernel_size", C.c_int32), ("dim", C.c_int32), ("codebook_size", C.c_int32),
        ("num_quantizers", C.c_int32),
    ]


class HilError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"hilcodec_b200 error {code}: {msg}")
        self.code = code


_P = C.c_void_p
_F = C.c_float
_I = C.c_int32

# name -> (restype, argtypes); every symbol include/hilcodec_b200.h declares
SIGNATURES = {
    "hil_abi_version": (_I, []),
    "hil_last_error": (C.c_char_p, []),
    "hil_config_default": (None, [C.POINTER(HilConfig), _I]),
    "hil_model_create": (_I, [C.POINTER(HilConfig), C.POINTER(_P)]),
    "hil_model_set_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "hil_model_set_graph": (_I, [_P, _I]),
    "hil_model_graph": (_I, [_P]),
    "hil_model_finalize": (_I, [_P]),
    "hil_model_destroy": (None, [_P]),
    "hil_model_hop": (_I, [_P]),
    "hil_model_num_caches": (_I, [_P, _I]),
    "hil_model_cache_shape": (_I, [_P, _I, _I, _I, C.POINTER(C.c_int64)]),
    "hil_state_create": (_I, [_P, _I, C.POINTER(_P)]),
    "hil_state_reset": (_I, [_P, _P]),
    "hil_state_export_cache": (_I, [_P, _I, _I, _P, _P]),
    "hil_state_import_cache": (_I, [_P, _I, _I, _P, _P]),
    "hil_state_destroy": (None, [_P]),
    "hil_state_workspace_bytes": (C.c_size_t, [_P]),
    "hil_state_range_flag": (_I, [_P, _I, _P, C.POINTER(_I)]),
    "hil_state_rollback": (_I, [_P, _I, _I]),
    "hil_set_exact_fp32": (_I, [_I]),
    "hil_encode": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "hil_encode_caches": (_I, [_P, _P, _P, _I, _I, _P, C.POINTER(_P), C.POINTER(_P), _P]),
    "hil_encode_ragged": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "hil_rvq_encode": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "hil_rvq_decode": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "hil_decode": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "hil_decode_caches": (_I, [_P, _P, _P, _I, _I, _P, C.POINTER(_P), C.POINTER(_P), _P]),
    "hil_codec_forward": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "hil_codec_forward_graph": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "hil_codec_forward_host": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "hil_bitstream_bytes_per_frame": (_I, [_P, _I]),
    "hil_pack_indices": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "hil_unpack_indices": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "hil_launch_count": (C.c_uint64, []),
    "hil_set_tensor_cores": (_I, [_I]),
    "hil_profile_begin": (_I, []),
    "hil_profile_end": (_I, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                             C.POINTER(C.c_int64), _I]),
    "hil_profile_launches": (_I, [C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), _I]),
    "hil_op_dwconv": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P]),
    "hil_op_dwconv_transpose": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "hil_op_pointwise": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "hil_op_dws": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _I, _F, _P]),
    "hil_op_upsample": (_I, [_P] * 8 + [_I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "hil_op_resblock": (_I, [_P] * 13 + [_I, _I, _I, _I, _F, _I, _P]),
    "hil_op_downsample": (_I, [_P] * 8 + [_I, _I, _I, _I, _I, _I, _F, _I, _P]),
    "hil_op_stft_logmag": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
}

_lock = threading.Lock()
_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the shared library and bind every signature."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise FileNotFoundError(LIB_PATH)
            from . import build as _build
            _build.build_library()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale: loud by design
            fn.restype = res
            fn.argtypes = args
        if lib.hil_abi_version() != 1:
            raise RuntimeError("libhilcodec_b200.so ABI version mismatch; rebuild with hilcodec_b200/build.py")
        _lib = lib
        return lib


def check(code: int) -> None:
    if code != 0:
        raise HilError(code, load().hil_last_error().decode("utf-8", "replace"))
