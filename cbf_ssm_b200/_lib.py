"""ctypes binding of ``libcbfssm_b200.so`` (C ABI: ``include/cbfssm_b200.h``).

The product path has no CPU fallback: if the shared library is missing, or the
process has no CUDA device when a compute entry point is called, this module
raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C cbf_ssm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CBFSSM_B200_LIB") or os.path.join(_HERE, "libcbfssm_b200.so")   # override: kernel experiments

CBF_ERR = {-1: "CBF_ERR_INVALID_SHAPE", -2: "CBF_ERR_UNSUPPORTED_DIMS", -3: "CBF_ERR_UNSUPPORTED_M",
           -4: "CBF_ERR_ALIGNMENT", -5: "CBF_ERR_NULL"}


class CbfError(RuntimeError):
    def __init__(self, code, msg):
        self.code = code
        super().__init__(f"cbfssm_b200: {CBF_ERR.get(code, 'cudaError %d' % code)}: {msg}")


class cbf_shape(C.Structure):
    _fields_ = [("B", C.c_int32), ("S", C.c_int32), ("T", C.c_int32), ("M", C.c_int32),
                ("dx", C.c_int32), ("du", C.c_int32), ("dy", C.c_int32), ("R", C.c_int32),
                ("condition", C.c_int32), ("n_offset", C.c_int32), ("n_local", C.c_int32),
                ("k_factor", C.c_float), ("flags", C.c_int32)]


class cbf_gp(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("Z", "ell", "sig2", "P", "alpha", "S", "state")]


class cbf_grad_layout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "f_P", "f_alpha", "f_S", "f_Z", "f_ell", "f_sig2",
        "b_P", "b_alpha", "b_S", "b_Z", "b_ell", "b_sig2", "var_x", "var_y", "total")]


_P = C.c_void_p
_SIGNATURES = {
    "cbf_abi_version": (C.c_int, []),
    "cbf_last_error_string": (C.c_char_p, []),
    "cbf_supported": (C.c_int, [C.c_int32] * 4),
    "cbf_workspace_bytes": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(C.c_size_t)]),
    "cbf_grad_layout_get": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(cbf_grad_layout)]),
    "cbf_elbo_forward": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(cbf_gp), C.POINTER(cbf_gp)] + [_P] * 10),
    "cbf_elbo_backward": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(cbf_gp), C.POINTER(cbf_gp)] + [_P] * 7
                          + [C.POINTER(C.c_double), _P, _P, _P]),
    "cbf_elbo_forward_half": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(cbf_gp)] + [_P] * 9),
    "cbf_elbo_backward_half": (C.c_int, [C.POINTER(cbf_shape), C.POINTER(cbf_gp)] + [_P] * 6
                               + [C.POINTER(C.c_double), _P, _P, _P, _P]),
    "cbf_export_states": (C.c_int, [C.POINTER(cbf_shape), _P, _P, _P, _P, _P]),
    "cbf_state_sums": (C.c_int, [C.POINTER(cbf_shape), _P, _P, _P]),
    "cbf_moments": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "cbf_gp_prologue_state_doubles": (C.c_int64, [C.c_int32] * 3),
    "cbf_gp_prologue": (C.c_int, [C.c_int32] * 3 + [_P] * 14),
    "cbf_gp_prologue_backward": (C.c_int, [C.c_int32] * 3 + [_P] * 6 + [C.c_double] + [_P] * 7),
    "cbf_noise_forward": (C.c_int, [C.c_int32] + [_P] * 5),
    "cbf_noise_backward": (C.c_int, [C.c_int32] + [_P] * 7),
    "cbf_adam_step": (C.c_int, [C.c_int64, _P, _P, _P, _P, C.c_int64, C.c_double, C.c_double, C.c_double,
                                C.c_double, _P]),
    "cbf_fill_normal": (C.c_int, [_P, C.c_int64, C.c_uint64, C.c_uint64, _P]),
    "cbf_timing_enable": (C.c_int, [C.c_int]),
    "cbf_timing_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "cbf_launches_read": (C.c_int, [C.POINTER(C.c_int64), C.c_int]),
    "cbf_measure_fp32_peak": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library once; raise (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
            "`python -c \"import __graft_entry__ as g; g.build()\"` (there is no CPU fallback).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.cbf_abi_version() != 2:
        raise ImportError("libcbfssm_b200.so has an unexpected ABI version")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CbfError(rc, load().cbf_last_error_string().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())
