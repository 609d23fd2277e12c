"""ctypes binding of libxde_b200.so (include/xde_b200.h).

There is NO CPU fallback: if the CUDA library is missing or a field is not covered by a fused kernel
the call raises.  PyTorch is used only as plumbing (device memory, streams, DLPack interchange)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libxde_b200.so")

XDE_OK, XDE_E_BAD_ARG, XDE_E_UNSUPPORTED_FIELD, XDE_E_CUDA = 0, -1, -2, -3
ST_OK, ST_DT_UNDERFLOW, ST_NONFINITE_STATE, ST_MAX_STEPS, ST_INTERP_RANGE, ST_TC_RANGE = 0, 1, 2, 3, 5, 6

PRE = {"id": 0, "identity": 0, None: 0, "square": 1, "cube": 2}
CTRL = {"trajectory": 0, "batch": 1}
ADJ_NORM = {"mixed": 0, "default": 0, "seminorm": 1}
FIXED = {"euler": 0, "rk4": 1, "midpoint": 2}
RK = {"dopri5": 0, "bosh3": 1, "fehlberg2": 2, "adaptive_heun": 3, "dopri8": 4, "dopri5_table": 100}
SDE = {"em": 0, "euler": 0, "milstein": 1}
INTERP = {"linear": 0, "cubic": 1, "hermite": 1, "bez": 2, "bezier": 2}


class XdeError(RuntimeError):
    """CUDA/launch failure inside libxde_b200."""


class UnsupportedFieldError(NotImplementedError):
    """The vector field / shape / option has no fused sm_100a kernel.  By design there is no fallback."""


class MlpFieldC(C.Structure):
    _fields_ = [("d", C.c_int32), ("h", C.c_int32), ("pre", C.c_int32), ("_pad", C.c_int32),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


class CtrlOptsC(C.Structure):
    _fields_ = [("rtol", C.c_float), ("atol", C.c_float), ("min_step", C.c_float), ("max_step", C.c_float),
                ("first_step", C.c_float), ("safety", C.c_float), ("ifactor", C.c_float),
                ("dfactor", C.c_float), ("max_num_steps", C.c_int32), ("_pad", C.c_int32)]


class StatsC(C.Structure):
    _fields_ = [("n_attempts", C.c_ulonglong), ("n_accepted", C.c_ulonglong), ("nfe", C.c_ulonglong),
                ("status", C.c_int32), ("_pad", C.c_int32)]


class AttemptLogC(C.Structure):
    _fields_ = [("records", C.c_void_p), ("counts", C.c_void_p), ("cap", C.c_int32), ("_pad", C.c_int32)]


_EXPORTS = {
    "xde_abi_version": (C.c_int, []),
    "xde_last_error": (C.c_char_p, []),
    "xde_launch_count": (C.c_ulonglong, []),
    "xde_default_ctrl_opts": (None, [C.POINTER(CtrlOptsC)]),
    "xde_probe_ffma_f32": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "xde_probe_ffma2_f32": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]),
    "xde_dopri5_mlp_f32": (C.c_int, [C.POINTER(MlpFieldC), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                     C.POINTER(CtrlOptsC), C.c_int32, C.c_void_p, C.c_void_p,
                                     C.POINTER(AttemptLogC), C.c_void_p]),
    "xde_adaptive_rk_mlp_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int32, C.POINTER(CtrlOptsC), C.c_int32, C.c_void_p, C.c_void_p,
                                          C.POINTER(AttemptLogC), C.c_void_p]),
    "xde_adaptive_rk_mlp_grid_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.c_void_p, C.c_int64, C.c_void_p,
                                               C.c_int32, C.POINTER(CtrlOptsC), C.c_int32, C.c_void_p, C.c_int32,
                                               C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                               C.POINTER(AttemptLogC), C.c_void_p]),
    "xde_dopri5_mlp_adjoint_f32": (C.c_int, [C.POINTER(MlpFieldC), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_int64, C.POINTER(CtrlOptsC), C.c_int32, C.c_int32, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(AttemptLogC),
                                             C.c_void_p]),
    "xde_rk_fixed_mlp_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "xde_fixed_interp_linear_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                              C.c_void_p, C.c_void_p]),
    "xde_sde_mlp_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.POINTER(MlpFieldC), C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "xde_rk_fixed_mlp_tc_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xde_sde_mlp_tc_f32": (C.c_int, [C.c_int32, C.POINTER(MlpFieldC), C.POINTER(MlpFieldC), C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "xde_brownian_increments_f32": (C.c_int, [C.c_uint64, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                              C.c_void_p, C.c_void_p]),
    "xde_sde_mlp_philox_f32": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(MlpFieldC), C.POINTER(MlpFieldC), C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_int32, C.c_uint64, C.c_int64, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "xde_sde_mlp_adjoint_f32": (C.c_int, [C.POINTER(MlpFieldC), C.POINTER(MlpFieldC), C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "xde_history_gather_f32": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "xde_history_gather_bwd_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                             C.c_void_p, C.c_void_p]),
    "xde_dde_fuse_f32": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "xde_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "xde_set_device": (C.c_int, [C.c_int32]),
    "xde_get_device": (C.c_int, [C.POINTER(C.c_int32)]),
    "xde_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64, C.c_void_p]),
    "xde_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "xde_memcpy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p]),
    "xde_memset_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p]),
    "xde_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "xde_host_free": (C.c_int, [C.c_void_p]),
    "xde_stream_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "xde_stream_destroy": (C.c_int, [C.c_void_p]),
    "xde_stream_synchronize": (C.c_int, [C.c_void_p]),
    "xde_dlpack_wrap": (C.c_void_p, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32]),
    "xde_dlpack_release": (None, [C.c_void_p]),
    "xde_allreduce_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "xde_dde_fuse_bwd_f32": (C.c_int, [C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def library_path() -> str:
    return _SO


def lib():
    """Load libxde_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(
                f"{_SO} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C paddlexde_b200/csrc`). paddlexde_b200 has no CPU fallback.")
        handle = C.CDLL(_SO)
        for name, (res, args) in _EXPORTS.items():
            fn = getattr(handle, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if handle.xde_abi_version() != 3:
            raise ImportError("libxde_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def exported_symbols():
    return list(_EXPORTS)


def check(rc: int):
    if rc == XDE_OK:
        return
    msg = lib().xde_last_error().decode("utf-8", "replace")
    if rc == XDE_E_BAD_ARG:
        raise ValueError(msg)
    if rc == XDE_E_UNSUPPORTED_FIELD:
        raise UnsupportedFieldError(msg)
    raise XdeError(msg)


def raise_for_status(status: int):
    """Re-raise the reference's Python asserts (solver/base_adaptive_solver_rk.py:120-122,200-203,
    utils/ode_utils.py:65-67) from the device status word."""
    if status == ST_OK:
        return
    text = {ST_DT_UNDERFLOW: "underflow in dt", ST_NONFINITE_STATE: "non-finite values in state `y`",
            ST_MAX_STEPS: "max_num_steps exceeded", ST_INTERP_RANGE: "invalid interpolation, fails `t0 <= t <= t1`"}
    raise AssertionError(text.get(status, f"solver status {status}"))


def launch_count() -> int:
    return int(lib().xde_launch_count())
