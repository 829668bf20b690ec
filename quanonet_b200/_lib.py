"""ctypes binding of libquanonet_b200.so (the C-ABI declared in include/quanonet_b200.h).

There is NO CPU fallback: if the library is missing or fails to load, every compute entry point
raises.  Build it with ``python -m quanonet_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# QON_LIB_PATH selects an experiment build of the same library (scripts/build_variant.sh); never a fallback
LIB_PATH = os.environ.get("QON_LIB_PATH") or os.path.join(PKG_DIR, "libquanonet_b200.so")

QON_F32, QON_F64 = 0, 1
QON_HAM_DIAG, QON_HAM_PAULI_X, QON_HAM_PAULI_Y = 0, 1, 2
QON_DIAG_LSB0, QON_DIAG_MSB0 = 0, 1
ABI_VERSION = 3

# every symbol include/quanonet_b200.h declares
EXPORTED_SYMBOLS = (
    "qon_abi_version",
    "qon_last_error",
    "qon_workspace_bytes",
    "qon_hea_forward",
    "qon_hea_forward_backward",
    "qon_hea_mse_forward_backward",
    "qon_encoded_forward",
    "qon_encoded_mse_step",
    "qon_plan_tier",
    "qon_latency_tier_max_batch",
    "qon_encoded_supported",
    "qon_encoded_supported_for",
    "qon_peer_buffer_bytes",
    "qon_peer_allreduce_f32",
    "qon_encoded_mse_step_dp",
    "qon_measure_fp32_peak_tflops",
    "qon_tensor_tier",
)

_lock = threading.Lock()
_lib = None


class QonLibraryError(RuntimeError):
    pass


def _declare(lib):
    c = ctypes
    vp, i64, i32, dbl, sz = c.c_void_p, c.c_int64, c.c_int, c.c_double, c.c_size_t
    ip = c.POINTER(c.c_int)
    lib.qon_abi_version.restype = i32
    lib.qon_abi_version.argtypes = []
    lib.qon_last_error.restype = c.c_char_p
    lib.qon_last_error.argtypes = []
    lib.qon_workspace_bytes.restype = sz
    lib.qon_workspace_bytes.argtypes = [i64, i32, i32, ip, i32, i32]
    lib.qon_hea_forward.restype = i32
    lib.qon_hea_forward.argtypes = [vp, i64, vp, vp, i64, i32, i32, ip, vp, i32, dbl, dbl, i32, i32, vp, sz, vp]
    lib.qon_hea_forward_backward.restype = i32
    lib.qon_hea_forward_backward.argtypes = [vp, i64, vp, vp, vp, vp, i64, vp, i64, i32, i32, ip,
                                             vp, i32, dbl, dbl, i32, i32, vp, sz, vp]
    lib.qon_hea_mse_forward_backward.restype = i32
    lib.qon_hea_mse_forward_backward.argtypes = [vp, i64, vp, vp, vp, dbl, vp, vp, vp, i64, vp, i64, i32, i32, ip,
                                                 vp, i32, dbl, dbl, i32, i32, vp, sz, vp]
    ham_tail = [vp, i32, dbl, dbl, i32, i32, vp, sz, vp]   # ham_diag, diag_order, offset, coeff, kind, dtype, ws, bytes, stream
    enc_head = [vp, i64, i32, i32, vp, i64, i32, vp, vp]   # u0, ldu0, in0, K0, u1, ldu1, in1, fw, fb
    lib.qon_encoded_forward.restype = i32
    lib.qon_encoded_forward.argtypes = enc_head + [vp, vp, i64, i32, i32, ip] + ham_tail
    lib.qon_encoded_mse_step.restype = i32
    lib.qon_encoded_mse_step.argtypes = enc_head + [vp, vp, vp, dbl, vp, vp, vp, vp, vp, i64, i32, i32, ip] + ham_tail
    lib.qon_encoded_supported.restype = i32
    lib.qon_encoded_supported.argtypes = [i64, i32, i32, i32]
    lib.qon_encoded_supported_for.restype = i32
    lib.qon_encoded_supported_for.argtypes = [i64, i32, i32, ip, i32, i32]
    lib.qon_latency_tier_max_batch.restype = i64
    lib.qon_latency_tier_max_batch.argtypes = []
    lib.qon_peer_buffer_bytes.restype = sz
    lib.qon_peer_buffer_bytes.argtypes = [i64, i32]
    lib.qon_peer_allreduce_f32.restype = i32
    lib.qon_peer_allreduce_f32.argtypes = [vp, vp, i64, c.POINTER(c.c_void_p), i32, i32, i64, vp]
    lib.qon_encoded_mse_step_dp.restype = i32
    lib.qon_encoded_mse_step_dp.argtypes = [vp, i64, i32, i32, vp, i64, i32, vp, vp, vp, vp, vp, dbl, vp,
                                            vp, i64, i64, i64, i64, i64, c.POINTER(c.c_void_p), i32, i32, i64,
                                            i64, i32, i32, ip, vp, i32, dbl, dbl, i32, vp, sz, vp]
    lib.qon_plan_tier.restype = i32
    lib.qon_plan_tier.argtypes = [i64, i32, i32, i32, ip]
    lib.qon_tensor_tier.restype = i32
    lib.qon_tensor_tier.argtypes = [i32, i64, vp, vp]
    lib.qon_measure_fp32_peak_tflops.restype = dbl
    lib.qon_measure_fp32_peak_tflops.argtypes = [i32, vp]


def load():
    """Return the loaded library, loading it on first use.  Raises QonLibraryError if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise QonLibraryError(
                f"{LIB_PATH} not found: the CUDA extension has not been built. "
                "Run `python -m quanonet_b200.build` (needs nvcc). There is no CPU fallback.")
        try:
            lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        except OSError as e:  # pragma: no cover - depends on the host
            raise QonLibraryError(f"failed to load {LIB_PATH}: {e}") from e
        _declare(lib)
        got = lib.qon_abi_version()
        if got != ABI_VERSION:
            raise QonLibraryError(f"ABI version mismatch: library {got}, binding {ABI_VERSION}; rebuild")
        _lib = lib
        return lib


def last_error() -> str:
    return load().qon_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        kind = "invalid argument" if rc < 0 else "CUDA error"
        raise RuntimeError(f"{what} failed ({kind} {rc}): {last_error()}")


def int_array(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr
