"""CPU test: the C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(qon_\w+)\s*\(", text))
    return sorted(syms)


@pytest.fixture(scope="module")
def lib():
    from quanonet_b200 import _lib
    from quanonet_b200.build import build
    build()                      # nvcc cross-compiles without a GPU; no-op when up to date
    return _lib.load()


def test_header_declares_the_documented_entry_points():
    syms = _declared_symbols()
    for s in ("qon_hea_forward", "qon_hea_forward_backward", "qon_workspace_bytes", "qon_last_error",
              "qon_abi_version"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    from quanonet_b200 import _lib
    for s in _declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) == set(_declared_symbols())
    assert lib.qon_abi_version() == _lib.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu(lib):
    """Host-side validation runs before any CUDA call that needs a device."""
    depth = (ctypes.c_int * 2)(1, 0)
    assert lib.qon_workspace_bytes(10, 5, 2, depth, 0, 1) == 0
    assert b"depth_per_block" in lib.qon_last_error()
    depth = (ctypes.c_int * 2)(1, 1)
    assert lib.qon_workspace_bytes(10, 99, 2, depth, 0, 1) == 0
    assert b"n must be" in lib.qon_last_error()
    assert lib.qon_workspace_bytes(10, 5, 2, depth, 7, 1) == 0
    rc = lib.qon_hea_forward(None, 10, None, None, 4, 5, 2, depth, None, 0, 0.0, 1.0, 0, 3, None, 0, None)
    assert rc < 0


def test_host_only_entry_points(lib):
    """Planning helpers and the host-side checks of the exchange entry points need no device."""
    assert lib.qon_latency_tier_max_batch() >= 0
    assert lib.qon_peer_buffer_bytes(2402, 8) == 256 + 2 * 8 * 2402 * 4
    assert lib.qon_peer_buffer_bytes(2402, 9) == 0            # one node: at most 8 ranks
    bufs = (ctypes.c_void_p * 2)(None, None)
    assert lib.qon_peer_allreduce_f32(None, None, 10, bufs, 2, 0, 10, None) < 0
    assert b"non-NULL" in lib.qon_last_error()
    assert lib.qon_peer_allreduce_f32(ctypes.c_void_p(256), ctypes.c_void_p(256), 10, bufs, 2, 5, 10, None) < 0
    assert b"rank" in lib.qon_last_error()
