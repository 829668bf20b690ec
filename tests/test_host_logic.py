"""CPU tests of the host-side mirror of the reference interface (no compute calls)."""
import numpy as np
import pytest
import torch

from quanonet_b200.core import quantum_circuits_tq as qtq
from quanonet_b200.core.models_pt import HEAQNNPT, QuanONetPT, _ScaleRepeat, _TiledElementWise


def test_public_surface_matches_reference_module():
    # names the reference's callers import (core/models_pt.py:75, core/quantum_circuits_pl.py:17, ibm_inference.py:18)
    for name in ("build_quanonet_tq", "build_heaqnn_tq", "_make_block_configs", "_ham_params", "_TQHEACircuit"):
        assert hasattr(qtq, name)
    assert qtq._make_block_configs(5, 20, 2, 40, 2) == [(5, 2)] * 20 + [(5, 2)] * 40
    assert qtq._make_block_configs(3, 1, 4, 2, 1) == [(3, 4), (3, 1), (3, 1)]          # trunk first
    assert qtq._ham_params(5, -5.0, 5.0) == (0.0, 1.0)
    off, co = qtq._ham_params(2, -5.0, 5.0)
    assert (off, co) == (0.0, 2.5)
    assert qtq._ham_params(4, -1.0, 3.0) == (1.0, 0.5)


def test_module_attributes_and_state_dict():
    torch.manual_seed(0)
    m = qtq.build_quanonet_tq(5, 100, 2, (40, 2, 20, 2))
    assert m.n_wires == 5 and len(m.block_configs) == 60 and not m.use_full_ham
    assert m.ham_offset == 0.0 and m.ham_coeff == 1.0
    assert tuple(m.ansatz_weights.shape) == (120, 3, 5)
    assert float(m.ansatz_weights.detach().abs().max()) <= np.pi
    assert list(m.state_dict().keys()) == ["ansatz_weights"]
    d = qtq.build_heaqnn_tq(2, 6, (2, 1, 0, 0), ham_diag=[-5.0, -2.5, 2.5, 5.0])
    assert d.use_full_ham and "ham_diag" in d.state_dict() and d.block_configs == [(2, 1)] * 2
    with pytest.raises(ValueError):
        qtq.build_quanonet_tq(2, 1, 1, (1, 1, 1, 1), ham_diag=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        qtq._TQHEACircuit(2, [(2, 1)], ham_pauli="Q")


def test_model_parameter_names_and_counts_match_reference():
    m = QuanONetPT(5, 100, 2, (40, 2, 20, 2), scale_coeff=0.1, if_trainable_freq=True)
    sd = m.state_dict()
    assert set(sd) == {"bias", "branch_freq.weights", "branch_freq.bias", "trunk_freq.weights", "trunk_freq.bias",
                       "quantum_layer.ansatz_weights"}
    assert sum(p.numel() for p in m.parameters()) == 2401          # 1800 + 2*200 + 2*100 + 1
    assert float(m.branch_freq.weights[0]) == pytest.approx(0.1) and float(m.branch_freq.bias.abs().max()) == 0.0
    h = HEAQNNPT(2, 6, (2, 1, 0, 0), scale_coeff=0.1, if_trainable_freq=True)
    assert set(h.state_dict()) == {"freq.weights", "freq.bias", "quantum_layer.ansatz_weights"}
    f = QuanONetPT(3, 7, 4, (3, 2, 1, 1), scale_coeff=0.7, if_trainable_freq=False)
    assert set(f.state_dict()) == {"bias", "quantum_layer.ansatz_weights"}
    with pytest.raises(ValueError):
        QuanONetPT(2, 3, 1, (1, 1, 1, 1), quantum_backend="qiskit")


def test_frequency_layers_tile_like_the_reference():
    x = torch.arange(6.0).reshape(2, 3)
    t = _TiledElementWise(3, 7, 0.5)
    with torch.no_grad():
        t.bias.copy_(torch.arange(7.0))
    exp = torch.stack([x[:, j % 3] * 0.5 + j for j in range(7)], dim=1)
    assert torch.allclose(t(x), exp)
    s = _ScaleRepeat(3, 2, 2.0)                                   # in > out: extra inputs dropped
    assert torch.equal(s(x), x[:, :2] * 2.0)


def test_canonical_plan_standard_ragged_and_merged_blocks():
    plan = qtq._canonical_plan
    depths, groups, ident = plan(3, [(3, 2), (3, 1)], 6)
    assert depths == [2, 1] and ident and groups == [[0], [1], [2], [3], [4], [5]]
    # ragged: x has only 4 columns, later gates are skipped (quantum_circuits_tq.py:83)
    depths, groups, ident = plan(3, [(3, 1), (3, 1)], 4)
    assert not ident and groups == [[0], [1], [2], [3], [], []]
    # n_encode > n: two angles land on wire 0 and add; n_encode < n: wire 2 gets none
    depths, groups, ident = plan(3, [(4, 1), (2, 2)], 6)
    assert depths == [1, 2] and groups == [[0, 3], [1], [2], [4], [5], []]
    # a depth-0 block merges its encoding into the next block
    depths, groups, ident = plan(2, [(2, 0), (2, 3)], 4)
    assert depths == [3] and groups == [[0, 2], [1, 3]]
    with pytest.raises(NotImplementedError):
        plan(2, [(2, 1), (2, 0)], 4)
    # extra x columns beyond what the circuit consumes are ignored
    depths, groups, ident = plan(2, [(2, 1)], 5)
    assert not ident and groups == [[0], [1]]


def test_canonical_inputs_gather_sums_columns():
    m = qtq._TQHEACircuit(3, [(4, 1), (2, 2)])
    x = torch.arange(12.0).reshape(2, 6)
    xc, depths = m.canonical_inputs(x)
    assert depths == [1, 2] and xc.shape == (2, 6)
    assert torch.equal(xc[:, 0], x[:, 0] + x[:, 3]) and torch.equal(xc[:, 5], torch.zeros(2))


def test_cpu_tensors_fail_loudly_no_fallback():
    m = QuanONetPT(2, 3, 1, (1, 1, 1, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(4, 3), torch.zeros(4, 1))


def test_backend_routing():
    from quanonet_b200.utils.backend import backend
    assert backend.check_compatibility("QuanONet", "torchquantum") == "pytorch_quantum"
    assert backend.check_compatibility("HEAQNN", "torchquantum") == "pytorch_quantum"
    with pytest.raises(ImportError):
        backend.check_compatibility("QuanONet", "mindquantum")
    with pytest.raises(ValueError):
        backend.check_compatibility("QuanONet", "nope")
    assert backend.check_compatibility("Whatever") == "unknown"


def test_product_code_never_imports_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "quanonet_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_encoded_supported_static_rules():
    """The size-independent part of the fused-encoding rule needs no device (the per-batch part asks the planner)."""
    import torch
    from quanonet_b200.ops import encoded_supported
    assert encoded_supported(5, torch.float32) and encoded_supported(4, torch.float64)
    assert not encoded_supported(5, torch.float64) and not encoded_supported(7, torch.float32)
    assert not encoded_supported(7, torch.float64, batch=100) and not encoded_supported(10, torch.float32, batch=100)
