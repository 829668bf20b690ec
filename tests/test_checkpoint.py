"""CPU tests of the checkpoint readers (MindSpore .ckpt protobuf, .npz, name mapping)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from quanonet_b200 import checkpoint as ck


def _pb_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _pb_field(num, payload):
    return _pb_varint((num << 3) | 2) + _pb_varint(len(payload)) + payload


def _write_ms_ckpt(path, tensors):
    """Minimal writer of the MindSpore checkpoint wire format."""
    blob = b""
    for name, arr in tensors.items():
        arr = np.asarray(arr, dtype=np.float32)
        dims = b"".join(_pb_varint((1 << 3) | 0) + _pb_varint(d) for d in (arr.shape or (0,)))
        tensor = dims + _pb_field(2, b"Float32") + _pb_field(3, arr.astype("<f4").tobytes())
        blob += _pb_field(1, _pb_field(1, name.encode()) + _pb_field(2, tensor))
    with open(path, "wb") as f:
        f.write(blob)


def test_ckpt_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    src = {"bias": np.float32(0.25), "QuanONet.weight": rng.standard_normal(24).astype(np.float32),
           "branch_LinearLayer.Net2.weights": rng.standard_normal(4).astype(np.float32),
           "branch_LinearLayer.Net2.bias": rng.standard_normal(4).astype(np.float32),
           "trunk_LinearLayer.Net2.weights": rng.standard_normal(2).astype(np.float32),
           "trunk_LinearLayer.Net2.bias": rng.standard_normal(2).astype(np.float32)}
    p = str(tmp_path / "m.ckpt")
    _write_ms_ckpt(p, src)
    raw = ck.load_raw(p)
    assert raw["bias"].shape == () and float(raw["bias"]) == 0.25
    for k, v in src.items():
        assert np.array_equal(raw[k], v)
    sd = ck.ms_to_pt_arrays(raw, (2, 1, 1, 2), 2)
    assert sd["quantum_layer.ansatz_weights"].shape == (4, 3, 2)
    assert np.array_equal(sd["quantum_layer.ansatz_weights"].reshape(-1), src["QuanONet.weight"])
    assert sd["bias"].shape == (1,) and set(sd) == {"bias", "branch_freq.weights", "branch_freq.bias",
                                                   "trunk_freq.weights", "trunk_freq.bias",
                                                   "quantum_layer.ansatz_weights"}
    with pytest.raises(ValueError, match="expected"):
        ck.ms_to_pt_arrays(raw, (3, 1, 1, 2), 2)
    del raw["trunk_LinearLayer.Net2.bias"]
    with pytest.raises(KeyError):
        ck.ms_to_pt_arrays(raw, (2, 1, 1, 2), 2)


def test_corrupt_and_unknown_files(tmp_path):
    p = tmp_path / "bad.ckpt"
    p.write_bytes(b"\x0a\xff\xff\xff")
    with pytest.raises(ValueError):
        ck.load_raw(str(p))
    with pytest.raises(ValueError, match="extension"):
        ck.load_raw(str(tmp_path / "x.bin"))


def test_parse_experiment_dir():
    cfg = ck.parse_experiment_dir("pretrained_weights/Darcy/Darcy_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x25_Seed0/best_model.ckpt")
    assert cfg["net_size"] == (40, 2, 20, 2) and cfg["num_qubits"] == 5 and cfg["if_trainable_freq"]
    assert cfg["scale_coeff"] == 0.1 and cfg["num_points"] == 25 and cfg["operator"] == "Darcy"
    cfg = ck.parse_experiment_dir("x/Antideriv_HEAQNN_Net32-2_Q3_S0.01_100x10_Seed3")
    assert cfg["model_type"] == "HEAQNN" and "if_trainable_freq" not in cfg and cfg["net_size"] == (32, 2)
    with pytest.raises(ValueError):
        ck.parse_experiment_dir("nothing/here")


def test_parse_experiment_dir_reference_naming_scheme():
    """Every suffix the reference's utils/logger.py:55-118 can write: _TF/_FF, _Pauli*, _Diag*/_Ham*, the backend tag."""
    P = ck.parse_experiment_dir
    cfg = P("outputs/Advection/Advection_QuanONet_Net40-2-20-2_Q5_TF_S0.1_TQ_1000x100_Seed0/best_model.pt")
    assert cfg["quantum_backend"] == "torchquantum" and cfg["if_trainable_freq"] is True
    assert (cfg["num_train"], cfg["num_points"], cfg["seed"]) == (1000, 100, 0) and cfg["scale_coeff"] == 0.1
    cfg = P("Antideriv_QuanONet_Net5-1-5-1_Q2_FF_S0.001_1000x100_Seed0")
    assert cfg["if_trainable_freq"] is False and cfg["net_size"] == (5, 1, 5, 1) and cfg["scale_coeff"] == 0.001
    cfg = P("RDiffusion_HEAQNN_Net64-2_Q5_TF_S0.01_PauliX_1000x100_Seed3")
    assert cfg["ham_pauli"] == "X" and cfg["model_type"] == "HEAQNN" and cfg["net_size"] == (64, 2) and cfg["seed"] == 3
    cfg = P("Darcy_QuanONet_Net20-2-10-2_Q5_TF_S0.01_Ham-1-1_1000x25_Seed0")
    assert cfg["ham_bound"] == (-1.0, 1.0) and cfg["num_points"] == 25 and "ham_pauli" not in cfg
    cfg = P("Darcy_QuanONet_Net20-2-10-2_Q5_TF_S0.01_PauliY_Ham-10-10_Qiskit_1000x25_Seed0")
    assert cfg["ham_bound"] == (-10.0, 10.0) and cfg["ham_pauli"] == "Y" and cfg["quantum_backend"] == "qiskit"
    cfg = P("Antideriv_QuanONet_Net50-2-50-2_Q2_TF_S0.01_Diag-5--2.5-2.5-5_PL_1000x100_Seed1")
    assert cfg["ham_diag"] == [-5.0, -2.5, 2.5, 5.0] and cfg["quantum_backend"] == "pennylane"
    cfg = P("Antideriv_QuanONet_Net50-2-50-2_Q2_TF_S0.01_Diag-5-5-5-5_1000x100_Seed1")
    assert cfg["ham_diag"] == [-5.0, 5.0, 5.0, 5.0]


def test_infer_config_from_reference_style_names():
    from quanonet_b200.infer import _resolve_config
    cfg = _resolve_config("o/Advection_QuanONet_Net40-2-20-2_Q5_FF_S0.1_TQ_1000x100_Seed0/best_model.pt", {})
    assert cfg["model_type"] == "QuanONet" and cfg["if_trainable_freq"] is False
    cfg = _resolve_config("o/RDiffusion_HEAQNN_Net64-2_Q5_TF_S0.01_PauliX_Ham-1-1_1000x100_Seed3/final_model.npz", {})
    assert cfg["ham_pauli"] == "X" and tuple(cfg["ham_bound"]) == (-1.0, 1.0) and cfg["if_trainable_freq"] is True


def test_shipped_checkpoint_fixture_values():
    """First values recorded in SURVEY Appendix B for the shipped checkpoints."""
    z = np.load(os.path.join(GOLDEN, "pretrained.npz"))
    assert np.allclose(z["Advection/quantum_layer.ansatz_weights"].reshape(-1)[:3], [-1.78024, -0.18713, -3.07548], atol=1e-5)
    assert z["Advection/quantum_layer.ansatz_weights"].shape == (120, 3, 5)
    assert abs(float(z["Advection/bias"][0]) - 0.07399) < 1e-5
    assert abs(float(z["Darcy/bias"][0]) - 0.06565) < 1e-5
    assert abs(float(z["RDiffusion/bias"][0]) - 0.00752) < 1e-5
    assert abs(float(z["Antideriv/bias"][0]) - 0.032724) < 1e-6
    assert z["Antideriv/quantum_layer.ansatz_weights"].shape == (10, 3, 2)


@pytest.mark.skipif(not os.path.isdir("/root/reference/pretrained_weights"), reason="reference checkout absent")
def test_reads_the_reference_files_directly():
    z = np.load(os.path.join(GOLDEN, "pretrained.npz"))
    p = "/root/reference/pretrained_weights/RDiffusion/RDiffusion_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x100_Seed0/best_model.ckpt"
    sd = ck.ms_npz_to_pt_state_dict(p)
    assert np.array_equal(sd["quantum_layer.ansatz_weights"].numpy(), z["RDiffusion/quantum_layer.ansatz_weights"])
