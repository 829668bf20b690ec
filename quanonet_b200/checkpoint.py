"""Checkpoint readers for the weights the reference ships and writes.

* MindSpore ``.npz`` (written by the reference's MindSpore solver, ``solvers/solver_ms.py:254-263``)
  and PyTorch ``.npz`` (``solvers/solver_pt.py:250-257``);
* MindSpore ``.ckpt`` — a protobuf file, decoded here with a ~40-line wire-format reader so the
  three shipped Q5 checkpoints (``pretrained_weights/{Advection,Darcy,RDiffusion}/…/best_model.ckpt``)
  load without MindSpore.

Key mapping MindSpore → PyTorch state_dict follows ``utils/weight_transfer.py:37-43,73-96``:
the flat circuit-order vector ``QuanONet.weight (3nS,)`` reshapes directly to
``quantum_layer.ansatz_weights (S,3,n)``.
"""
from __future__ import annotations

import os
import re
from typing import Dict, Tuple

import numpy as np

_MS_FREQ_KEYS = {
    "branch_LinearLayer.Net2.weights": "branch_freq.weights",
    "branch_LinearLayer.Net2.bias": "branch_freq.bias",
    "trunk_LinearLayer.Net2.weights": "trunk_freq.weights",
    "trunk_LinearLayer.Net2.bias": "trunk_freq.bias",
    # HEAQNN (core/models_ms.py:92-124) keeps a single frequency layer
    "LinearLayer.Net2.weights": "freq.weights",
    "LinearLayer.Net2.bias": "freq.bias",
}
_MS_CIRCUIT_KEYS = ("QuanONet.weight", "HEAQNN.weight")

_MS_DTYPES = {
    "Float32": np.float32,
    "Float64": np.float64,
    "Float16": np.float16,
    "Int32": np.int32,
    "Int64": np.int64,
}


# ------------------------------------------------------------------ protobuf wire format


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _fields(buf: bytes):
    """Yield ``(field_number, wire_type, value)``; value is int for varint, bytes for
    length-delimited.  Only wire types 0/1/2/5 occur."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise ValueError("truncated length-delimited field")
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def read_mindspore_ckpt(path: str) -> Dict[str, np.ndarray]:
    """Decode a MindSpore ``.ckpt``: top-level repeated field 1 = one entry per parameter;
    entry field 1 = name, field 2 = tensor {1: dims (repeated varint), 2: type string,
    3: raw little-endian bytes}."""
    with open(path, "rb") as f:
        buf = f.read()
    out: Dict[str, np.ndarray] = {}
    for field, wt, entry in _fields(buf):
        if field != 1 or wt != 2:
            continue
        name = None
        tensor = None
        for f2, wt2, v in _fields(entry):
            if f2 == 1 and wt2 == 2:
                name = v.decode("utf-8")
            elif f2 == 2 and wt2 == 2:
                tensor = v
        if name is None or tensor is None:
            continue
        dims = []
        dtype = None
        raw = b""
        for f3, wt3, v in _fields(tensor):
            if f3 == 1 and wt3 == 0:
                dims.append(v)
            elif f3 == 1 and wt3 == 2:  # packed dims
                p = 0
                while p < len(v):
                    d, p = _varint(v, p)
                    dims.append(d)
            elif f3 == 2 and wt3 == 2:
                dtype = v.decode("ascii")
            elif f3 == 3 and wt3 == 2:
                raw = v
        if dtype not in _MS_DTYPES:
            raise ValueError(f"{path}: parameter {name!r} has unsupported dtype {dtype!r}")
        arr = np.frombuffer(raw, dtype=np.dtype(_MS_DTYPES[dtype]).newbyteorder("<")).copy()
        count = int(np.prod(dims)) if dims else 1
        if count == 0 and arr.size == 1:  # MindSpore writes a 0-d parameter with dims=[0]
            dims = []
            count = 1
        if arr.size != count:
            raise ValueError(f"{path}: parameter {name!r}: {arr.size} values for dims {dims}")
        out[name] = arr.reshape(dims) if dims else arr.reshape(())
    if not out:
        raise ValueError(f"{path}: no parameters found (not a MindSpore checkpoint?)")
    return out


# ------------------------------------------------------------------ name / shape mapping


def load_raw(path: str) -> Dict[str, np.ndarray]:
    """Read ``.npz`` or MindSpore ``.ckpt`` into a ``{name: ndarray}`` dict, names untouched."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if ext == ".ckpt":
        return read_mindspore_ckpt(path)
    raise ValueError(f"unsupported checkpoint extension {ext!r} (expected .npz or .ckpt)")


def ms_to_pt_arrays(raw: Dict[str, np.ndarray], net_size, num_qubits, if_trainable_freq=True,
                    model_type="QuanONet") -> Dict[str, np.ndarray]:
    """MindSpore-named arrays → PyTorch state_dict-named float32 arrays.

    Mirrors ``utils/weight_transfer.py:46-98`` (same errors for missing keys / wrong sizes).
    Arrays already carrying PyTorch names (``quantum_layer.ansatz_weights`` …) pass through.
    """
    if "quantum_layer.ansatz_weights" in raw:
        return {k: np.asarray(v, dtype=np.float32) for k, v in raw.items()}
    if model_type == "QuanONet":
        b_d, b_l, t_d, t_l = net_size
        n_sub = b_d * b_l + t_d * t_l
    else:
        n_sub = net_size[0] * net_size[1]
    sd: Dict[str, np.ndarray] = {}
    if "bias" in raw:
        sd["bias"] = np.asarray(raw["bias"], dtype=np.float32).reshape(1)
    elif model_type == "QuanONet":
        raise KeyError(f"Expected key 'bias' not found. Available: {sorted(raw)}")
    if if_trainable_freq:
        wanted = [k for k in _MS_FREQ_KEYS
                  if (k.startswith(("branch_", "trunk_")) == (model_type == "QuanONet"))]
        for ms_key in wanted:
            if ms_key not in raw:
                raise KeyError(f"Expected key '{ms_key}' not found. Available: {sorted(raw)}")
            sd[_MS_FREQ_KEYS[ms_key]] = np.asarray(raw[ms_key], dtype=np.float32)
    circ = next((k for k in _MS_CIRCUIT_KEYS if k in raw), None)
    if circ is None:
        raise KeyError(f"no circuit weight ({' / '.join(_MS_CIRCUIT_KEYS)}) in checkpoint; "
                       f"available: {sorted(raw)}")
    flat = np.asarray(raw[circ], dtype=np.float32).reshape(-1)
    expected = n_sub * 3 * num_qubits
    if flat.size != expected:
        raise ValueError(f"{circ} has {flat.size} elements but expected {expected} "
                         f"({n_sub}×3×{num_qubits}). Check net_size and num_qubits.")
    sd["quantum_layer.ansatz_weights"] = flat.reshape(n_sub, 3, num_qubits)
    return sd


def ms_npz_to_pt_state_dict(path, net_size=(40, 2, 20, 2), num_qubits=5, if_trainable_freq=True):
    """Same name, arguments and result as ``utils/weight_transfer.py:46`` but also accepts the
    shipped MindSpore ``.ckpt`` files."""
    import torch

    arrays = ms_to_pt_arrays(load_raw(path), net_size, num_qubits, if_trainable_freq)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in arrays.items()}


# One independent pattern per field, the way the reference reads its own directory names (infer.py:60-86);
# the names are written by utils/logger.py:55-118:
#   <Op>_<Model>_Net<a-b[-c-d]>_Q<n>_<TF|FF>_S<scale>[_Pauli<P>][_Diag<v-v-..>|_Ham<lb-ub>][_<TQ|Qiskit|PL>]_<N>x<P>_Seed<k>
_MODEL_RE = re.compile(r"(?:^|_)(QuanONet|HEAQNN)(?:_|$)")
_OP_RE = re.compile(r"^([A-Za-z0-9]+?)_(?:QuanONet|HEAQNN)(?:_|$)")
_NET_RE = re.compile(r"_Net(\d+(?:-\d+)*)(?:_|$)")
_Q_RE = re.compile(r"_Q(\d+)(?:_|$)")
_TF_RE = re.compile(r"_(TF|FF)(?:_|$)")
_S_RE = re.compile(r"_S(\d[\d.]*(?:[eE][+-]?\d+)?)(?:_|$)")
_PAULI_RE = re.compile(r"_Pauli([XYZ])(?:_|$)")
_NUM = r"-?\d[\d.]*(?:[eE][+-]?\d+)?"
_DIAG_RE = re.compile(r"_Diag(" + _NUM + r"(?:-" + _NUM + r")*)(?:_|$)")
_HAM_RE = re.compile(r"_Ham(" + _NUM + r")-(" + _NUM + r")(?:_|$)")
_QB_RE = re.compile(r"_(TQ|Qiskit|PL)(?:_|$)")
_DATA_RE = re.compile(r"_(\d+)x(\d+)_Seed(\d+)(?:_|$)")
_QB_MAP = {"TQ": "torchquantum", "Qiskit": "qiskit", "PL": "pennylane"}


def _split_signed(text: str):
    """'-5--2.5-2.5-5' -> [-5.0, -2.5, 2.5, 5.0]: values joined by '-' (utils/logger.py:95), negatives included.
    A '-' that follows a digit is the separator; any other '-' is a sign."""
    return [float(v) for v in re.findall(_NUM, re.sub(r"(?<=[\d.])-", " ", text))]


def parse_experiment_dir(path: str) -> dict:
    """Recover hyper-parameters from a reference experiment directory name such as
    ``Advection_QuanONet_Net40-2-20-2_Q5_TF_S0.1_1000x100_Seed0`` or
    ``RDiffusion_HEAQNN_Net64-2_Q5_FF_S0.01_PauliX_Ham-1-1_TQ_1000x100_Seed3`` (naming scheme:
    ``utils/logger.py:55-118``; the reference's own parser is ``infer.py:60-86``).  Fields the name does not
    carry are absent from the result (``infer._resolve_config`` fills the reference's defaults, ``infer.py:47-57``);
    ``if_trainable_freq`` is True for ``_TF`` and False for ``_FF``."""
    name = None
    for part in reversed(os.path.normpath(path).split(os.sep)):
        if _MODEL_RE.search(part):
            name = part
            break
    if name is None:
        raise ValueError(f"cannot parse experiment hyper-parameters from {path!r}")
    cfg = {"model_type": _MODEL_RE.search(name).group(1)}
    m = _OP_RE.search(name)
    if m:
        cfg["operator"] = m.group(1)
    m = _NET_RE.search(name)
    if m:
        cfg["net_size"] = tuple(int(v) for v in m.group(1).split("-"))
    m = _Q_RE.search(name)
    if m:
        cfg["num_qubits"] = int(m.group(1))
    m = _TF_RE.search(name)
    if m:
        cfg["if_trainable_freq"] = m.group(1) == "TF"
    m = _S_RE.search(name)
    if m:
        cfg["scale_coeff"] = float(m.group(1))
    m = _PAULI_RE.search(name)
    if m:
        cfg["ham_pauli"] = m.group(1)
    m = _DIAG_RE.search(name)
    if m:
        cfg["ham_diag"] = _split_signed(m.group(1))
    m = _HAM_RE.search(name)
    if m:
        cfg["ham_bound"] = (float(m.group(1)), float(m.group(2)))
    m = _QB_RE.search(name)
    if m:
        cfg["quantum_backend"] = _QB_MAP[m.group(1)]
    m = _DATA_RE.search(name)
    if m:
        cfg["num_train"], cfg["num_points"], cfg["seed"] = int(m.group(1)), int(m.group(2)), int(m.group(3))
    return cfg
