"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_circuit_cases():
    z = np.load(os.path.join(GOLDEN, "circuit_cases.npz"))
    with open(os.path.join(GOLDEN, "circuit_cases.json")) as f:
        meta = json.load(f)
    return z, meta


def ham_kwargs_for_op(meta_ham, n):
    """Golden-case Hamiltonian description -> keyword arguments of quanonet::hea_expval."""
    from quanonet_b200 import _lib
    if meta_ham["kind"] == "diag":
        order = _lib.QON_DIAG_MSB0 if meta_ham["order"] == "msb0" else _lib.QON_DIAG_LSB0
        return dict(ham_diag=np.asarray(meta_ham["diag"]), diag_order=order, ham_offset=0.0, ham_coeff=0.0,
                    ham_kind=_lib.QON_HAM_DIAG)
    lb, ub = meta_ham["bound"]
    width = ub - lb
    kind = {"Z": _lib.QON_HAM_DIAG, "X": _lib.QON_HAM_PAULI_X, "Y": _lib.QON_HAM_PAULI_Y}[meta_ham["pauli"]]
    return dict(ham_diag=None, diag_order=_lib.QON_DIAG_LSB0, ham_offset=lb + width / 2.0,
                ham_coeff=width / 2.0 / n, ham_kind=kind)


def oracle_ham(meta_ham, n):
    from oracle import hea_oracle as orc
    if meta_ham["kind"] == "diag":
        return orc.ham_from_diag(meta_ham["diag"], n, meta_ham["order"])
    return orc.ham_from_bound(n, meta_ham["bound"][0], meta_ham["bound"][1], pauli=meta_ham["pauli"])
