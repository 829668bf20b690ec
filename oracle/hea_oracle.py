"""TEST INFRASTRUCTURE ONLY — fp64 statevector + adjoint-gradient oracle of the HEA hot path.

A NumPy restatement of what the reference computes on this path, written from its
*specification*, not from its code.  Each function cites the reference file:line it follows
(paths relative to the reference checkout).

Where the arithmetic really lives: third-party, un-vendored packages that are absent from the
reference tree and from this image —

* ``torchquantum==0.1.8`` (reference ``requirements.txt:28``) — complex64, autograd backprop;
  call sites ``core/quantum_circuits_tq.py:74,84-101,109``;
* ``mindquantum==0.11.0`` (reference ``requirements.txt:21``) — complex128 ``mqvector`` simulator
  with a built-in adjoint gradient; call sites ``core/quantum_circuits_ms.py:229-233``.

Their published gate definitions are restated here: ``RX(t)=exp(-i t X/2)``, ``RY(t)=exp(-i t Y/2)``,
``RZ(t)=exp(-i t Z/2)``, CNOT with an explicit (control, target).

Parity status: PINNED by the reference's published outputs — the twelve MSE/MAE figures printed
in ``visualization.ipynb`` for the three shipped Q5 checkpoints and the closed-form Antideriv
answers of ``ibm_inference.py:177-189`` (``tests/test_oracle_golden.py``,
``tests/golden/make_golden.py``).  No *runnable* reference test exists for this path because
every simulator backend is absent; see DESIGN.md.

Index convention of this oracle (and of the CUDA kernels): amplitude index ``k`` has qubit ``q``
at bit ``q`` (qubit 0 = least-significant bit) — MindQuantum's convention
(``core/quantum_circuits_ms.py:56-60``).  TorchQuantum flattens with wire 0 as the MOST
significant bit; ``diag_msb0_to_lsb0`` converts a diagonal given in that order.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# circuit structure
# --------------------------------------------------------------------------------------


def make_block_configs(num_qubits, trunk_depth, trunk_linear_depth, branch_depth, branch_linear_depth):
    """Trunk blocks first, then branch blocks; every block encodes ``num_qubits`` angles.

    Follows ``core/quantum_circuits_tq.py:130-138`` and ``core/quantum_circuits_ms.py:175-196``.
    """
    return [(num_qubits, trunk_linear_depth)] * trunk_depth + [(num_qubits, branch_linear_depth)] * branch_depth


def heaqnn_block_configs(num_qubits, depth, linear_depth):
    """``core/quantum_circuits_tq.py:194-196`` / ``core/quantum_circuits_ms.py:207-221``."""
    return [(num_qubits, linear_depth)] * depth


def gate_list(n: int, blocks: Sequence[Tuple[int, int]], n_cols: int):
    """Flatten the circuit into a gate list in application order.

    Entries: ``('rx', qubit, col)``, ``('ry'|'rz', qubit, (s, g, i))``, ``('cnot', control, target)``.
    Order follows ``core/quantum_circuits_tq.py:79-102``: per block, ``n_encode`` RX gates on wire
    ``j % n`` (a gate whose column is past the end of ``x`` is skipped, ``:83``), then
    ``linear_depth`` sublayers of per-qubit RY·RZ·RY followed by the CNOT ring
    ``control=(i+1)%n, target=i`` for ``i=0..n-1`` (``:98-101``).  For ``n == 1`` the ring is
    skipped, as MindQuantum does (``core/quantum_circuits_ms.py:140``).
    """
    gates = []
    col = 0
    s = 0
    for n_enc, depth in blocks:
        for j in range(n_enc):
            if col < n_cols:
                gates.append(("rx", j % n, col))
            col += 1
        for _ in range(depth):
            for i in range(n):
                gates.append(("ry", i, (s, 0, i)))
                gates.append(("rz", i, (s, 1, i)))
                gates.append(("ry", i, (s, 2, i)))
            if n > 1:
                for i in range(n):
                    gates.append(("cnot", (i + 1) % n, i))
            s += 1
    return gates


def num_sublayers(blocks):
    return int(sum(d for _, d in blocks))


def num_encode_cols(blocks):
    return int(sum(e for e, _ in blocks))


# --------------------------------------------------------------------------------------
# Hamiltonians
# --------------------------------------------------------------------------------------


@dataclass
class Ham:
    """``kind='pauli'``: ``offset*I + coeff * sum_i P_i`` with ``P`` in X/Y/Z
    (``core/quantum_circuits_ms.py:28-39``); ``kind='diag'``: arbitrary real diagonal in LSB0 order
    (``core/quantum_circuits_tq.py:112-114``, ``core/quantum_circuits_ms.py:41-63``)."""

    kind: str = "pauli"
    pauli: str = "Z"
    offset: float = 0.0
    coeff: float = 1.0
    diag: Optional[np.ndarray] = None


def ham_params(num_qubits, lower_bound=-5.0, upper_bound=5.0):
    """``core/quantum_circuits_tq.py:141-146``: offset = midpoint, coeff = half-width / n."""
    width = upper_bound - lower_bound
    return lower_bound + width / 2.0, width / 2.0 / num_qubits


def ham_from_bound(num_qubits, lower_bound=-5.0, upper_bound=5.0, pauli="Z") -> Ham:
    off, c = ham_params(num_qubits, lower_bound, upper_bound)
    return Ham("pauli", pauli, off, c)


def diag_msb0_to_lsb0(diag, n):
    """Re-index a diagonal given with wire 0 as the most-significant bit (TorchQuantum's
    ``get_states_1d`` order, what ``core/quantum_circuits_tq.py:112-114`` multiplies against)
    into this oracle's qubit-0-is-LSB order."""
    diag = np.asarray(diag, dtype=np.float64)
    k = np.arange(1 << n)
    rev = np.zeros_like(k)
    for q in range(n):
        rev |= ((k >> q) & 1) << (n - 1 - q)
    return diag[rev]


def ham_from_diag(diag, n, order="lsb0") -> Ham:
    d = np.asarray(diag, dtype=np.float64)
    if d.shape != (1 << n,):
        raise ValueError("diagonal must have 2**n entries")
    if order == "msb0":
        d = diag_msb0_to_lsb0(d, n)
    elif order != "lsb0":
        raise ValueError(order)
    return Ham("diag", diag=d)


def zero_state_ham(n, lower_bound=0.0, upper_bound=1.0) -> Ham:
    """``lb*I + (ub-lb)|0..0><0..0|`` (``core/quantum_circuits_ms.py:17-25``)."""
    d = np.full(1 << n, float(lower_bound))
    d[0] = float(upper_bound)
    return Ham("diag", diag=d)


def z_sum_diag(n):
    k = np.arange(1 << n)
    return sum(1.0 - 2.0 * ((k >> q) & 1) for q in range(n))


def apply_ham(psi: np.ndarray, n: int, ham: Ham) -> np.ndarray:
    """Return ``H psi`` for a batch of states ``(B, 2**n)``."""
    if ham.kind == "diag":
        return psi * ham.diag.astype(psi.real.dtype)[None, :]
    if ham.pauli == "Z":
        d = (ham.offset + ham.coeff * z_sum_diag(n)).astype(psi.real.dtype)
        return psi * d[None, :]
    out = ham.offset * psi
    k = np.arange(1 << n)
    for q in range(n):
        flipped = psi[:, k ^ (1 << q)]
        if ham.pauli == "X":
            out = out + ham.coeff * flipped
        elif ham.pauli == "Y":
            # (Y psi)_k = -i psi_{k^b} when bit q of k is 0, +i psi_{k^b} when it is 1
            sign = (2.0 * ((k >> q) & 1) - 1.0).astype(psi.real.dtype)
            out = out + ham.coeff * (1j * sign)[None, :] * flipped
        else:
            raise ValueError(ham.pauli)
    return out


# --------------------------------------------------------------------------------------
# gates
# --------------------------------------------------------------------------------------


def _rot_matrix(kind, theta, cdtype):
    """(…,2,2) matrices. RX=[[c,-is],[-is,c]], RY=[[c,-s],[s,c]], RZ=diag(e^{-it/2},e^{+it/2})."""
    theta = np.asarray(theta)
    c = np.cos(theta / 2.0)
    s = np.sin(theta / 2.0)
    m = np.zeros(theta.shape + (2, 2), dtype=cdtype)
    if kind == "rx":
        m[..., 0, 0] = c
        m[..., 1, 1] = c
        m[..., 0, 1] = -1j * s
        m[..., 1, 0] = -1j * s
    elif kind == "ry":
        m[..., 0, 0] = c
        m[..., 1, 1] = c
        m[..., 0, 1] = -s
        m[..., 1, 0] = s
    elif kind == "rz":
        m[..., 0, 0] = c - 1j * s
        m[..., 1, 1] = c + 1j * s
    else:
        raise ValueError(kind)
    return m


def _apply_1q(psi, q, m):
    """Apply 2x2 matrix/matrices ``m`` (shape (2,2) or (B,2,2)) on qubit ``q`` of ``psi (B, N)``."""
    B, N = psi.shape
    v = psi.reshape(B, N >> (q + 1), 2, 1 << q)
    if m.ndim == 2:
        out = np.einsum("ij,bhjl->bhil", m, v)
    else:
        out = np.einsum("bij,bhjl->bhil", m, v)
    return out.reshape(B, N)


def _apply_pauli(psi, q, kind):
    if kind == "rx":
        p = np.array([[0, 1], [1, 0]], dtype=psi.dtype)
    elif kind == "ry":
        p = np.array([[0, -1j], [1j, 0]], dtype=psi.dtype)
    else:
        p = np.array([[1, 0], [0, -1]], dtype=psi.dtype)
    return _apply_1q(psi, q, p)


def _cnot_perm(n, control, target):
    k = np.arange(1 << n)
    return k ^ (((k >> control) & 1) << target)


# --------------------------------------------------------------------------------------
# forward / adjoint backward
# --------------------------------------------------------------------------------------


def hea_state(x, w, n, blocks, cdtype=np.complex128):
    """Final statevector ``(B, 2**n)`` of the HEA circuit."""
    x = np.asarray(x)
    w = np.asarray(w)
    B = x.shape[0]
    rdtype = np.float64 if cdtype == np.complex128 else np.float32
    x = x.astype(rdtype)
    w = w.astype(rdtype)
    psi = np.zeros((B, 1 << n), dtype=cdtype)
    psi[:, 0] = 1.0
    for kind, a, b in gate_list(n, blocks, x.shape[1]):
        if kind == "rx":
            psi = _apply_1q(psi, a, _rot_matrix("rx", x[:, b], cdtype))
        elif kind == "cnot":
            psi = psi[:, _cnot_perm(n, a, b)]
        else:
            psi = _apply_1q(psi, a, _rot_matrix(kind, w[b], cdtype))
    return psi


def hea_forward(x, w, n, blocks, ham: Ham, cdtype=np.complex128):
    """Expectation values ``(B,)``: ``<psi|H|psi>`` (``core/quantum_circuits_tq.py:106-127``)."""
    psi = hea_state(x, w, n, blocks, cdtype)
    return np.real(np.sum(np.conj(psi) * apply_ham(psi, n, ham), axis=1))


def hea_forward_backward(x, w, n, blocks, ham: Ham, grad_out=None, cdtype=np.complex128, return_state=False):
    """Adjoint differentiation (what MindQuantum's ``get_expectation_with_grad`` computes,
    ``core/quantum_circuits_ms.py:229-233``; equals autograd through
    ``core/quantum_circuits_tq.py:65-127`` mathematically).

    Returns ``(E (B,), grad_x (B, n_cols), grad_w (S,3,n))`` where, with ``g = grad_out`` (ones if
    None), ``grad_x[b,c] = g_b dE_b/dx[b,c]`` and ``grad_w = sum_b g_b dE_b/dw``.
    Rule: for a rotation ``exp(-i t P/2)`` with ``psi`` the state just after it and
    ``lam = U_after^dagger H psi_final``: ``dE/dt = Im <lam|P|psi>``.
    """
    x = np.asarray(x)
    w = np.asarray(w)
    B, n_cols = x.shape
    rdtype = np.float64 if cdtype == np.complex128 else np.float32
    x = x.astype(rdtype)
    w = w.astype(rdtype)
    g = np.ones(B, dtype=rdtype) if grad_out is None else np.asarray(grad_out, dtype=rdtype).reshape(B)
    gates = gate_list(n, blocks, n_cols)
    psi = hea_state(x, w, n, blocks, cdtype)
    lam = apply_ham(psi, n, ham)
    E = np.real(np.sum(np.conj(psi) * lam, axis=1))
    grad_x = np.zeros((B, n_cols), dtype=rdtype)
    grad_w = np.zeros(w.shape, dtype=rdtype)
    for kind, a, b in reversed(gates):
        if kind == "cnot":
            perm = _cnot_perm(n, a, b)  # self-inverse
            psi = psi[:, perm]
            lam = lam[:, perm]
            continue
        d = np.imag(np.sum(np.conj(lam) * _apply_pauli(psi, a, kind), axis=1))  # (B,)
        if kind == "rx":
            grad_x[:, b] = g * d
            m = _rot_matrix("rx", -x[:, b], cdtype)
        else:
            grad_w[b] += np.sum(g * d)
            m = _rot_matrix(kind, -w[b], cdtype)
        psi = _apply_1q(psi, a, m)
        lam = _apply_1q(lam, a, m)
    if return_state:  # after the reverse sweep psi must be back at |0..0> (free self-check)
        return E, grad_x, grad_w, psi
    return E, grad_x, grad_w


# --------------------------------------------------------------------------------------
# model wrappers (frequency layer + bias), restated from core/models_pt.py
# --------------------------------------------------------------------------------------


def tiled_elementwise(u, out_features, weights, bias):
    """``enc[b,j] = u[b, j % in] * w[j] + bias[j]`` for ``j < out`` (``core/models_pt.py:38-41``)."""
    u = np.asarray(u, dtype=np.float64)
    reps = int(np.ceil(out_features / u.shape[1]))
    tiled = np.tile(u, (1, reps))[:, :out_features]
    return tiled * np.asarray(weights, dtype=np.float64) + np.asarray(bias, dtype=np.float64)


def scale_repeat(u, out_features, scale):
    """``enc[b,j] = scale * u[b, j % in]`` (``core/models_pt.py:63-68``)."""
    u = np.asarray(u, dtype=np.float64) * scale
    reps = int(np.ceil(out_features / u.shape[1]))
    return np.tile(u, (1, reps))[:, :out_features]


def quanonet_forward(branch, trunk, params, n, net_size, ham: Ham, trainable_freq=True, scale=1.0):
    """``QuanONetPT.forward`` (``core/models_pt.py:153-166``): trunk encoding first, then branch,
    quantum layer, plus bias.  ``params`` keys follow the PyTorch state_dict names."""
    b_d, b_l, t_d, t_l = net_size
    if trainable_freq:
        be = tiled_elementwise(branch, b_d * n, params["branch_freq.weights"], params["branch_freq.bias"])
        te = tiled_elementwise(trunk, t_d * n, params["trunk_freq.weights"], params["trunk_freq.bias"])
    else:
        be = scale_repeat(branch, b_d * n, scale)
        te = scale_repeat(trunk, t_d * n, scale)
    xenc = np.concatenate([te, be], axis=1)
    blocks = make_block_configs(n, t_d, t_l, b_d, b_l)
    e = hea_forward(xenc, params["quantum_layer.ansatz_weights"], n, blocks, ham)
    return e + float(np.asarray(params["bias"]).reshape(-1)[0])
