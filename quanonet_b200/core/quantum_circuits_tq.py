"""Drop-in replacement for the reference's ``core/quantum_circuits_tq.py``.

Same public surface (``build_quanonet_tq``, ``build_heaqnn_tq``, ``_make_block_configs``,
``_ham_params``, ``_TQHEACircuit``; reference ``core/quantum_circuits_tq.py:39-202``), same
``forward(x: (B,E)) -> (B,1)`` contract, same parameter name/shape/initialisation
(``ansatz_weights (S,3,n) ~ U(-pi,pi)``, buffer ``ham_diag`` when a diagonal is given), so
``core/models_pt.py``, ``solvers/solver_pt.py``, ``infer.py``, ``compare_backends.py`` and
``utils/weight_transfer.py`` use it unchanged.  Instead of TorchQuantum's gate-by-gate PyTorch
ops with autograd, ``forward`` calls the ``quanonet::hea_expval`` custom op: one fused CUDA kernel
for the forward pass and one for forward+adjoint-backward (csrc/).  No TorchQuantum import, no
backend dispatch, no CPU fallback — a CPU tensor raises.

Extensions beyond the reference (all keyword-only, defaults reproduce the reference):
``ham_pauli`` ('Z' | 'X' | 'Y') — the reference's PyTorch path silently ignores ``--ham_pauli``
(``solvers/solver_pt.py:88-93``); ``diag_order`` ('msb0' = TorchQuantum's flattening, the default
here, or 'lsb0' = MindQuantum's); float64 parameters/inputs select the complex128 kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..ops import hea_expval_autograd as hea_expval

_PAULI_KIND = {"Z": _lib.QON_HAM_DIAG, "X": _lib.QON_HAM_PAULI_X, "Y": _lib.QON_HAM_PAULI_Y}
_DIAG_ORDER = {"lsb0": _lib.QON_DIAG_LSB0, "msb0": _lib.QON_DIAG_MSB0}


def _canonical_plan(n_wires: int, block_configs: Sequence[Tuple[int, int]], n_cols: int):
    """Map the reference's ``block_configs`` / ``x`` columns onto the C-ABI's canonical form.

    Reference semantics (``core/quantum_circuits_tq.py:79-102``): block k applies ``n_encode`` RX
    gates on wires ``j % n`` consuming successive columns of ``x`` (gates whose column is past the
    end are skipped), then ``linear_depth`` sublayers.  Canonical form: every block has exactly one
    angle per qubit and depth >= 1.  Because RX rotations on one wire add, a canonical angle is the
    SUM of the source columns that land on that wire before the next sublayer; missing ones are 0
    (RX(0) is exactly the identity).  A block with ``linear_depth == 0`` therefore merges into the
    next block.  Returns ``(depths, groups, identity)`` where ``groups[c]`` lists the source
    columns feeding canonical column c and ``identity`` says x can be passed through untouched.
    """
    depths: List[int] = []
    groups: List[List[int]] = []
    pending: List[List[int]] = [[] for _ in range(n_wires)]
    col = 0
    for n_enc, depth in block_configs:
        if n_enc < 0 or depth < 0:
            raise ValueError(f"invalid block config {(n_enc, depth)}")
        for j in range(n_enc):
            if col < n_cols:
                pending[j % n_wires].append(col)
            col += 1
        if depth > 0:
            depths.append(int(depth))
            groups.extend(pending)
            pending = [[] for _ in range(n_wires)]
    if any(pending):
        raise NotImplementedError(
            "block_configs end with an encoding layer that no ansatz sublayer follows "
            "(trailing linear_depth == 0); not supported by the B200 kernels")
    if not depths:
        raise ValueError("circuit has no ansatz sublayer")
    identity = (len(groups) == n_cols and all(g == [i] for i, g in enumerate(groups)))
    return depths, groups, identity


class _TQHEACircuit(nn.Module):
    """HEA circuit module with the constructor and attributes of the reference class
    (``core/quantum_circuits_tq.py:39-63``), evaluated by the B200 kernels."""

    def __init__(self, n_wires, block_configs, ham_offset=0.0, ham_coeff_per_qubit=0.0, ham_diag=None,
                 *, ham_pauli: str = "Z", diag_order: str = "msb0"):
        super().__init__()
        if ham_pauli not in _PAULI_KIND:
            raise ValueError(f"ham_pauli must be one of X, Y, Z (got {ham_pauli!r})")
        if diag_order not in _DIAG_ORDER:
            raise ValueError(f"diag_order must be 'msb0' or 'lsb0' (got {diag_order!r})")
        self.n_wires = int(n_wires)
        self.block_configs = [(int(e), int(d)) for e, d in block_configs]
        self.ham_pauli = ham_pauli
        self.diag_order = diag_order
        total_ansatz_blocks = sum(d for _, d in self.block_configs)
        self.ansatz_weights = nn.Parameter(torch.empty(total_ansatz_blocks, 3, self.n_wires))
        nn.init.uniform_(self.ansatz_weights, -np.pi, np.pi)
        if ham_diag is not None:
            if ham_pauli != "Z":
                raise ValueError("ham_diag is a computational-basis diagonal; it excludes ham_pauli X/Y")
            self.register_buffer("ham_diag", torch.tensor(np.asarray(ham_diag), dtype=torch.float32))
            if self.ham_diag.numel() != 2 ** self.n_wires:
                raise ValueError(f"ham_diag must have 2**n_wires = {2 ** self.n_wires} entries")
            self.use_full_ham = True
        else:
            self.ham_offset = float(ham_offset)
            self.ham_coeff = float(ham_coeff_per_qubit)
            self.use_full_ham = False
        self._plans = {}

    def _plan(self, n_cols: int):
        plan = self._plans.get(n_cols)
        if plan is None:
            depths, groups, identity = _canonical_plan(self.n_wires, self.block_configs, n_cols)
            gather = None
            if not identity:
                width = max((len(g) for g in groups), default=0)
                # index n_cols points at an appended zero column
                idx = torch.full((len(groups), max(width, 1)), n_cols, dtype=torch.long)
                for c, g in enumerate(groups):
                    for t, src in enumerate(g):
                        idx[c, t] = src
                gather = idx
            plan = (depths, gather)
            self._plans[n_cols] = plan
        return plan

    def canonical_inputs(self, x: torch.Tensor):
        """(x_canonical (B, n*K), depth_per_block) for an encoding-angle matrix ``x (B, E)``."""
        depths, gather = self._plan(int(x.shape[1]))
        if gather is not None:
            xz = torch.cat([x, x.new_zeros(x.shape[0], 1)], dim=1)
            x = xz[:, gather.to(x.device)].sum(dim=2)
        return x, depths

    def forward(self, x):
        """x: (batch, total_encode_params) -> (batch, 1) Hamiltonian expectation value."""
        if x.dim() != 2:
            raise ValueError(f"x must be 2-D (batch, encode_params), got shape {tuple(x.shape)}")
        w = self.ansatz_weights
        if x.dtype != w.dtype:
            x = x.to(w.dtype)
        xc, depths = self.canonical_inputs(x)
        if self.use_full_ham:
            return hea_expval(xc, w, self.n_wires, depths, self.ham_diag.to(device=x.device, dtype=w.dtype),
                              _DIAG_ORDER[self.diag_order], 0.0, 0.0, _lib.QON_HAM_DIAG)
        return hea_expval(xc, w, self.n_wires, depths, None, _lib.QON_DIAG_LSB0,
                          self.ham_offset, self.ham_coeff, _PAULI_KIND[self.ham_pauli])

    def extra_repr(self):
        ham = "diag" if self.use_full_ham else f"{self.ham_offset:g}+{self.ham_coeff:g}*sum {self.ham_pauli}_i"
        return f"n_wires={self.n_wires}, blocks={len(self.block_configs)}, H={ham}, backend=quanonet_b200"


HEACircuitB200 = _TQHEACircuit


def _make_block_configs(num_qubits, trunk_depth, trunk_linear_depth, branch_depth, branch_linear_depth):
    """Trunk blocks first, then branch blocks (reference ``core/quantum_circuits_tq.py:130-138``)."""
    return ([(num_qubits, trunk_linear_depth)] * trunk_depth
            + [(num_qubits, branch_linear_depth)] * branch_depth)


def _ham_params(num_qubits, lower_bound=-5.0, upper_bound=5.0):
    """(offset, coeff_per_qubit) of ``offset + coeff * sum_i P_i`` spanning [lower, upper]
    (reference ``core/quantum_circuits_tq.py:141-146``)."""
    width = upper_bound - lower_bound
    return lower_bound + width / 2.0, width / 2.0 / num_qubits


def _build(num_qubits, block_configs, ham_bound, ham_diag, ham_pauli, diag_order):
    if ham_diag is not None:
        return _TQHEACircuit(num_qubits, block_configs, ham_diag=ham_diag, diag_order=diag_order)
    offset, coeff = _ham_params(num_qubits, ham_bound[0], ham_bound[1])
    return _TQHEACircuit(num_qubits, block_configs, ham_offset=offset, ham_coeff_per_qubit=coeff,
                         ham_pauli=ham_pauli)


def build_quanonet_tq(num_qubits, branch_input_size, trunk_input_size, net_size,
                      ham_bound=(-5.0, 5.0), ham_diag=None, *, ham_pauli="Z", diag_order="msb0"):
    """QuanONet circuit; ``net_size = (branch_depth, branch_linear_depth, trunk_depth,
    trunk_linear_depth)`` (reference ``core/quantum_circuits_tq.py:149-175``)."""
    branch_depth, branch_linear_depth, trunk_depth, trunk_linear_depth = net_size
    cfg = _make_block_configs(num_qubits, trunk_depth, trunk_linear_depth, branch_depth, branch_linear_depth)
    return _build(num_qubits, cfg, ham_bound, ham_diag, ham_pauli, diag_order)


def build_heaqnn_tq(num_qubits, input_size, net_size, ham_bound=(-5.0, 5.0), ham_diag=None,
                    *, ham_pauli="Z", diag_order="msb0"):
    """HEAQNN circuit; ``net_size = (depth, linear_depth, _, _)``
    (reference ``core/quantum_circuits_tq.py:178-202``)."""
    cfg = [(num_qubits, net_size[1])] * net_size[0]
    return _build(num_qubits, cfg, ham_bound, ham_diag, ham_pauli, diag_order)
