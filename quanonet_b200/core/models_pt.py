"""``QuanONetPT`` / ``HEAQNNPT`` with the reference's constructor and forward signatures
(reference ``core/models_pt.py:103-213``), wired to the B200 quantum layer.

The reference's own ``core/models_pt.py`` works unchanged on top of
``quanonet_b200.core.quantum_circuits_tq``; this module exists so the package
is usable — and testable on the GPU box — without the reference checkout.  Parameter names match
the reference state_dict: ``branch_freq.weights/bias``, ``trunk_freq.weights/bias`` (``freq.*`` for
HEAQNN), ``quantum_layer.ansatz_weights``, ``bias``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _tile_to(x: torch.Tensor, out_features: int) -> torch.Tensor:
    """Repeat the columns of ``x (B, in)`` cyclically up to ``out_features`` columns
    (column j of the result is input column ``j % in``; extra inputs are dropped)."""
    reps = math.ceil(out_features / x.shape[1])
    return x.repeat(1, reps)[:, :out_features]


class _TiledElementWise(nn.Module):
    """Trainable frequency layer ``enc[b,j] = x[b, j % in] * weights[j] + bias[j]``
    (reference ``core/models_pt.py:14-41``; MindSpore twin ``core/layers.py:14-30,96-107``)."""

    def __init__(self, in_features, out_features, init_scale=0.1):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.repeats = math.ceil(out_features / in_features)
        self.weights = nn.Parameter(torch.full((out_features,), float(init_scale)))
        self.bias = nn.Parameter(torch.zeros(out_features))

    def forward(self, x):
        return _tile_to(x, self.out_features) * self.weights + self.bias


class _ScaleRepeat(nn.Module):
    """Fixed frequency layer ``enc[b,j] = scale * x[b, j % in]`` (reference ``core/models_pt.py:44-68``)."""

    def __init__(self, in_features, out_features, scale=0.01):
        super().__init__()
        self.scale = scale
        self.in_features = in_features
        self.out_features = out_features
        self.repeats = math.ceil(out_features / in_features)

    def forward(self, x):
        return _tile_to(x * self.scale, self.out_features)


def _build_quantum_layer(quantum_backend, num_qubits, total_input_size, net_size, ham_bound, ham_diag,
                         branch_input_size=None, trunk_input_size=None, **extra):
    """Only the 'torchquantum' route exists here, and it resolves to the B200 module
    (reference dispatch: ``core/models_pt.py:71-100``)."""
    if quantum_backend not in ("torchquantum", "b200", "quanonet_b200"):
        raise ValueError(f"quanonet_b200 implements the PyTorch quantum path only; "
                         f"quantum_backend={quantum_backend!r} is not available here")
    from .quantum_circuits_tq import build_heaqnn_tq, build_quanonet_tq
    if branch_input_size is not None:
        return build_quanonet_tq(num_qubits, branch_input_size, trunk_input_size, net_size,
                                 ham_bound=ham_bound, ham_diag=ham_diag, **extra)
    return build_heaqnn_tq(num_qubits, total_input_size, net_size, ham_bound=ham_bound, ham_diag=ham_diag, **extra)


def _can_fuse_inference(qlayer, u):
    """Inference (no autograd) on CUDA in the register tier: evaluate the frequency layers inside the
    kernel (``quanonet::encoded_expval``) instead of materialising the encoding matrix."""
    if torch.is_grad_enabled() or not u.is_cuda:
        return False
    from ..ops import encoded_supported
    w = qlayer.ansatz_weights
    if not all(e == qlayer.n_wires and d >= 1 for e, d in qlayer.block_configs):
        return False
    return encoded_supported(qlayer.n_wires, w.dtype, int(u.shape[0]), need_grad=False,
                             depths=[d for _, d in qlayer.block_configs])


def _fused_inference(qlayer, layers, u0, u1, enc0):
    from .. import _lib
    from ..ops import encoded_expval
    from .quantum_circuits_tq import _DIAG_ORDER, _PAULI_KIND
    w = qlayer.ansatz_weights
    if isinstance(layers[0], _TiledElementWise):
        fw = torch.cat([l.weights for l in layers])
        fb = torch.cat([l.bias for l in layers])
    else:
        fw = torch.cat([torch.full((l.out_features,), float(l.scale), dtype=w.dtype, device=w.device) for l in layers])
        fb = None
    depths = [d for _, d in qlayer.block_configs]
    u1 = u1.to(w.dtype)
    u0 = None if u0 is None else u0.to(w.dtype)
    if qlayer.use_full_ham:
        ham = (qlayer.ham_diag.to(device=w.device, dtype=w.dtype), _DIAG_ORDER[qlayer.diag_order], 0.0, 0.0,
               _lib.QON_HAM_DIAG)
    else:
        ham = (None, _lib.QON_DIAG_LSB0, qlayer.ham_offset, qlayer.ham_coeff, _PAULI_KIND[qlayer.ham_pauli])
    return encoded_expval(u0, u1, fw, fb, enc0 // qlayer.n_wires, w, qlayer.n_wires, depths, *ham)


class QuanONetPT(nn.Module):
    """``forward(branch_input (B, b_in), trunk_input (B, t_in)) -> (B,1)``:
    frequency layers -> ``cat([trunk_enc, branch_enc])`` -> quantum layer -> ``+ bias``
    (reference ``core/models_pt.py:124-166``)."""

    def __init__(self, num_qubits, branch_input_size, trunk_input_size, net_size, scale_coeff=1.0,
                 if_trainable_freq=False, quantum_backend="torchquantum", ham_bound=(-5.0, 5.0), ham_diag=None,
                 **extra):
        super().__init__()
        branch_depth, _, trunk_depth, _ = net_size
        self.if_trainable_freq = if_trainable_freq
        self.branch_enc_size = branch_depth * num_qubits
        self.trunk_enc_size = trunk_depth * num_qubits
        layer = _TiledElementWise if if_trainable_freq else _ScaleRepeat
        self.branch_freq = layer(branch_input_size, self.branch_enc_size, scale_coeff)
        self.trunk_freq = layer(trunk_input_size, self.trunk_enc_size, scale_coeff)
        self.quantum_layer = _build_quantum_layer(
            quantum_backend, num_qubits, self.trunk_enc_size + self.branch_enc_size, net_size, ham_bound, ham_diag,
            branch_input_size=branch_input_size, trunk_input_size=trunk_input_size, **extra)
        self.bias = nn.Parameter(torch.zeros(1))

    def forward(self, branch_input, trunk_input):
        if _can_fuse_inference(self.quantum_layer, branch_input):
            return _fused_inference(self.quantum_layer, (self.trunk_freq, self.branch_freq),
                                    trunk_input, branch_input, self.trunk_enc_size) + self.bias
        x = torch.cat([self.trunk_freq(trunk_input), self.branch_freq(branch_input)], dim=1)
        return self.quantum_layer(x) + self.bias


class HEAQNNPT(nn.Module):
    """``forward(x (B, in)) -> (B,1)``; no bias parameter (reference ``core/models_pt.py:169-213``)."""

    def __init__(self, num_qubits, input_size, net_size, scale_coeff=1.0, if_trainable_freq=False,
                 quantum_backend="torchquantum", ham_bound=(-5.0, 5.0), ham_diag=None, **extra):
        super().__init__()
        enc_size = net_size[0] * num_qubits
        self.if_trainable_freq = if_trainable_freq
        layer = _TiledElementWise if if_trainable_freq else _ScaleRepeat
        self.freq = layer(input_size, enc_size, scale_coeff)
        self.quantum_layer = _build_quantum_layer(quantum_backend, num_qubits, enc_size, net_size, ham_bound,
                                                  ham_diag, **extra)

    def forward(self, x):
        if _can_fuse_inference(self.quantum_layer, x):
            return _fused_inference(self.quantum_layer, (self.freq,), None, x, 0)
        return self.quantum_layer(self.freq(x))
