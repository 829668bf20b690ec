"""TEST INFRASTRUCTURE ONLY — TorchQuantum-faithful complex64 restatement (the timed CPU baseline).

``torchquantum==0.1.8`` (reference ``requirements.txt:28``) is not vendored in the reference and
not installed here, so this module restates what its functional gates do on the reference's call
sequence (``core/quantum_circuits_tq.py:65-127``) with the same dtypes and the same differentiation
method (plain PyTorch autograd through every gate):

* state: complex64 tensor ``(B, 2, …, 2)`` initialised to |0…0>, wire ``w`` on tensor axis ``w+1``
  (so wire 0 is the most-significant bit of the flattened index);
* a one-qubit gate builds a per-sample ``(B,2,2)`` matrix from ``cos/sin(theta/2)``, moves the
  wire's axis last, applies a batched matmul and moves it back;
* ``RX=[[c,-is],[-is,c]]``, ``RY=[[c,-s],[s,c]]``, ``RZ=diag(e^{-it/2}, e^{it/2})``; CNOT's first wire
  is the control;
* measurement exactly as ``_measure`` (``core/quantum_circuits_tq.py:106-127``).

Parity status: pinned against oracle/hea_oracle.py (itself pinned by published reference outputs)
in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch


def _gate_matrix(kind, theta):
    """theta (B,) float32 → (B,2,2) complex64."""
    half = theta * 0.5
    c = torch.cos(half)
    s = torch.sin(half)
    z = torch.zeros_like(c)
    if kind == "rx":
        m = torch.stack([torch.complex(c, z), torch.complex(z, -s),
                         torch.complex(z, -s), torch.complex(c, z)], dim=-1)
    elif kind == "ry":
        m = torch.stack([torch.complex(c, z), torch.complex(-s, z),
                         torch.complex(s, z), torch.complex(c, z)], dim=-1)
    elif kind == "rz":
        m = torch.stack([torch.complex(c, -s), torch.complex(z, z),
                         torch.complex(z, z), torch.complex(c, s)], dim=-1)
    else:
        raise ValueError(kind)
    return m.reshape(-1, 2, 2)


def _apply_1q(state, wire, mat):
    """state (B,2,..,2); mat (B,2,2). new[..., i, ...] = sum_j mat[b,i,j] state[..., j, ...]."""
    n = state.dim() - 1
    axis = wire + 1
    st = state.movedim(axis, -1)                       # (B, ..., 2)
    shp = st.shape
    st = st.reshape(shp[0], -1, 2)                     # (B, M, 2)
    st = torch.bmm(st, mat.transpose(1, 2))            # (B, M, 2)
    return st.reshape(shp).movedim(-1, axis)


def _apply_cnot(state, control, target):
    c_ax, t_ax = control + 1, target + 1
    idx0 = [slice(None)] * state.dim()
    idx1 = [slice(None)] * state.dim()
    idx0[c_ax] = 0
    idx1[c_ax] = 1
    s0 = state[tuple(idx0)]
    s1 = state[tuple(idx1)]
    t_rel = t_ax - 1 if t_ax > c_ax else t_ax
    s1 = s1.flip(t_rel)
    return torch.stack([s0, s1], dim=c_ax)


def tq_forward(x, weights, n_wires, block_configs, ham_offset=0.0, ham_coeff=1.0, ham_diag=None):
    """``_TQHEACircuit.forward`` + ``_measure`` restated. x (B,E) float32, weights (S,3,n).
    Returns (B,1) float32, differentiable w.r.t. ``x`` and ``weights`` by autograd."""
    B = x.shape[0]
    n = n_wires
    state = torch.zeros((B,) + (2,) * n, dtype=torch.complex64, device=x.device)
    state[(slice(None),) + (0,) * n] = 1.0
    col = 0
    s = 0
    for n_enc, depth in block_configs:
        for j in range(n_enc):
            if col < x.shape[1]:
                state = _apply_1q(state, j % n, _gate_matrix("rx", x[:, col]))
            col += 1
        for _ in range(depth):
            w = weights[s]
            for i in range(n):
                state = _apply_1q(state, i, _gate_matrix("ry", w[0, i].expand(B)))
                state = _apply_1q(state, i, _gate_matrix("rz", w[1, i].expand(B)))
                state = _apply_1q(state, i, _gate_matrix("ry", w[2, i].expand(B)))
            if n > 1:
                for i in range(n):
                    state = _apply_cnot(state, (i + 1) % n, i)
            s += 1
    probs = state.reshape(B, -1).abs().pow(2)
    if ham_diag is not None:
        d = torch.as_tensor(ham_diag, dtype=torch.float32, device=x.device)
        return (probs * d.unsqueeze(0)).sum(dim=1, keepdim=True)
    k = torch.arange(2 ** n, device=x.device)
    z_sum = torch.zeros(B, 1, device=x.device)
    for i in range(n):
        sign = 1 - 2 * ((k >> i) & 1).float()
        z_sum = z_sum + (probs * sign.unsqueeze(0)).sum(dim=1, keepdim=True)
    return ham_offset + ham_coeff * z_sum


def tq_forward_backward(x, weights, n_wires, block_configs, grad_out, need_grad_x=True, **ham):
    """Forward + autograd backward, the way ``loss.backward()`` drives the reference
    (``solvers/solver_pt.py:233-235``).  Returns (out, grad_x or None, grad_w)."""
    x = x.detach().clone().requires_grad_(need_grad_x)
    w = weights.detach().clone().requires_grad_(True)
    out = tq_forward(x, w, n_wires, block_configs, **ham)
    out.backward(grad_out.reshape_as(out))
    return out.detach(), (x.grad if need_grad_x else None), w.grad
