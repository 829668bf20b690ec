"""Batch-sharded data-parallel training of QuanONetPT / HEAQNNPT on the B200 kernels.

Mirrors what ``PTSolver.train`` does per batch (reference ``solvers/solver_pt.py:225-241``:
forward, ``nn.MSELoss``, ``loss.backward()``, ``optimizer.step()``) with three differences that the
reference cannot have because it is single-process and autograd-driven:

* one fused kernel pass per step (``qon_hea_mse_forward_backward``): forward, MSE upstream
  gradient and adjoint backward, instead of forward + a second (recomputing) backward launch;
* the batch is sharded across ranks (one process per GPU); every rank produces a full-length
  partial gradient of the 2,401-ish parameters in ONE flat buffer, summed with a single
  ``all_reduce`` (NCCL over NVLink on GPUs, gloo in the CPU tests); the MSE mean uses the GLOBAL
  batch size, so the update equals the single-process update on the concatenated batch;
* no host synchronisation inside a step (the reference calls ``.item()`` twice per batch,
  ``solver_pt.py:238-241``); the loss stays a device scalar until the caller reads it.

Nothing here imports ``oracle``; CPU tests inject a stand-in for the kernel call through
``kernel_fn`` only to exercise the sharding / all-reduce / optimiser plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .core.models_pt import HEAQNNPT, QuanONetPT, _TiledElementWise, _tile_to

_PAULI_KIND = {"Z": _lib.QON_HAM_DIAG, "X": _lib.QON_HAM_PAULI_X, "Y": _lib.QON_HAM_PAULI_Y}
_DIAG_ORDER = {"lsb0": _lib.QON_DIAG_LSB0, "msb0": _lib.QON_DIAG_MSB0}


def _default_kernel(x, w, y, bias, grad_scale, qlayer, depths, need_gx):
    from .ops import hea_mse_backward
    if qlayer.use_full_ham:
        return hea_mse_backward(x, w, y, bias, grad_scale, qlayer.n_wires, depths,
                                qlayer.ham_diag.to(device=x.device, dtype=w.dtype),
                                _DIAG_ORDER[qlayer.diag_order], 0.0, 0.0, _lib.QON_HAM_DIAG, need_gx)
    return hea_mse_backward(x, w, y, bias, grad_scale, qlayer.n_wires, depths, None, _lib.QON_DIAG_LSB0,
                            qlayer.ham_offset, qlayer.ham_coeff, _PAULI_KIND[qlayer.ham_pauli], need_gx)


def _freq_grads(layer, u, gx):
    """Chain rule through ``enc[b,j] = u[b, j % in] * w[j] + bias[j]`` (``core/models_pt.py:38-41``)."""
    out, fin = layer.out_features, u.shape[1]
    if out % fin == 0:
        gw = (gx.view(gx.shape[0], out // fin, fin) * u.unsqueeze(1)).sum(0).reshape(out)
    else:
        gw = (gx * _tile_to(u, out)).sum(0)
    return gw, gx.sum(0)


class DataParallelTrainer:
    """One optimiser step per call on this rank's shard of the global batch.

    ``model``: ``QuanONetPT`` or ``HEAQNNPT`` (this package's or the reference's — same attributes).
    ``process_group``: a ``torch.distributed`` group, or None for single-process training.
    """

    def __init__(self, model: nn.Module, lr: float = 1e-3, optimizer: str = "adam", optimizer_kwargs=None,
                 process_group=None, kernel_fn: Optional[Callable] = None, use_fused_encoding: bool = True):
        self.model = model
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1
        self.rank = dist.get_rank(process_group) if self.distributed else 0
        self.kernel_fn = kernel_fn or _default_kernel
        self.is_onet = hasattr(model, "branch_freq")
        self.params = [p for p in model.parameters() if p.requires_grad]
        # one flat gradient buffer; every p.grad is a view into it -> a single all-reduce per step
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat_grad = torch.zeros(total + 1, dtype=p0.dtype, device=p0.device)   # last slot: sum of squared residuals
        off = 0
        for p in self.params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        opt_map = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD,
                   "rmsprop": torch.optim.RMSprop}       # solvers/solver_pt.py:156-161
        kw = dict(optimizer_kwargs or {})
        if optimizer in ("adam", "adamw") and p0.is_cuda:
            kw.setdefault("fused", True)
        self.optimizer = opt_map[optimizer](self.params, lr=lr, **kw)
        if self.distributed:                               # replicas must start identical
            for p in model.parameters():
                dist.broadcast(p.data, src=0, group=process_group)
        # whole-model fused kernel (frequency layers evaluated in-kernel, x / grad_x never materialised):
        # needs the register tier with one thread per sample and the standard n-angles-per-block layout
        self.kernel_events = None      # set to a list to collect (start, end) CUDA events around the kernel call
        self.fused_encoding = False
        if kernel_fn is None and use_fused_encoding and p0.is_cuda:
            from .ops import encoded_supported
            q = model.quantum_layer
            standard = all(e == q.n_wires and d >= 1 for e, d in q.block_configs)
            self.fused_encoding = standard and encoded_supported(q.n_wires, p0.dtype)

    # -- one step -------------------------------------------------------------------------------
    @torch.no_grad()
    def compute_grads(self, inputs: Sequence[torch.Tensor], y: torch.Tensor, global_batch: Optional[int] = None):
        """Fill ``p.grad`` (already all-reduced) for the MSE loss over the global batch; returns the
        global mean-squared error as a device scalar."""
        m = self.model
        q = m.quantum_layer
        B = y.shape[0]
        gB = global_batch if global_batch is not None else B * self.world_size
        scale = 2.0 / gB
        if self.fused_encoding:
            return self._compute_grads_fused(inputs, y, scale, gB)
        if self.is_onet:
            branch, trunk = inputs
            x = torch.cat([m.trunk_freq(trunk), m.branch_freq(branch)], dim=1)
        else:
            (u,) = inputs
            x = m.freq(u)
        xc, depths = q.canonical_inputs(x)
        if xc is not x and m.if_trainable_freq:
            raise NotImplementedError("trainable frequency layers need the standard block layout "
                                      "(n encoding angles per block) in the fused training step")
        need_gx = bool(m.if_trainable_freq)
        bias = m.bias if hasattr(m, "bias") else None
        ev = self._event_start()
        out, g, gx, gw = self.kernel_fn(xc, q.ansatz_weights, y.reshape(-1), bias, scale, q, depths, need_gx)
        self._event_end(ev)
        self.flat_grad.zero_()
        q.ansatz_weights.grad.copy_(gw)
        if bias is not None:
            m.bias.grad.copy_(g.sum().reshape(1))
        if need_gx:
            if self.is_onet:
                et = m.trunk_enc_size
                for layer, u_in, gpart in ((m.trunk_freq, trunk, gx[:, :et]), (m.branch_freq, branch, gx[:, et:])):
                    gw_f, gb_f = _freq_grads(layer, u_in, gpart)
                    layer.weights.grad.copy_(gw_f)
                    layer.bias.grad.copy_(gb_f)
            else:
                gw_f, gb_f = _freq_grads(m.freq, u, gx)
                m.freq.weights.grad.copy_(gw_f)
                m.freq.bias.grad.copy_(gb_f)
        self.flat_grad[-1] = (g * g).sum() / (scale * scale)          # sum of squared residuals on this shard
        if self.distributed and self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat_grad[-1] / gB

    def _event_start(self):
        if self.kernel_events is None:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return e0

    def _event_end(self, e0):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.kernel_events.append((e0, e1))

    def _freq_vectors(self):
        """(fw, fb) over all encoding columns in circuit order (trunk first, core/models_pt.py:163-164)."""
        m = self.model
        layers = (m.trunk_freq, m.branch_freq) if self.is_onet else (m.freq,)
        if m.if_trainable_freq:
            return torch.cat([l.weights for l in layers]), torch.cat([l.bias for l in layers])
        p0 = self.params[0]
        fw = torch.cat([torch.full((l.out_features,), float(l.scale), dtype=p0.dtype, device=p0.device) for l in layers])
        return fw, None

    def _compute_grads_fused(self, inputs, y, scale, gB):
        from .ops import encoded_mse_step
        m = self.model
        q = m.quantum_layer
        depths = [d for _, d in q.block_configs]
        fw, fb = self._freq_vectors()
        if self.is_onet:
            branch, trunk = inputs
            u0, u1, K0 = trunk, branch, m.trunk_enc_size // q.n_wires
        else:
            (u1,) = inputs
            u0, K0 = None, 0
        bias = m.bias if hasattr(m, "bias") else None
        tf = bool(m.if_trainable_freq)
        if q.use_full_ham:
            ham = (q.ham_diag.to(device=u1.device, dtype=fw.dtype), _DIAG_ORDER[q.diag_order], 0.0, 0.0, _lib.QON_HAM_DIAG)
        else:
            ham = (None, _lib.QON_DIAG_LSB0, q.ham_offset, q.ham_coeff, _PAULI_KIND[q.ham_pauli])
        ev = self._event_start()
        gw, gfw, gfb, sums = encoded_mse_step(u0, u1, fw, fb, K0, q.ansatz_weights, y.reshape(-1), bias, scale,
                                              q.n_wires, depths, *ham, tf)
        self._event_end(ev)
        self.flat_grad.zero_()
        q.ansatz_weights.grad.copy_(gw)
        if bias is not None:
            m.bias.grad.copy_(sums[0:1])
        if tf:
            off = 0
            for layer in ((m.trunk_freq, m.branch_freq) if self.is_onet else (m.freq,)):
                layer.weights.grad.copy_(gfw[off:off + layer.out_features])
                layer.bias.grad.copy_(gfb[off:off + layer.out_features])
                off += layer.out_features
        self.flat_grad[-1] = sums[1]
        if self.distributed and self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat_grad[-1] / gB

    def step(self, inputs, y, global_batch=None):
        loss = self.compute_grads(inputs, y, global_batch)
        self.optimizer.step()
        return loss


def autograd_step(model, optimizer, inputs, y, loss_fn=None):
    """The reference's own per-batch sequence (``solvers/solver_pt.py:231-236``) — works on the
    drop-in module through the registered autograd of ``quanonet::hea_expval``."""
    loss_fn = loss_fn or nn.MSELoss()
    optimizer.zero_grad()
    pred = model(*inputs)
    loss = loss_fn(pred, y)
    loss.backward()
    optimizer.step()
    return loss.detach()
