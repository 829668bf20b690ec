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
    from .ops import _mse_backward_impl as hea_mse_backward      # same body as the custom op, no dispatcher
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
        # ONE flat parameter buffer and ONE flat gradient buffer; every p.data / p.grad is a view into them, in
        # the order the fused kernel writes its outputs: [ansatz | freq weights (circuit order: trunk, branch) |
        # freq biases | other parameters | model bias | sum of squared residuals | pad].  The kernel then writes
        # every gradient in place (no copies), the frequency vectors it reads are plain slices (no torch.cat),
        # and a single all-reduce over the flat buffer is the whole exchange step.
        q = model.quantum_layer
        layers = [l for l in ((model.trunk_freq, model.branch_freq) if self.is_onet else (model.freq,))
                  if isinstance(getattr(l, "weights", None), nn.Parameter)]
        bias_p = model.bias if isinstance(getattr(model, "bias", None), nn.Parameter) and model.bias.requires_grad else None
        head = [q.ansatz_weights] + [l.weights for l in layers] + [l.bias for l in layers]
        head = [p for p in head if p.requires_grad]
        seen = {id(p) for p in head} | ({id(bias_p)} if bias_p is not None else set())
        ordered = head + [p for p in self.params if id(p) not in seen] + ([bias_p] if bias_p is not None else [])
        assert len(ordered) == len(self.params)
        total = sum(p.numel() for p in ordered)
        p0 = self.params[0]
        self.flat_param = torch.cat([p.data.reshape(-1) for p in ordered])
        self.flat_grad = torch.zeros(total + 2, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in ordered:
            n_el = p.numel()
            p.data = self.flat_param[off:off + n_el].view_as(p)
            p.grad = self.flat_grad[off:off + n_el].view_as(p)
            off += n_el
        # [sum g (= dL/dbias), sum squared residuals] are adjacent: the kernel's `sums` output lands on them
        self.sums_off = total - 1 if bias_p is not None else total
        self.sse_idx = self.sums_off + 1
        self._freq_span = None
        if layers and all(l.weights.requires_grad and l.bias.requires_grad for l in layers):
            nw, ne = q.ansatz_weights.numel(), sum(l.weights.numel() for l in layers)
            self._freq_span = (nw, ne)            # fw = flat[nw : nw+ne], fb = flat[nw+ne : nw+2ne]
        opt_map = {"adam": torch.optim.Adam, "adamw": torch.optim.AdamW, "sgd": torch.optim.SGD,
                   "rmsprop": torch.optim.RMSprop}       # solvers/solver_pt.py:156-161
        kw = dict(optimizer_kwargs or {})
        if optimizer in ("adam", "adamw") and p0.is_cuda:
            kw.setdefault("fused", True)
        self.optimizer = opt_map[optimizer](self.params, lr=lr, **kw)
        if self.distributed:                               # replicas must start identical
            for p in model.parameters():
                dist.broadcast(p.data, src=0, group=process_group)
        self._views = [(p, p.data_ptr(), p.grad.data_ptr()) for p in ordered]
        # the exchange step: the library's one-kernel NVLink peer-memory all-reduce on CUDA/NCCL groups
        # (quanonet_b200/comm.py), torch.distributed's all_reduce elsewhere (gloo in the CPU tests)
        self._all_reduce = None
        if self.distributed:
            from .comm import make_all_reduce
            self._all_reduce = make_all_reduce(self.flat_grad, process_group)
        # whole-model fused kernel (frequency layers evaluated in-kernel, x / grad_x never materialised):
        # needs the register tier with one thread per sample and the standard n-angles-per-block layout
        self.kernel_events = None      # set to a list to collect (start, end) CUDA events around the kernel call
        self.fused_encoding = False
        self._enc_by_batch = None
        self._const_freq = None
        if kernel_fn is None and use_fused_encoding and p0.is_cuda:
            from .ops import encoded_supported
            q = model.quantum_layer
            standard = all(e == q.n_wires and d >= 1 for e, d in q.block_configs)
            self.fused_encoding = standard and encoded_supported(q.n_wires, p0.dtype)
            # n = 6..9 (fp32): fused-encoding kernels exist for small batches only (wide latency tier) — decided per step
            self._enc_by_batch = {} if standard and not self.fused_encoding and p0.dtype == torch.float32 \
                and 6 <= q.n_wires <= 9 else None
        # with the peer-memory exchange available, the fused step also does the all-reduce (one finalize kernel)
        from .comm import PeerAllReduce
        self._fused_exchange = ((self.fused_encoding or self._enc_by_batch is not None) and self.distributed
                                and p0.dtype == torch.float32
                                and isinstance(self._all_reduce, PeerAllReduce)
                                and (self._freq_span is not None or not getattr(model, "if_trainable_freq", False))
                                and q.ansatz_weights.requires_grad)

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
        fused = self.fused_encoding
        if not fused and self._enc_by_batch is not None:
            fused = self._enc_by_batch.get(B)
            if fused is None:
                from .ops import encoded_supported
                fused = self._enc_by_batch[B] = encoded_supported(q.n_wires, q.ansatz_weights.dtype, B,
                                                                  depths=[d for _, d in q.block_configs])
        if fused:
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
        self.flat_grad[self.sums_off] = g.sum()      # dL/dbias when the model has a bias (same slot), else a spare
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
        self.flat_grad[self.sse_idx] = (g * g).sum() / (scale * scale)   # sum of squared residuals on this shard
        if self.distributed and self.world_size > 1:
            self._all_reduce(self.flat_grad)
        return self.flat_grad[self.sse_idx] / gB

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
        from .ops import encoded_mse_step_into
        m = self.model
        q = m.quantum_layer
        depths = [d for _, d in q.block_configs]
        tf = bool(m.if_trainable_freq)
        fg = self.flat_grad
        if tf and self._freq_span is not None:
            nw, ne = self._freq_span
            fw, fb = self.flat_param[nw:nw + ne], self.flat_param[nw + ne:nw + 2 * ne]
            gfw, gfb = fg[nw:nw + ne], fg[nw + ne:nw + 2 * ne]
        else:
            if self._const_freq is None:
                self._const_freq = self._freq_vectors()
            (fw, fb), gfw, gfb = self._const_freq, None, None
            if tf:
                raise NotImplementedError("frozen frequency-layer parameters are not supported by the fused step")
        if self.is_onet:
            branch, trunk = inputs
            u0, u1, K0 = trunk, branch, m.trunk_enc_size // q.n_wires
        else:
            (u1,) = inputs
            u0, K0 = None, 0
        bias = m.bias if hasattr(m, "bias") else None
        if q.use_full_ham:
            ham = (q.ham_diag.to(device=u1.device, dtype=fw.dtype), _DIAG_ORDER[q.diag_order], 0.0, 0.0, _lib.QON_HAM_DIAG)
        else:
            ham = (None, _lib.QON_DIAG_LSB0, q.ham_offset, q.ham_coeff, _PAULI_KIND[q.ham_pauli])
        if self._fused_exchange:
            # compute step + exchange step in one pass: the finalize kernel pushes the gradients to the peers and
            # leaves the all-reduced flat gradient in fg (quanonet_b200/comm.py, csrc/qon_capi.cu)
            from .ops import encoded_mse_step_dp
            ev = self._event_start()
            nw = q.ansatz_weights.numel()
            ne = self._freq_span[1] if gfw is not None else 0
            encoded_mse_step_dp(u0, u1, fw, fb, K0, q.ansatz_weights.data, y.reshape(-1), bias, scale, q.n_wires, depths,
                                *ham, fg, 0, nw if gfw is not None else -1, nw + ne if gfw is not None else -1,
                                self.sums_off, self._all_reduce)
            self._event_end(ev)
            return fg[self.sse_idx] / gB
        ev = self._event_start()
        # every output is a view of the flat gradient buffer: grad_w, grad_fw, grad_fb, [dL/dbias, sum sq. residuals]
        encoded_mse_step_into(u0, u1, fw, fb, K0, q.ansatz_weights.data, y.reshape(-1), bias, scale, q.n_wires, depths,
                              *ham, q.ansatz_weights.grad, gfw, gfb, fg[self.sums_off:self.sums_off + 2])
        self._event_end(ev)
        if self.distributed and self.world_size > 1:
            self._all_reduce(fg)
        return fg[self.sse_idx] / gB

    def _check_views(self):
        """``model.to(...)`` / ``.double()`` / ``.cuda()`` after construction re-allocates the parameters and silently
        detaches them from the flat buffers the kernels and the optimiser work on: refuse to train on such a model."""
        for p, ptr, gptr in self._views:
            if p.data_ptr() != ptr or p.grad is None or p.grad.data_ptr() != gptr:
                raise RuntimeError(
                    "the model's parameters no longer alias the trainer's flat buffers (was the model moved or cast with "
                    ".to()/.cuda()/.double(), or were its gradients set to None, after DataParallelTrainer was built?). "
                    "Move the model first, then construct the trainer; use optimizer.zero_grad(set_to_none=False).")

    def step(self, inputs, y, global_batch=None):
        self._check_views()
        loss = self.compute_grads(inputs, y, global_batch)
        self.optimizer.step()
        return loss

    def run_host_fed(self, batches, global_batch=None, on_loss=None):
        """Train on batches that live in HOST memory: ``batches`` yields ``(inputs, y)`` with ``inputs`` a tuple of CPU
        tensors (pinned memory for asynchronous copies: ``tensor.pin_memory()``).  The copy of batch i+1 to the GPU
        runs on its own stream while batch i trains (two device slots); every step's loss is copied back to a pinned
        host scalar without stalling the stream.  Returns the list of per-step losses (floats) after one final sync —
        the data path of a dataset larger than HBM, and what bench.py's ``e2e`` times.  ``on_loss(i, pinned_scalar)``
        may be given to consume losses as they land."""
        dev = self.flat_param.device
        if dev.type != "cuda":
            raise RuntimeError("run_host_fed needs the model on a CUDA device")
        cur = torch.cuda.current_stream(dev)
        hf = getattr(self, "_hf", None)
        if hf is None:      # copy stream, events, device slots and the pinned loss landing zone live as long as the trainer
            hf = self._hf = {"stream": torch.cuda.Stream(device=dev), "slots": [None, None],
                             "ready": [torch.cuda.Event(), torch.cuda.Event()],
                             "consumed": [torch.cuda.Event(), torch.cuda.Event()],
                             "chunks": [torch.empty(1024, dtype=torch.float32).pin_memory()]}
        copy_stream, slots, ready, consumed = hf["stream"], hf["slots"], hf["ready"], hf["consumed"]
        for c in consumed:
            c.record(cur)
        losses = []

        def enqueue(slot, inputs, y):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])            # the step that last read this slot is done
                if slots[slot] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slots[slot], (*inputs, y))):
                    slots[slot] = tuple(torch.empty(h.shape, dtype=h.dtype, device=dev) for h in (*inputs, y))
                for d, h in zip(slots[slot], (*inputs, y)):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        it = iter(batches)
        nxt = next(it, None)
        i = 0
        if nxt is not None:
            enqueue(0, *nxt)
        while nxt is not None:
            slot = i % 2
            nxt = next(it, None)
            if nxt is not None:
                enqueue((i + 1) % 2, *nxt)
            cur.wait_event(ready[slot])
            *ins, yy = slots[slot]
            loss = self.step(tuple(ins), yy, global_batch)
            consumed[slot].record(cur)
            if i // 1024 >= len(hf["chunks"]):                     # pinned landing zone for the losses, 4 KB at a time
                hf["chunks"].append(torch.empty(1024, dtype=torch.float32).pin_memory())
            host = hf["chunks"][i // 1024][i % 1024:i % 1024 + 1]
            host.copy_(loss.detach().reshape(1).float(), non_blocking=True)
            losses.append(host)
            if on_loss is not None:
                on_loss(i, host)
            i += 1
        torch.cuda.synchronize(dev)
        return [float(h[0]) for h in losses]

    def exchange_timed_out(self) -> bool:
        """True if the peer-memory exchange ever gave up waiting for a rank (~30 s): the gradients of that step were
        NaN-poisoned on purpose, so the loss and the parameters are NaN from then on.  One host sync — poll it per
        epoch (``B200Solver`` does), not per step."""
        ar = self._all_reduce
        return bool(hasattr(ar, "timed_out") and ar.timed_out())


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
