"""torch custom ops over the C-ABI: ``quanonet::hea_expval`` and ``quanonet::hea_expval_backward``.

PyTorch is plumbing here — device memory, the current stream, autograd wiring.  The arithmetic is
the hand-written CUDA in csrc/, reached through ``_lib`` (ctypes).  CUDA tensors only; a CPU tensor
raises (no fallback).

Inputs are in the C-ABI's canonical form (see include/quanonet_b200.h): ``x (B, n*K)``,
``weights (S,3,n)``, ``depth_per_block`` (K ints >= 1).  The reference-facing module
(core/quantum_circuits_tq.py) canonicalises arbitrary ``block_configs`` before calling in.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.QON_F32, torch.float64: _lib.QON_F64}


def _check_inputs(x, weights, n_wires, depth_per_block, ham_diag):
    if not x.is_cuda:
        raise RuntimeError("quanonet::hea_expval runs on CUDA tensors only (there is no CPU fallback); "
                           f"got x on {x.device}")
    if x.dtype not in _DTYPES:
        raise TypeError(f"x must be float32 or float64, got {x.dtype}")
    if weights.dtype != x.dtype or weights.device != x.device:
        raise TypeError("weights must match x in dtype and device")
    K = len(depth_per_block)
    S = int(sum(depth_per_block))
    if x.dim() != 2 or x.shape[1] != n_wires * K:
        raise ValueError(f"x must be (B, n*K) = (B, {n_wires * K}), got {tuple(x.shape)}")
    if tuple(weights.shape) != (S, 3, n_wires):
        raise ValueError(f"weights must be (S,3,n) = ({S},3,{n_wires}), got {tuple(weights.shape)}")
    if ham_diag is not None:
        if ham_diag.numel() != (1 << n_wires):
            raise ValueError(f"ham_diag must have 2**n = {1 << n_wires} entries, got {ham_diag.numel()}")


def _rowmajor(x):
    """Rows contiguous (stride(1) == 1, or a single column); returns (tensor, row_stride)."""
    if x.shape[0] <= 1 or x.shape[1] == 0:
        x = x.contiguous()
        return x, max(int(x.shape[1]), 1)
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    return x, int(x.stride(0))


def _workspace(B, n, depth, dtype_code, need_grad, device):
    lib = _lib.load()
    nbytes = lib.qon_workspace_bytes(B, n, len(depth), depth, dtype_code, int(need_grad))
    if nbytes == 0:
        raise RuntimeError(f"qon_workspace_bytes failed: {_lib.last_error()}")
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


@torch.library.custom_op("quanonet::hea_expval", mutates_args=())
def hea_expval(x: torch.Tensor, weights: torch.Tensor, n_wires: int, depth_per_block: List[int],
               ham_diag: Optional[torch.Tensor], diag_order: int, ham_offset: float, ham_coeff: float,
               ham_kind: int) -> torch.Tensor:
    return _forward_impl(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind)


def _forward_impl(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind):
    _check_inputs(x, weights, n_wires, depth_per_block, ham_diag)
    lib = _lib.load()
    code = _DTYPES[x.dtype]
    B = x.shape[0]
    out = torch.empty((B, 1), dtype=x.dtype, device=x.device)
    if B == 0:
        return out
    with torch.cuda.device(x.device):
        xc, ldx = _rowmajor(x)
        wc = weights.contiguous()
        hd = None if ham_diag is None else ham_diag.to(dtype=x.dtype, device=x.device).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, code, False, x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = lib.qon_hea_forward(xc.data_ptr(), ldx, wc.data_ptr(), out.data_ptr(), B, n_wires, len(depth_per_block),
                                 depth, None if hd is None else hd.data_ptr(), diag_order, ham_offset, ham_coeff,
                                 ham_kind, code, ws.data_ptr(), nbytes, stream)
        _lib.check(rc, "qon_hea_forward")
    return out


@hea_expval.register_fake
def _(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind):
    return x.new_empty((x.shape[0], 1))


@torch.library.custom_op("quanonet::hea_expval_backward", mutates_args=())
def hea_expval_backward(grad_out: torch.Tensor, x: torch.Tensor, weights: torch.Tensor, n_wires: int,
                        depth_per_block: List[int], ham_diag: Optional[torch.Tensor], diag_order: int,
                        ham_offset: float, ham_coeff: float, ham_kind: int,
                        need_grad_x: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One fused forward + adjoint-backward pass.  Returns (out (B,1), grad_x (B,n*K) or empty, grad_w)."""
    return _backward_impl(grad_out, x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff,
                          ham_kind, need_grad_x)


def _backward_impl(grad_out, x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff,
                   ham_kind, need_grad_x):
    _check_inputs(x, weights, n_wires, depth_per_block, ham_diag)
    lib = _lib.load()
    code = _DTYPES[x.dtype]
    B = x.shape[0]
    out = torch.empty((B, 1), dtype=x.dtype, device=x.device)
    grad_w = torch.empty_like(weights, memory_format=torch.contiguous_format)
    grad_x = torch.empty((B, x.shape[1]) if need_grad_x else (0,), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        xc, ldx = _rowmajor(x)
        wc = weights.contiguous()
        g = grad_out.to(dtype=x.dtype).reshape(-1).contiguous()
        if g.numel() != B:
            raise ValueError(f"grad_out must have B = {B} elements, got {g.numel()}")
        hd = None if ham_diag is None else ham_diag.to(dtype=x.dtype, device=x.device).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, code, True, x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = lib.qon_hea_forward_backward(
            xc.data_ptr(), ldx, wc.data_ptr(), g.data_ptr(), out.data_ptr(),
            grad_x.data_ptr() if need_grad_x else None, max(int(x.shape[1]), 1), grad_w.data_ptr(),
            B, n_wires, len(depth_per_block), depth, None if hd is None else hd.data_ptr(), diag_order,
            ham_offset, ham_coeff, ham_kind, code, ws.data_ptr(), nbytes, stream)
        _lib.check(rc, "qon_hea_forward_backward")
    return out, grad_x, grad_w


@hea_expval_backward.register_fake
def _(grad_out, x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind,
      need_grad_x):
    gx = x.new_empty(x.shape if need_grad_x else (0,))
    return x.new_empty((x.shape[0], 1)), gx, torch.empty_like(weights)


def _setup_context(ctx, inputs, output):
    x, weights, n_wires, depth, ham_diag, diag_order, off, coeff, kind = inputs
    ctx.save_for_backward(x, weights, ham_diag if ham_diag is not None else x.new_empty(0))
    ctx.has_diag = ham_diag is not None
    ctx.cfg = (n_wires, list(depth), diag_order, off, coeff, kind)


def _backward(ctx, grad_out):
    x, weights, hd = ctx.saved_tensors
    n_wires, depth, diag_order, off, coeff, kind = ctx.cfg
    need_gx = ctx.needs_input_grad[0]
    _, gx, gw = hea_expval_backward(grad_out.contiguous(), x, weights, n_wires, depth,
                                    hd if ctx.has_diag else None, diag_order, off, coeff, kind, need_gx)
    return (gx if need_gx else None), (gw if ctx.needs_input_grad[1] else None), None, None, None, None, None, None, None


hea_expval.register_autograd(_backward, setup_context=_setup_context)


class _HeaExpvalFn(torch.autograd.Function):
    """Eager-mode twin of ``quanonet::hea_expval`` + its registered autograd: the same two C-ABI calls without the
    custom-op dispatcher, which costs ~150 us per differentiable call with a 60-entry ``depth_per_block`` —
    more than the 69 us kernel of a 100-sample batch."""

    @staticmethod
    def forward(ctx, x, weights, n_wires, depth, ham_diag, diag_order, off, coeff, kind):
        ctx.save_for_backward(x, weights)
        ctx.ham_diag = ham_diag
        ctx.cfg = (n_wires, depth, diag_order, off, coeff, kind)
        return _forward_impl(x, weights, n_wires, depth, ham_diag, diag_order, off, coeff, kind)

    @staticmethod
    def backward(ctx, grad_out):
        x, weights = ctx.saved_tensors
        n_wires, depth, diag_order, off, coeff, kind = ctx.cfg
        need_gx = ctx.needs_input_grad[0]
        _, gx, gw = _backward_impl(grad_out.contiguous(), x, weights, n_wires, depth, ctx.ham_diag, diag_order, off,
                                   coeff, kind, need_gx)
        return (gx if need_gx else None), (gw if ctx.needs_input_grad[1] else None), None, None, None, None, None, None, None


def hea_expval_autograd(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind):
    """What the drop-in module calls: the custom op while torch.compile traces, the plain autograd Function in
    eager mode (identical results; see ``_HeaExpvalFn``)."""
    if torch.compiler.is_compiling():
        return hea_expval(x, weights, n_wires, list(depth_per_block), ham_diag, diag_order, ham_offset, ham_coeff, ham_kind)
    if not (torch.is_grad_enabled() and (x.requires_grad or weights.requires_grad)):
        return _forward_impl(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind)
    return _HeaExpvalFn.apply(x, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind)


def fp32_peak_tflops(iters: int = 2000) -> float:
    """Measured FFMA throughput of the current device (TFLOP/s) — the '% of FP32 peak' denominator."""
    lib = _lib.load()
    v = lib.qon_measure_fp32_peak_tflops(int(iters), torch.cuda.current_stream().cuda_stream)
    if v <= 0:
        raise RuntimeError(f"FP32 peak probe failed: {_lib.last_error()}")
    return float(v)


def tensor_tier(enable: Optional[bool] = None, min_batch: Optional[int] = None, debug_state=None, error_flag=None) -> bool:
    """Switch / query the tensor-core tier (``qon_tensor_tier``): n = 5, fp32, diagonal observables, batches of at
    least ``min_batch`` samples.  Returns the previous setting.  ``debug_state`` / ``error_flag`` are device tensors
    used by the bring-up scripts only."""
    prev = _lib.load().qon_tensor_tier(-1 if enable is None else int(bool(enable)) if not isinstance(enable, int) else int(enable),
                                       -1 if min_batch is None else int(min_batch),
                                       None if debug_state is None else debug_state.data_ptr(),
                                       None if error_flag is None else error_flag.data_ptr())
    return bool(prev)


def latency_tier_max_batch() -> int:
    """Largest per-call batch served by the small-batch latency tier (n <= 5, angles given)."""
    return int(_lib.load().qon_latency_tier_max_batch())


def plan_tier(B: int, n: int, dtype=torch.float32, need_grad=True):
    """(tier, lanes_log2) the library would use: tier 0 register, 1 shared memory, 2 HBM-streamed."""
    import ctypes
    lib = _lib.load()
    lq = ctypes.c_int(-1)
    t = lib.qon_plan_tier(B, n, _DTYPES[dtype], int(need_grad), ctypes.byref(lq))
    if t < 0:
        raise RuntimeError(_lib.last_error())
    return t, lq.value


@torch.library.custom_op("quanonet::hea_mse_backward", mutates_args=())
def hea_mse_backward(x: torch.Tensor, weights: torch.Tensor, target: torch.Tensor, bias: Optional[torch.Tensor],
                     grad_scale: float, n_wires: int, depth_per_block: List[int], ham_diag: Optional[torch.Tensor],
                     diag_order: int, ham_offset: float, ham_coeff: float, ham_kind: int,
                     need_grad_x: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Training-step kernel: forward, MSE upstream gradient ``g = grad_scale*(out+bias-target)`` and
    adjoint backward in ONE pass (``qon_hea_mse_forward_backward``).
    Returns (out (B,1) without bias, g (B,), grad_x (B,n*K) or empty, grad_w (S,3,n))."""
    return _mse_backward_impl(x, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag, diag_order,
                              ham_offset, ham_coeff, ham_kind, need_grad_x)


def _mse_backward_impl(x, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag, diag_order, ham_offset,
                       ham_coeff, ham_kind, need_grad_x):
    """Body of ``quanonet::hea_mse_backward``; the trainer calls it directly (no dispatcher overhead)."""
    _check_inputs(x, weights, n_wires, depth_per_block, ham_diag)
    lib = _lib.load()
    code = _DTYPES[x.dtype]
    B = x.shape[0]
    out = torch.empty((B, 1), dtype=x.dtype, device=x.device)
    g = torch.empty((B,), dtype=x.dtype, device=x.device)
    grad_w = torch.empty_like(weights, memory_format=torch.contiguous_format)
    grad_x = torch.empty((B, x.shape[1]) if need_grad_x else (0,), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        xc, ldx = _rowmajor(x)
        wc = weights.contiguous()
        y = target.to(dtype=x.dtype).reshape(-1).contiguous()
        if y.numel() != B:
            raise ValueError(f"target must have B = {B} elements, got {y.numel()}")
        bptr = None
        if bias is not None:
            bc = bias.to(dtype=x.dtype).reshape(-1).contiguous()
            bptr = bc.data_ptr()
        hd = None if ham_diag is None else ham_diag.to(dtype=x.dtype, device=x.device).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, code, True, x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        rc = lib.qon_hea_mse_forward_backward(
            xc.data_ptr(), ldx, wc.data_ptr(), y.data_ptr(), bptr, float(grad_scale), out.data_ptr(), g.data_ptr(),
            grad_x.data_ptr() if need_grad_x else None, max(int(x.shape[1]), 1), grad_w.data_ptr(),
            B, n_wires, len(depth_per_block), depth, None if hd is None else hd.data_ptr(), diag_order,
            ham_offset, ham_coeff, ham_kind, code, ws.data_ptr(), nbytes, stream)
        _lib.check(rc, "qon_hea_mse_forward_backward")
    return out, g, grad_x, grad_w


@hea_mse_backward.register_fake
def _(x, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff,
      ham_kind, need_grad_x):
    gx = x.new_empty(x.shape if need_grad_x else (0,))
    return x.new_empty((x.shape[0], 1)), x.new_empty((x.shape[0],)), gx, torch.empty_like(weights)


# ------------------------------------------------------------------------------------------------
# fused encoding: frequency layers evaluated inside the kernel (x / grad_x never materialised)
# ------------------------------------------------------------------------------------------------
def _enc_args(u0, u1, fw, fb, K0, n_wires, depth_per_block):
    if not u1.is_cuda:
        raise RuntimeError("quanonet encoded ops run on CUDA tensors only (there is no CPU fallback)")
    if u1.dtype not in _DTYPES:
        raise TypeError(f"inputs must be float32 or float64, got {u1.dtype}")
    E = n_wires * len(depth_per_block)
    if fw.numel() != E or (fb is not None and fb.numel() != E):
        raise ValueError(f"fw / fb must have n*K = {E} entries")
    if u0 is None and K0 != 0:
        raise ValueError("K0 > 0 needs a source-0 input")
    u1c, ld1 = _rowmajor(u1)
    if u0 is not None:
        u0c, ld0 = _rowmajor(u0.to(u1.dtype))
        head = [u0c.data_ptr(), ld0, int(u0c.shape[1]), int(K0)]
    else:
        u0c, head = None, [None, 0, 0, 0]
    fwc = fw.to(u1.dtype).contiguous()
    fbc = None if fb is None else fb.to(u1.dtype).contiguous()
    head += [u1c.data_ptr(), ld1, int(u1c.shape[1]), fwc.data_ptr(), None if fbc is None else fbc.data_ptr()]
    return head, (u0c, u1c, fwc, fbc)


@torch.library.custom_op("quanonet::encoded_expval", mutates_args=())
def encoded_expval(u0: Optional[torch.Tensor], u1: torch.Tensor, fw: torch.Tensor, fb: Optional[torch.Tensor],
                   K0: int, weights: torch.Tensor, n_wires: int, depth_per_block: List[int],
                   ham_diag: Optional[torch.Tensor], diag_order: int, ham_offset: float, ham_coeff: float,
                   ham_kind: int) -> torch.Tensor:
    """Forward of the whole model body (frequency layers + circuit) in one kernel: ``qon_encoded_forward``."""
    lib = _lib.load()
    B = u1.shape[0]
    out = torch.empty((B, 1), dtype=u1.dtype, device=u1.device)
    if B == 0:
        return out
    with torch.cuda.device(u1.device):
        head, keep = _enc_args(u0, u1, fw, fb, K0, n_wires, depth_per_block)
        code = _DTYPES[u1.dtype]
        wc = weights.to(u1.dtype).contiguous()
        hd = None if ham_diag is None else ham_diag.to(dtype=u1.dtype, device=u1.device).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, code, False, u1.device)
        rc = lib.qon_encoded_forward(*head, wc.data_ptr(), out.data_ptr(), B, n_wires, len(depth_per_block), depth,
                                     None if hd is None else hd.data_ptr(), diag_order, ham_offset, ham_coeff, ham_kind,
                                     code, ws.data_ptr(), nbytes, torch.cuda.current_stream(u1.device).cuda_stream)
        _lib.check(rc, "qon_encoded_forward")
    return out


@encoded_expval.register_fake
def _(u0, u1, fw, fb, K0, weights, n_wires, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind):
    return u1.new_empty((u1.shape[0], 1))


@torch.library.custom_op("quanonet::encoded_mse_step", mutates_args=())
def encoded_mse_step(u0: Optional[torch.Tensor], u1: torch.Tensor, fw: torch.Tensor, fb: Optional[torch.Tensor],
                     K0: int, weights: torch.Tensor, target: torch.Tensor, bias: Optional[torch.Tensor],
                     grad_scale: float, n_wires: int, depth_per_block: List[int], ham_diag: Optional[torch.Tensor],
                     diag_order: int, ham_offset: float, ham_coeff: float, ham_kind: int,
                     need_freq_grad: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``qon_encoded_mse_step``: returns (grad_w (S,3,n), grad_fw (E,) or empty, grad_fb (E,) or empty,
    sums (2,) = [sum g, sum squared residual])."""
    dt, dev = u1.dtype, u1.device
    E = n_wires * len(depth_per_block)
    grad_w = torch.empty(weights.shape, dtype=dt, device=dev)
    gfw = torch.empty((E,) if need_freq_grad else (0,), dtype=dt, device=dev)
    gfb = torch.empty((E,) if need_freq_grad else (0,), dtype=dt, device=dev)
    sums = torch.empty((2,), dtype=dt, device=dev)
    encoded_mse_step_into(u0, u1, fw, fb, K0, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag,
                          diag_order, ham_offset, ham_coeff, ham_kind, grad_w, gfw if need_freq_grad else None,
                          gfb if need_freq_grad else None, sums)
    return grad_w, gfw, gfb, sums


def encoded_mse_step_into(u0, u1, fw, fb, K0, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag,
                          diag_order, ham_offset, ham_coeff, ham_kind, grad_w, grad_fw, grad_fb, sums):
    """``qon_encoded_mse_step`` writing into caller-owned tensors (e.g. views of the trainer's flat gradient
    buffer, so a training step needs no copies): grad_w (S,3,n), grad_fw / grad_fb (E,) or None, sums (2,)."""
    lib = _lib.load()
    B = u1.shape[0]
    dt, dev = u1.dtype, u1.device
    for name, t in (("grad_w", grad_w), ("grad_fw", grad_fw), ("grad_fb", grad_fb), ("sums", sums)):
        if t is not None and (t.dtype != dt or t.device != dev or not t.is_contiguous()):
            raise ValueError(f"{name} must be a contiguous {dt} tensor on {dev}")
    with torch.cuda.device(dev):
        head, keep = _enc_args(u0, u1, fw, fb, K0, n_wires, depth_per_block)
        code = _DTYPES[dt]
        wc = weights.to(dt).contiguous()
        y = target.to(dt).reshape(-1).contiguous()
        if y.numel() != B:
            raise ValueError(f"target must have B = {B} elements, got {y.numel()}")
        bc = None if bias is None else bias.to(dt).reshape(-1).contiguous()
        hd = None if ham_diag is None else ham_diag.to(dtype=dt, device=dev).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, code, True, dev)
        rc = lib.qon_encoded_mse_step(
            *head, wc.data_ptr(), y.data_ptr(), None if bc is None else bc.data_ptr(), float(grad_scale), None,
            grad_w.data_ptr(), None if grad_fw is None else grad_fw.data_ptr(),
            None if grad_fb is None else grad_fb.data_ptr(), sums.data_ptr(), B, n_wires, len(depth_per_block), depth,
            None if hd is None else hd.data_ptr(), diag_order, ham_offset, ham_coeff, ham_kind, code, ws.data_ptr(),
            nbytes, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "qon_encoded_mse_step")


def encoded_mse_step_dp(u0, u1, fw, fb, K0, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag,
                        diag_order, ham_offset, ham_coeff, ham_kind, flat, w_off, fw_off, fb_off, sums_off, peer):
    """``qon_encoded_mse_step_dp``: the training step's gradient AND its all-reduce over the ranks in one pass —
    the finalize kernel pushes every gradient into the peers' symmetric buffers and the summed result lands in
    ``flat`` (fp32).  ``peer`` is a ``quanonet_b200.comm.PeerAllReduce``; ``fw_off = fb_off = -1`` for fixed
    frequency layers."""
    lib = _lib.load()
    B = u1.shape[0]
    dt, dev = u1.dtype, u1.device
    if dt != torch.float32 or flat.dtype != torch.float32 or not flat.is_contiguous() or flat.device != dev:
        raise ValueError("the fused exchange takes float32 inputs and a contiguous float32 flat buffer on the same device")
    if flat.numel() > peer.max_len:
        raise ValueError(f"flat buffer ({flat.numel()}) exceeds the peer buffers' max_len ({peer.max_len})")
    with torch.cuda.device(dev):
        head, keep = _enc_args(u0, u1, fw, fb, K0, n_wires, depth_per_block)
        wc = weights.to(dt).contiguous()
        y = target.to(dt).reshape(-1).contiguous()
        if y.numel() != B:
            raise ValueError(f"target must have B = {B} elements, got {y.numel()}")
        bc = None if bias is None else bias.to(dt).reshape(-1).contiguous()
        hd = None if ham_diag is None else ham_diag.to(dtype=dt, device=dev).contiguous()
        depth = _lib.int_array(depth_per_block)
        ws, nbytes = _workspace(B, n_wires, depth, _DTYPES[dt], True, dev)
        rc = lib.qon_encoded_mse_step_dp(
            *head, wc.data_ptr(), y.data_ptr(), None if bc is None else bc.data_ptr(), float(grad_scale), None,
            flat.data_ptr(), flat.numel(), int(w_off), int(fw_off), int(fb_off), int(sums_off), peer.ptrs, peer.world,
            peer.rank, peer.max_len, B, n_wires, len(depth_per_block), depth, None if hd is None else hd.data_ptr(),
            diag_order, ham_offset, ham_coeff, ham_kind, ws.data_ptr(), nbytes,
            torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "qon_encoded_mse_step_dp")


@encoded_mse_step.register_fake
def _(u0, u1, fw, fb, K0, weights, target, bias, grad_scale, n_wires, depth_per_block, ham_diag, diag_order,
      ham_offset, ham_coeff, ham_kind, need_freq_grad):
    E = n_wires * len(depth_per_block)
    g = u1.new_empty((E,) if need_freq_grad else (0,))
    return torch.empty_like(weights), g, u1.new_empty(g.shape), u1.new_empty((2,))


def encoded_supported(n_wires: int, dtype=torch.float32, batch: Optional[int] = None, need_grad: bool = True,
                      depths: Optional[List[int]] = None) -> bool:
    """Whether the fused-encoding kernels serve this problem: always for n <= 5 (fp32) / n <= 4 (fp64); for
    n = 6..9 in fp32 only when ``batch`` is given and small enough for the wide latency tier — and, because that
    tier's shared-memory footprint grows with the circuit, only if the planner accepts the real circuit
    (``depths``; ``qon_encoded_supported_for``).  Without ``depths`` a one-block probe is planned."""
    if n_wires <= (5 if dtype == torch.float32 else 4):
        return True
    if batch is None or dtype != torch.float32 or not 6 <= n_wires <= 9:
        return False
    lib = _lib.load()
    if depths:
        return bool(lib.qon_encoded_supported_for(int(batch), int(n_wires), len(depths), _lib.int_array(depths),
                                                  _DTYPES[dtype], int(need_grad)))
    return bool(lib.qon_encoded_supported(int(batch), int(n_wires), _DTYPES[dtype], int(need_grad)))
