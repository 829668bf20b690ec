"""Exchange step of the data-parallel training step: sum all-reduce of the flat gradient buffer.

``PeerAllReduce`` runs the library's one-kernel NVLink peer-memory all-reduce (``csrc/qon_peer.cuh``) over a
symmetric buffer that torch's symmetric memory allocates and exchanges between the ranks — torch is plumbing
here (allocation + handle exchange); the data path is the library's own kernel.  ``make_all_reduce`` picks it
for CUDA/NCCL process groups on one node (world <= 8) and otherwise returns ``dist.all_reduce`` (gloo CPU tests,
or ``QON_COLLECTIVE=nccl``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from . import _lib


class PeerAllReduce:
    """In-place sum all-reduce of a contiguous fp32 CUDA tensor with at most ``max_len`` elements."""

    def __init__(self, max_len: int, device: torch.device, group=None, connect: bool = True):
        """Construction has a LOCAL half (allocation, capability checks — may raise on one rank only) and a
        COLLECTIVE half (``connect``: handle exchange + barrier).  ``make_all_reduce`` runs the local half on every
        rank, agrees on the outcome with an all-reduce(MIN), and only then enters the collective half — so a rank
        whose allocation fails never leaves its peers blocked in a rendezvous."""
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.max_len = int(max_len)
        self.device = device
        self.lib = _lib.load()
        nbytes = int(self.lib.qon_peer_buffer_bytes(self.max_len, self.world))
        if nbytes == 0:
            raise RuntimeError(f"peer all-reduce supports world <= 8 (got {self.world})")
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.ptrs = None
        if connect:
            self.connect()

    def connect(self):
        """Collective: every rank of the group must call it."""
        import torch.distributed._symmetric_memory as symm_mem

        self.handle = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != self.buf.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned unexpected peer pointers")
        self.ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # every rank's buffer is zeroed before anyone pushes

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        if flat.dtype != torch.float32 or not flat.is_cuda or not flat.is_contiguous():
            raise ValueError("peer all-reduce takes a contiguous float32 CUDA tensor")
        n = flat.numel()
        if n > self.max_len:
            raise ValueError(f"{n} elements > max_len {self.max_len}")
        rc = self.lib.qon_peer_allreduce_f32(flat.data_ptr(), flat.data_ptr(), n, self.ptrs, self.world, self.rank,
                                             self.max_len, torch.cuda.current_stream(flat.device).cuda_stream)
        if rc != 0:
            raise RuntimeError(f"qon_peer_allreduce_f32 failed ({rc}): {_lib.last_error()}")
        return flat

    def timed_out(self) -> bool:
        """True if any call so far gave up waiting for a peer (its output was NaN-poisoned)."""
        return bool(self.buf[132:136].view(torch.int32).item() != 0)


def make_all_reduce(flat: torch.Tensor, group=None) -> Callable[[torch.Tensor], Optional[torch.Tensor]]:
    """The all-reduce the trainer calls on its flat gradient buffer."""
    def nccl(t):
        dist.all_reduce(t, group=group)
        return t

    choice = os.environ.get("QON_COLLECTIVE", "auto").lower()
    if choice == "nccl" or not flat.is_cuda or flat.dtype != torch.float32:
        return nccl
    if dist.get_backend(group) != "nccl" or dist.get_world_size(group) > 8:
        return nccl
    peer, err = None, None
    try:
        peer = PeerAllReduce(flat.numel(), flat.device, group, connect=False)     # local half only
    except Exception as exc:    # symmetric memory unavailable (no P2P, older driver ...)
        err = exc
    # every rank must take the same path: agree on the outcome BEFORE the collective half of the construction
    ok = torch.tensor([1 if peer is not None else 0], device=flat.device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 1:
        peer.connect()
        return peer
    if choice == "peer":
        raise RuntimeError(f"peer-memory all-reduce unavailable on at least one rank ({err})")
    import warnings
    warnings.warn(f"peer-memory all-reduce unavailable on at least one rank ({err}); NCCL does the exchange")
    return nccl
