"""Peer-memory plumbing of the multi-GPU optimizer step (csrc/p2p.cu): one CUDA-IPC buffer per rank holding what the
peers touch — the flat fp32 gradient, the flat fp16 shadow of the parameters and a flag block — plus this process's
mappings of every peer's buffer.  torch.distributed is used once, to exchange the 64-byte IPC handles.

All ranks must live on one NVLink/NVSwitch box (one process per GPU).  ``PeerRegion.create`` returns None — on every
rank alike — when that is not the case or any rank fails to allocate / map; the trainer then keeps the NCCL path.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from ._lib import check, load

_KEEP = []   # regions are never unmapped while the process lives: a peer may still be reading ours


class _Raw:
    """torch.as_tensor() view of raw device memory (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class PeerRegion:
    def __init__(self):
        self.ptr = None

    @classmethod
    def create(cls, total: int, dev: torch.device, group=None):
        """total = elements of the flat parameter buffer.  Collective over ``group``; None if unavailable."""
        if dev.type != "cuda" or not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        L = load()
        if world < 2 or world > 8:   # NGP_MAX_PEERS
            return None
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        self = cls()
        self.rank, self.world, self.dev, self.total = rank, world, dev, int(total)
        self.off_shadow = _align(total * 4)
        self.off_flags = self.off_shadow + _align(total * 2)
        self.bytes = self.off_flags + _align(int(L.ngp_p2p_flag_bytes()))
        ok, handle = local_world == world, b""
        if ok:
            ptr, buf = C.c_void_p(), (C.c_uint8 * 64)()
            with torch.cuda.device(dev):
                rc = L.ngp_p2p_alloc(self.bytes, C.byref(ptr), buf)
            ok = rc == 0 and bool(ptr.value)
            if ok:
                self.ptr, handle = int(ptr.value), bytes(buf)
        infos = [None] * world
        dist.all_gather_object(infos, (ok, handle), group=group)
        ok = all(i[0] for i in infos)
        self.peer_ptrs = [0] * world
        if ok:
            with torch.cuda.device(dev):
                for p, (_, h) in enumerate(infos):
                    if p == rank:
                        self.peer_ptrs[p] = self.ptr
                        continue
                    q = C.c_void_p()
                    if L.ngp_p2p_open((C.c_uint8 * 64).from_buffer_copy(h), C.byref(q)) != 0 or not q.value:
                        ok = False
                        break
                    self.peer_ptrs[p] = int(q.value)
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=group)
        if not all(oks):
            self._release()
            return None
        _KEEP.append(self)
        # views of this rank's own buffer
        self.grad = torch.as_tensor(_Raw(self.ptr, self.total, "<f4"), device=dev)
        self.shadow = torch.as_tensor(_Raw(self.ptr + self.off_shadow, self.total, "<f2"), device=dev)
        self.flags = torch.as_tensor(_Raw(self.ptr + self.off_flags, int(L.ngp_p2p_flag_bytes()) // 4, "<i4"),
                                     device=dev)
        assert self.grad.data_ptr() == self.ptr and self.grad.dtype == torch.float32
        assert self.shadow.dtype == torch.float16 and self.flags.dtype == torch.int32
        self.epoch = torch.zeros(1, device=dev, dtype=torch.int32)
        arr = C.c_void_p * world
        self.grad_tab = arr(*[p for p in self.peer_ptrs])
        self.shadow_tab = arr(*[p + self.off_shadow for p in self.peer_ptrs])
        self.flag_tab = arr(*[p + self.off_flags for p in self.peer_ptrs])
        return self

    def _release(self):
        L = load()
        for p, q in enumerate(getattr(self, "peer_ptrs", [])):
            if q and p != self.rank:
                L.ngp_p2p_close(C.c_void_p(q))
        if self.ptr:
            L.ngp_p2p_free(C.c_void_p(self.ptr))
        self.ptr, self.peer_ptrs = None, []

    # ---- enqueue (current stream; graph-capturable) -------------------------------------------------------------
    def barrier(self, found_inf=None):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(load().ngp_p2p_barrier(self.flag_tab, self.rank, self.world, C.c_void_p(self.epoch.data_ptr()),
                                     None if found_inf is None else C.c_void_p(found_inf.data_ptr()), st))

    def timed_out(self) -> bool:
        """Host check (synchronises): did a barrier of this rank give up waiting for a peer?"""
        return int(self.flags[16]) != 0   # word 2 * NGP_MAX_PEERS
