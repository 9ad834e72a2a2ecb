"""The communicator of ``--ddp-backend b200``: symmetric buffers + the hand-written collective kernels.

One ``SymmComm`` per process group owns

* the symmetric flag buffer of the cross-GPU barriers (``csrc/comm/comm_device.cuh``),
* the error channel: four words of pinned host memory the kernels write when a peer never arrives or two ranks pair
  up on different collectives (instead of trapping the context); ``check_health()`` turns that into a Python exception,
* two symmetric exchange buffers (fused optimizer tail / stand-alone statistics) of ``2 x world x 65`` doubles,

and launches the kernels of ``csrc/comm/allreduce.cu`` (one-shot / two-shot / NVLS all-reduce and their
reduce-scatter halves) and ``csrc/comm/fused_step.cu`` (statistics reduction, fused tail).  NCCL is not involved in
any of them; the process group is only used while buffers are being allocated (descriptor exchange).
"""
import logging
from typing import Optional

import torch
import torch.distributed as dist

from . import symm_mem

logger = logging.getLogger(__name__)

_DTYPE_TAG = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}

# tags of the flag protocol (never zero): what a rank is waiting for
TAG_BUCKET = 0x100      # + bucket index
TAG_TAIL = 0x7001
TAG_STATS = 0x7002
TAG_MISC = 0x7003


class CommunicatorError(RuntimeError):
    pass


class SymmComm:
    def __init__(self, group=None):
        from unicore_b200.ops._native import native

        self.native = native()
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.max_blocks = int(self.native.SYMM_MAX_BLOCKS)
        self.max_peers = int(self.native.SYMM_MAX_PEERS)
        self.row = int(self.native.SYMM_XCHG_DOUBLES_PER_RANK)
        self.max_stats = int(self.native.SYMM_MAX_STATS)
        self.flags = symm_mem.allocate(self.max_blocks * self.max_peers, torch.int32, self.device, self.group)
        self.rank, self.world = self.flags.rank, self.flags.world
        self.err_host, self.err_dev = self.native.symm_error_channel()
        self.xchg_tail = symm_mem.allocate(2 * self.world * self.row, torch.float64, self.device, self.group)
        self.xchg_stats = symm_mem.allocate(2 * self.world * self.row, torch.float64, self.device, self.group)
        self._stats_calls = 0
        self._stats_out = torch.zeros(self.max_stats, dtype=torch.float64, device=self.device)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # flags are zero everywhere before the first kernel spins on them
        self.provider = self.flags.provider

    # -- memory --------------------------------------------------------------------------------------------
    def allocate(self, numel: int, dtype: torch.dtype) -> symm_mem.SymmBuffer:
        return symm_mem.allocate(numel, dtype, self.device, self.group)

    # -- health ----------------------------------------------------------------------------------------------
    def check_health(self) -> None:
        """Raise if a collective kernel reported a dead peer / mismatched collective (reads host memory: free)."""
        code = int(self.err_host[0])
        if code == 0:
            return
        detail, who = int(self.err_host[1]) & 0xFFFFFFFF, int(self.err_host[2])
        if code == 1:
            msg = "rank {} waited ~10 s for a peer in collective tag {:#x}: a rank died or never reached it".format(who, detail)
        else:
            msg = ("rank {} met a peer that is inside a different collective (mine {:#x}, theirs {:#x}): the ranks "
                   "launched their gradient buckets in different orders".format(who, detail >> 16, detail & 0xFFFF))
        raise CommunicatorError("b200 communicator failed: " + msg)

    # -- gradient-path collectives -------------------------------------------------------------------------------
    def _range(self, buf, elem_offset, numel):
        t = buf.tensor
        numel = t.numel() - elem_offset if numel is None else numel
        esz = t.element_size()
        byte_off, nbytes = elem_offset * esz, numel * esz
        if byte_off % 16 or nbytes % 16:
            raise ValueError("symmetric collective ranges must be 16-byte aligned")
        return byte_off, nbytes, _DTYPE_TAG[t.dtype]

    def all_reduce(self, buf, elem_offset: int = 0, numel: Optional[int] = None, scale: float = 1.0, algo: int = 0,
                   blocks: int = 0, sq_out: Optional[torch.Tensor] = None, tag: int = TAG_MISC) -> None:
        """In-place sum over ranks of ``buf.tensor[elem_offset : elem_offset + numel]`` (x ``scale``).  ``sq_out``
        (>= 64 fp32 slots): CTA b stores the sum of squares of what it reduced for this rank's 1/world slice."""
        byte_off, nbytes, dt = self._range(buf, elem_offset, numel)
        self.native.symm_allreduce(buf.ptrs, self.flags.ptrs, buf.multicast_ptr, self.rank, byte_off, nbytes, dt,
                                   float(scale), int(algo), int(blocks), 0 if sq_out is None else sq_out.data_ptr(),
                                   self.err_dev, int(tag))

    def reduce_scatter(self, buf, elem_offset: int = 0, numel: Optional[int] = None, scale: float = 1.0, blocks: int = 0,
                       sq_out: Optional[torch.Tensor] = None, tag: int = TAG_MISC) -> None:
        """The first half only: rank r's buffer ends up with the reduced vectors ``[b + r*per, b + (r+1)*per)``."""
        byte_off, nbytes, dt = self._range(buf, elem_offset, numel)
        self.native.symm_reduce_scatter(buf.ptrs, self.flags.ptrs, buf.multicast_ptr, self.rank, byte_off, nbytes, dt,
                                        float(scale), int(blocks), 0 if sq_out is None else sq_out.data_ptr(),
                                        self.err_dev, int(tag))

    def slice_of(self, lo: int, hi: int, elem_size: int, rank: Optional[int] = None):
        """Elements of ``[lo, hi)`` that rank ``rank`` owns after a reduce-scatter (the kernels' own formula)."""
        rank = self.rank if rank is None else rank
        epv = 16 // elem_size
        nvec = (hi - lo) // epv
        per = -(-nvec // self.world)
        return lo + min(nvec, per * rank) * epv, lo + min(nvec, per * (rank + 1)) * epv

    # -- statistics ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def stats_allreduce(self, values: torch.Tensor) -> torch.Tensor:
        """Sum a vector of <= 64 doubles over the ranks: one 64-thread kernel, one flag round trip, no NCCL.
        Returns a NEW fp64 device tensor (identical on every rank: rows are added in rank order)."""
        k = values.numel()
        if k > self.max_stats:
            raise ValueError("at most {} statistics per reduction".format(self.max_stats))
        src = values.detach().to(device=self.device, dtype=torch.float64).contiguous()
        dst = torch.empty(k, dtype=torch.float64, device=self.device)
        self._stats_calls += 1
        self.native.symm_stats_allreduce(self.xchg_stats.ptrs, self.flags.ptrs, self.err_dev, self.rank, TAG_STATS, src,
                                         dst, self._stats_calls & 1)
        return dst
