"""Symmetric-memory data parallelism for one NVSwitch domain (``--ddp-backend b200``).

Instead of handing gradients to NCCL (reference: torch DDP reducer,
``unicore/models/distributed_unicore_model.py:37-46``), the flat 16-bit gradient AND parameter arenas that the
mixed-precision optimizer builds (``unicore/optim/fp16_optimizer.py``) are *allocated in symmetric memory*
(``symm_mem.py``): every rank maps every peer's arenas (and an NVLS multicast alias) into its address space.  Autograd
writes gradients directly where the hand-written kernels read them - no bucket copies, no NCCL call on the step path.

During backward: parameters are grouped into contiguous buckets of the arena, ordered the way gradients become
ready; per-parameter ``post_accumulate_grad`` hooks count a bucket down and buckets are launched STRICTLY IN INDEX ORDER
(bucket i only after bucket i-1, like torch DDP) on a high-priority side stream, so every rank issues the same sequence
of collectives; the flag protocol carries the bucket index as a tag and fails loudly on a mismatch.

Default (``FusedTail``): a bucket launch is only the reduce-scatter half (``multimem.ld_reduce`` through the switch or
peer loads; the rank keeps the reduced 1/N slice and the partial sum of its squares).  After backward ONE kernel
(``csrc/comm/fused_step.cu``) reduce-scatters what is left, exchanges squared norm + logging statistics over peer
memory, derives multiplier / clip / overflow on the device, runs Adam (+EMA) on the rank's shard of compact fp32 state
and stores the new 16-bit parameters into every rank's parameter arena.  Fallback (optimizers the tail does not cover):
full all-reduce kernels per bucket + the replicated optimizer step.

``no_sync()`` disables communication for gradient-accumulation micro-batches.  NCCL remains for bootstrap and cold
paths (parameter broadcast at construction, checkpoint consolidation).
"""
import contextlib
import logging
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

from .comm import TAG_BUCKET, TAG_MISC, SymmComm
from .fused_tail import FusedTail, plan_buckets
from .symm_mem import SymmBuffer  # noqa: F401  (re-exported)

logger = logging.getLogger(__name__)


def symm_available() -> bool:
    """CUDA + initialised NCCL group + native kernels loaded + 2..8 ranks."""
    try:
        from unicore_b200.ops import _native

        if not (_native.USE_NATIVE and hasattr(_native.native(), "symm_fused_tail")):
            return False
        if not (torch.cuda.is_available() and dist.is_available() and dist.is_initialized()):
            return False
        return 2 <= dist.get_world_size() <= _native.native().SYMM_MAX_PEERS
    except Exception as exc:  # noqa: BLE001
        logger.warning("symmetric-memory data parallelism unavailable: %r", exc)
        return False


class SymmAllReduce:
    """Callable front of ``SymmComm`` for benchmarks and tests: in-place (all-)reduce of a range of a symmetric buffer."""

    def __init__(self, group=None, comm: Optional[SymmComm] = None):
        self.comm = comm if comm is not None else SymmComm(group)
        self.native, self.group = self.comm.native, self.comm.group
        self.rank, self.world, self.flags = self.comm.rank, self.comm.world, self.comm.flags

    def allocate(self, numel: int, dtype: torch.dtype):
        return self.comm.allocate(numel, dtype)

    def sq_slots(self) -> torch.Tensor:
        return torch.zeros(self.comm.max_blocks, dtype=torch.float32, device=self.comm.device)

    def __call__(self, buf, elem_offset: int = 0, numel: Optional[int] = None, scale: float = 1.0, algo: int = 0,
                 blocks: int = 0, sq_out: Optional[torch.Tensor] = None, scatter_only: bool = False, tag: int = TAG_MISC):
        if scatter_only:
            self.comm.reduce_scatter(buf, elem_offset, numel, scale=scale, blocks=blocks, sq_out=sq_out, tag=tag)
        else:
            self.comm.all_reduce(buf, elem_offset, numel, scale=scale, algo=algo, blocks=blocks, sq_out=sq_out, tag=tag)


class _Bucket:
    __slots__ = ("buffer", "group", "lo", "hi", "index", "pending", "total", "launched")

    def __init__(self, buffer, group, lo, hi, index):
        self.buffer, self.group, self.lo, self.hi, self.index = buffer, group, lo, hi, index
        self.pending = self.total = 0
        self.launched = False


class SymmDataParallel(nn.Module):
    """Data-parallel wrapper whose gradient reduction and optimizer tail run on hand-written NVLink kernels."""

    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: int = 25):
        super().__init__()
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.comm = SymmComm(self.process_group)
        self.reducer = SymmAllReduce(comm=self.comm)
        self.bucket_bytes = max(1, int(bucket_cap_mb)) * 1024 * 1024
        self.accumulate_grads = False
        self.want_fused_tail = os.environ.get("UNICORE_B200_FUSED_TAIL", "1") != "0"
        self._grad_buffers: List = []
        self._param_buffers: List = []
        self._buckets: List[_Bucket] = []
        self._param_buckets: Dict[nn.Parameter, List[_Bucket]] = {}
        self._hooks = []
        self._next = 0                      # index of the next bucket to launch (strict order)
        self._seen = set()
        self._comm_stream = torch.cuda.Stream(priority=-1)
        self._started = False
        self._covers_all_params = False
        self.tail: Optional[FusedTail] = None
        self._sq_slots: Optional[torch.Tensor] = None   # fallback path: [n_buckets, 64] partial sums of squares
        self._sq_total: Optional[torch.Tensor] = None
        # replicas must start identical (reference: DDP broadcasts from rank 0 at construction)
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=self.process_group)

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        previous, self.accumulate_grads = self.accumulate_grads, True
        try:
            yield
        finally:
            self.accumulate_grads = previous

    # -- optimizer integration ----------------------------------------------------------------------------
    def _reuse(self, pool: List, numel: int, dtype: torch.dtype):
        """An arena of a previous optimizer build (checkpoint load, reinitialize) is reused instead of leaked."""
        for buf in pool:
            if not getattr(buf, "in_use", False) and buf.tensor.numel() == numel and buf.tensor.dtype == dtype:
                return buf
        return None

    def begin_optimizer_build(self) -> None:
        for buf in self._grad_buffers + self._param_buffers:
            buf.in_use = False

    def _alloc(self, pool: List, numel: int, dtype: torch.dtype) -> torch.Tensor:
        padded = -(-numel // 8) * 8  # whole 16-byte vectors
        buf = self._reuse(pool, padded, dtype)
        if buf is None:
            buf = self.comm.allocate(padded, dtype)
            pool.append(buf)
        buf.in_use = True
        buf.tensor.zero_()
        return buf.tensor[:numel]

    def alloc_grad_buffer(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        """Called by ``flatten_parameters``: the flat gradient arena lives in symmetric memory."""
        return self._alloc(self._grad_buffers, numel, dtype)

    def alloc_param_buffer(self, numel: int, dtype: torch.dtype, device: torch.device) -> Optional[torch.Tensor]:
        """Called by ``flatten_parameters``: with the fused tail the flat PARAMETER arena is symmetric too (every rank
        stores its shard of the new parameters into all of them); otherwise decline."""
        if not self.want_fused_tail or dtype not in (torch.float16, torch.bfloat16):
            return None
        return self._alloc(self._param_buffers, numel, dtype)

    def _buffer_of(self, pool, tensor):
        for buf in pool:
            if getattr(buf, "in_use", False) and buf.tensor.data_ptr() == tensor.data_ptr():
                return buf
        return None

    def attach_optimizer(self, optimizer, params=None) -> None:
        """Build the bucket plan over the flat arenas, install gradient-ready hooks, and hand the optimizer the fused
        tail when it can use it.  ``params``: every trainable parameter of the optimizer (model AND loss; default: the
        wrapped module's)."""
        params = list(params) if params is not None else [p for p in self.module.parameters() if p.requires_grad]
        for h in self._hooks:
            h.remove()
        self._hooks, self._buckets, self._param_buckets, self.tail = [], [], {}, None
        flats = [f for g in getattr(optimizer, "fp16_params", []) for f in g["params"]]
        gbufs = [self._buffer_of(self._grad_buffers, f.grad) for f in flats]
        pbufs = [self._buffer_of(self._param_buffers, f.data) for f in flats]
        covered = len(flats) > 0 and all(b is not None for b in gbufs)
        if not covered:
            self._covers_all_params = False
            return
        tail_ok = (
            self.want_fused_tail and all(b is not None for b in pbufs) and hasattr(optimizer, "enable_fused_tail")
            and all(len(g["params"]) == 1 for g in optimizer.fp16_params)
            and len({f.dtype for f in flats}) == 1 and flats[0].dtype in (torch.float16, torch.bfloat16)
        )
        if tail_ok:
            tail = FusedTail(self.comm, gbufs, pbufs, self.bucket_bytes, seed=getattr(optimizer.args, "seed", 0))
            if optimizer.enable_fused_tail(self, tail):
                self.tail = tail
        if self.tail is not None:
            self._buckets = [_Bucket(gbufs[b.group], b.group, b.lo, b.hi, b.index) for b in self.tail.buckets]
        else:
            esz = flats[0].element_size()
            plan = plan_buckets([b.tensor.numel() for b in gbufs], esz, self.bucket_bytes, 4096)
            self._buckets = [_Bucket(gbufs[g], g, lo, hi, i) for i, (g, lo, hi) in enumerate(plan)]
            self._sq_slots = torch.zeros(len(self._buckets) * self.comm.max_blocks, dtype=torch.float32, device=self.comm.device)
        # parameter -> the buckets its gradient overlaps
        for gi, flat in enumerate(flats):
            base, esz = flat.grad.data_ptr(), flat.grad.element_size()
            mine = sorted((b for b in self._buckets if b.group == gi), key=lambda b: b.lo)
            for p in params:
                if p.grad is None or not p.requires_grad or p.grad.dtype != flat.grad.dtype:
                    continue
                off = (p.grad.data_ptr() - base) // esz
                if not (0 <= off < flat.grad.numel()):
                    continue
                end = off + p.grad.numel()
                touched = [b for b in mine if b.lo < end and off < b.hi]
                for b in touched:
                    b.total += 1
                self._param_buckets[p] = touched
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
        # kernels that write parameter gradients straight into the arena (ops/grad_sink.py) bypass AccumulateGrad and its
        # hooks: they report through this callback instead
        from unicore_b200.ops import grad_sink

        self._sink_params = [p for p in params if p in self._param_buckets]
        grad_sink.enable(self._sink_params, on_written=self._on_grad_ready)
        self._reset_counters()
        self._covers_all_params = all(p in self._param_buckets for p in params if p.requires_grad)
        if self.tail is None and self._covers_all_params and hasattr(optimizer, "set_external_grad_sq_norm"):
            optimizer.set_external_grad_sq_norm(self.grad_sq_norm)
        if self.tail is not None and not self._covers_all_params:
            raise RuntimeError("fused tail: a trainable parameter lies outside the flat arenas")
        logger.info("b200 data parallel: %d buckets, %s, symmetric memory by %s%s", len(self._buckets),
                    "fused optimizer tail" if self.tail is not None else "all-reduce + replicated optimizer",
                    self.comm.provider, ", NVLS" if gbufs[0].multicast_ptr else "")

    # -- bucket scheduling ---------------------------------------------------------------------------------------
    def grad_sq_norm(self):
        """Fallback path: squared L2 norm of the (averaged) gradients of the step just reduced, or None."""
        return self._sq_total

    def _reset_counters(self):
        for b in self._buckets:
            b.pending = b.total
            b.launched = False
        self._next = 0
        self._started = False
        self._seen = set()
        for p in getattr(self, "_sink_params", ()):
            p._ub_pending = 0

    def _on_grad_ready(self, param):
        if self.accumulate_grads or id(param) in self._seen:
            return  # (a hook that fires twice for one parameter in one update must not count its buckets down twice)
        self._seen.add(id(param))
        for b in self._param_buckets.get(param, ()):
            if b.pending > 0:
                b.pending -= 1
        self._launch_ready_prefix()

    def _launch_ready_prefix(self, keep_last: bool = True):
        """Launch buckets next, next+1, ... while they are complete.  With the fused tail the LAST bucket is left to the
        tail kernel (it becomes ready when backward ends; the tail follows immediately and saves two launches)."""
        limit = len(self._buckets) - (1 if (self.tail is not None and keep_last) else 0)
        while self._next < limit and self._buckets[self._next].pending == 0:
            self._launch(self._buckets[self._next])
            self._next += 1

    def _launch(self, b: _Bucket):
        # the bucket's gradients were produced on the current (compute) stream
        self._comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            self._started = True
            if self.tail is not None:
                self.comm.reduce_scatter(b.buffer, b.lo, b.hi - b.lo, scale=1.0 / self.world_size,
                                         sq_out=self.tail.sq_slots(b.index), tag=TAG_BUCKET + b.index)
            else:
                n = self.comm.max_blocks
                self.comm.all_reduce(b.buffer, b.lo, b.hi - b.lo, scale=1.0 / self.world_size,
                                     sq_out=self._sq_slots[b.index * n:(b.index + 1) * n], tag=TAG_BUCKET + b.index)
        b.launched = True

    def flush_after_failure(self) -> None:
        """A rank whose backward died (out of memory) still has to issue the collectives its peers are waiting in:
        every bucket the hooks would have launched, in order (the gradients behind them were zeroed by the caller)."""
        if self.accumulate_grads or not self._buckets:
            return
        for b in self._buckets:
            b.pending = 0
        self._launch_ready_prefix()

    def pending_buckets(self) -> List[int]:
        return [b.index for b in self._buckets if not b.launched]

    def all_reduce_grads(self):
        """Fallback path: flush unreduced buckets (in order) and join the comm stream.  Fused tail: nothing to do here -
        whatever has not been launched is reduce-scattered by the tail kernel itself (``run_tail``)."""
        if self.accumulate_grads:
            return
        if not self._buckets:  # optimizer without flat arenas: NCCL per tensor
            for p in self.module.parameters():
                if p.grad is not None:
                    p.grad.div_(self.world_size)
                    dist.all_reduce(p.grad, group=self.process_group)
            return
        if self.tail is not None:
            return
        for b in self._buckets[self._next:]:
            self._launch(b)
        self._next = len(self._buckets)
        with torch.cuda.stream(self._comm_stream):
            # every rank holds the squares of its slices in per-CTA slots: fixed-order local sum, then one 64-thread
            # exchange kernel gives all ranks the same total
            local = self._sq_slots.double().sum().reshape(1)
            self._sq_total = self.comm.stats_allreduce(local)[0].float() if self._covers_all_params else None
            self._sq_slots.zero_()
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._reset_counters()

    def run_tail(self, **kwargs) -> torch.Tensor:
        """Launch the fused optimizer tail behind this update's bucket kernels (called by the optimizer's ``step``)."""
        pending = self.pending_buckets()
        self._comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            state = self.tail.launch(pending=pending, **kwargs)
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._reset_counters()
        return state

    # -- statistics ------------------------------------------------------------------------------------------
    def reduce_stats(self, values: torch.Tensor) -> torch.Tensor:
        """Sum of a small fp64 vector over the ranks on the peer-memory kernel (replaces the per-step NCCL all-reduce
        of the reference's ``_fast_stat_sync_sum``, ``unicore/trainer.py:1011-1049``)."""
        return self.comm.stats_allreduce(values)

    def check_health(self) -> None:
        self.comm.check_health()
