"""Symmetric-memory data parallelism for one NVSwitch domain (``--ddp-backend b200``).

Instead of handing gradients to NCCL (reference: torch DDP reducer,
``unicore/models/distributed_unicore_model.py:37-46``), the flat 16-bit gradient arena that the
mixed-precision optimizer builds (``unicore/optim/fp16_optimizer.py``) is *allocated in symmetric
memory*: every rank maps every peer's arena (and an NVLS multicast alias) into its address space.
Autograd therefore writes gradients directly where the hand-written reduction kernels
(``csrc/comm/allreduce.cu``: one-shot / two-shot peer loads+stores, or ``multimem.ld_reduce`` /
``multimem.st`` through the switch) read them - no bucket copies, no NCCL call on the gradient path.

Overlap with backward: parameters are grouped into contiguous buckets of the arena (reverse
registration order ~ gradient-ready order); per-parameter ``post_accumulate_grad`` hooks count a
bucket down and, when it is complete, launch its reduction on a high-priority side stream.
``all_reduce_grads()`` (called by the trainer after backward) flushes whatever is left and joins
the streams.  ``no_sync()`` disables communication for gradient-accumulation micro-batches.

Rendezvous uses ``torch.distributed._symmetric_memory`` (CUDA VMM handles exchanged over the
process group's store); NCCL remains only for bootstrap and cold paths (parameter broadcast,
checkpoints, tiny statistics).
"""
import contextlib
import logging
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
from torch import nn

logger = logging.getLogger(__name__)


def _symm_module():
    import torch.distributed._symmetric_memory as symm_mem

    return symm_mem


def symm_available() -> bool:
    """CUDA + initialised NCCL group + symmetric memory importable + native kernels loaded."""
    try:
        from unicore_b200.ops import _native

        if not (_native.USE_NATIVE and hasattr(_native.native(), "symm_allreduce")):
            return False
        if not (torch.cuda.is_available() and dist.is_available() and dist.is_initialized()):
            return False
        if dist.get_world_size() < 2 or dist.get_world_size() > _native.native().SYMM_MAX_PEERS:
            return False
        _symm_module()
        return True
    except Exception as exc:  # noqa: BLE001
        logger.warning("symmetric-memory data parallelism unavailable: %r", exc)
        return False


class SymmBuffer:
    """One symmetric allocation: local tensor + peer addresses (+ multicast alias)."""

    def __init__(self, numel: int, dtype: torch.dtype, device: torch.device, group):
        symm_mem = _symm_module()
        self.tensor = symm_mem.empty(numel, dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.tensor, group)
        self.rank = self.handle.rank
        self.world = self.handle.world_size
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = 0
        try:
            if getattr(self.handle, "has_multicast_support", False):
                mc = int(self.handle.multicast_ptr)
        except Exception:  # noqa: BLE001
            mc = 0
        self.multicast_ptr = mc


class SymmAllReduce:
    """Launches the peer-memory all-reduce kernels on ranges of a symmetric buffer."""

    def __init__(self, group=None):
        from unicore_b200.ops._native import native

        self.native = native()
        self.group = group if group is not None else dist.group.WORLD
        self.device = torch.device("cuda", torch.cuda.current_device())
        n_flags = int(self.native.SYMM_MAX_BLOCKS) * int(self.native.SYMM_MAX_PEERS)
        self.flags = SymmBuffer(n_flags, torch.int32, self.device, self.group)
        self.flags.tensor.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # flags are zero everywhere before the first kernel spins on them
        self.rank = self.flags.rank
        self.world = self.flags.world

    def allocate(self, numel: int, dtype: torch.dtype) -> SymmBuffer:
        return SymmBuffer(numel, dtype, self.device, self.group)

    def __call__(self, buf: SymmBuffer, elem_offset: int = 0, numel: Optional[int] = None, scale: float = 1.0,
                 algo: int = 0, blocks: int = 0, sq_acc: Optional[torch.Tensor] = None, scatter_only: bool = False):
        """In-place sum over ranks of ``buf.tensor[elem_offset : elem_offset + numel]`` (x ``scale``).

        ``sq_acc`` (fp32 device scalar): this rank adds the sum of squares of the reduced values of its
        1/world slice; summed over ranks that is the squared L2 norm of the result."""
        t = buf.tensor
        numel = t.numel() - elem_offset if numel is None else numel
        esz = t.element_size()
        byte_off, nbytes = elem_offset * esz, numel * esz
        if byte_off % 16 or nbytes % 16:
            raise ValueError("symmetric all-reduce ranges must be 16-byte aligned")
        tag = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[t.dtype]
        if scatter_only:  # reduce-scatter half only: rank r keeps the reduced slice r of the range (sharded optimizer)
            self.native.symm_reduce_scatter(
                buf.ptrs, self.flags.ptrs, buf.multicast_ptr, self.rank, byte_off, nbytes, tag, float(scale), int(blocks),
                0 if sq_acc is None else sq_acc.data_ptr(),
            )
            return
        self.native.symm_allreduce(
            buf.ptrs, self.flags.ptrs, buf.multicast_ptr, self.rank, byte_off, nbytes, tag, float(scale), int(algo),
            int(blocks), 0 if sq_acc is None else sq_acc.data_ptr(),
        )


class ShardedAdamStepper:
    """Optimizer step fused with the parameter all-gather (``csrc/comm/allreduce.cu::sharded_adam_kernel``).

    EXPERIMENTAL - enabled with ``UNICORE_B200_SHARD_OPTIMIZER=1``; written in round 1 without access to a
    multi-GPU box for validation.  Rank ``r`` owns elements ``[r*per, (r+1)*per)`` of every
    flat parameter group (``per`` a multiple of 8): one kernel runs Adam on that shard of the fp32 master /
    moments and stores the new 16-bit parameters into every rank's (symmetric) parameter arena with
    ``multimem.st`` (NVLS) or peer stores, closing with a flag barrier.
    """

    def __init__(self, reducer: SymmAllReduce, param_buffers: List[SymmBuffer], seed: int = 0):
        self.reducer = reducer
        self.rank, self.world = reducer.rank, reducer.world
        self.group = reducer.group
        self._by_ptr = {buf.tensor.data_ptr(): buf for buf in param_buffers}
        self.seed = int(seed)
        self._calls = 0
        # gradient-arena address -> [(lo, hi)] element ranges this rank owns (set by the engine when its buckets
        # run reduce-scatter-only); without it the shard is one contiguous 1/world slice of the group
        self.bucket_slices: Dict[int, List] = {}

    def bounds(self, numel: int):
        per = -(-(-(-numel // 8)) // self.world) * 8
        lo = min(numel, self.rank * per)
        return per, lo, min(numel, lo + per)

    def ranges(self, flat: torch.Tensor, numel: int):
        """Element ranges of the flat group (length ``numel``) whose update this rank performs."""
        slices = self.bucket_slices.get(flat.grad.data_ptr())
        if slices is None:
            _, lo, hi = self.bounds(numel)
            return [(lo, hi)] if hi > lo else []
        out = []
        for lo, hi in slices:
            lo, hi = min(lo, numel), min(hi, numel)
            if hi > lo:
                out.append((lo, hi))
        return out

    def covers(self, flat: torch.Tensor) -> bool:
        return flat.data_ptr() in self._by_ptr

    def step(self, flat, master, exp_avg, exp_avg_sq, *, lr, beta1, beta2, eps, step, bias_correction, weight_decay,
             grad_scale, stochastic_rounding=False):
        buf = self._by_ptr[flat.data_ptr()]
        ranges = self.ranges(flat, master.numel())
        scale_f, scale_dev = 1.0, None
        if torch.is_tensor(grad_scale):
            scale_dev = grad_scale.detach().float().reshape(1)
        else:
            scale_f = float(grad_scale)
        self._calls += 1
        self.reducer.native.symm_sharded_adam(
            buf.ptrs, self.reducer.flags.ptrs, buf.multicast_ptr, self.rank, flat.grad, master, exp_avg, exp_avg_sq,
            [r[0] for r in ranges], [r[1] for r in ranges], float(lr), float(beta1), float(beta2), float(eps), int(step),
            bool(bias_correction), float(weight_decay), scale_f, scale_dev, bool(stochastic_rounding), self.seed,
            self._calls, 0,
        )

    @torch.no_grad()
    def gather_(self, t: torch.Tensor, flat: torch.Tensor) -> None:
        """Refresh the full-length fp32 tensor ``t`` of ``flat``'s group from the ranges every rank keeps current
        (one NCCL all-reduce of a zero-filled copy; cold path, checkpoints only)."""
        mine = torch.zeros_like(t)
        for lo, hi in self.ranges(flat, t.numel()):
            mine[lo:hi].copy_(t[lo:hi])
        dist.all_reduce(mine, group=self.group)
        t.copy_(mine)


class _Bucket:
    __slots__ = ("buffer", "lo", "hi", "pending", "total", "launched")

    def __init__(self, buffer, lo, hi):
        self.buffer, self.lo, self.hi = buffer, lo, hi
        self.pending = self.total = 0
        self.launched = False


class SymmDataParallel(nn.Module):
    """Data-parallel wrapper whose gradient reduction runs on hand-written NVLink kernels."""

    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: int = 25):
        super().__init__()
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.reducer = SymmAllReduce(self.process_group)
        self.bucket_bytes = max(1, int(bucket_cap_mb)) * 1024 * 1024
        self.accumulate_grads = False
        self._buffers: List[SymmBuffer] = []
        self._buckets: List[_Bucket] = []
        self._param_bucket: Dict[nn.Parameter, _Bucket] = {}
        self._hooks = []
        self._comm_stream = torch.cuda.Stream(priority=-1)
        self._started = False
        # squared gradient norm, accumulated by the reduction kernels themselves (4 floats = one 16-byte vector)
        self._sq = self.reducer.allocate(4, torch.float32)
        self._sq.tensor.zero_()
        self._sq_valid = False
        self._covers_all_params = False
        # experimental: Adam on a 1/N shard + parameter all-gather in one kernel (see ShardedAdamStepper)
        #   1: contiguous 1/N shard on top of the full bucket all-reduce
        #   2: buckets stop after their reduce-scatter half; rank r updates its slice of every bucket
        self._shard_mode = int(os.environ.get("UNICORE_B200_SHARD_OPTIMIZER", "0") or 0)
        self.shard_optimizer = self._shard_mode in (1, 2)
        self._scatter_buckets = False
        self._stepper: Optional[ShardedAdamStepper] = None
        self._param_buffers: List[SymmBuffer] = []
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
    def alloc_grad_buffer(self, numel: int, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
        """Called by ``flatten_parameters``: the flat gradient arena lives in symmetric memory."""
        padded = -(-numel // 8) * 8  # whole 16-byte vectors
        buf = self.reducer.allocate(padded, dtype)
        buf.tensor.zero_()
        self._buffers.append(buf)
        return buf.tensor[:numel]

    def alloc_param_buffer(self, numel: int, dtype: torch.dtype, device: torch.device) -> Optional[torch.Tensor]:
        """Called by ``flatten_parameters``: with the sharded optimizer the flat PARAMETER arena is symmetric too
        (every rank stores its shard of the new parameters into all of them); otherwise decline."""
        if not self.shard_optimizer:
            return None
        padded = -(-numel // 8) * 8
        buf = self.reducer.allocate(padded, dtype)
        self._param_buffers.append(buf)
        return buf.tensor[:numel]

    def _maybe_enable_sharded_step(self, optimizer) -> None:
        if not (self.shard_optimizer and self._param_buffers and hasattr(optimizer, "enable_sharded_step")):
            return
        if getattr(getattr(optimizer, "args", None), "ema_decay", -1) > 0:
            logger.warning("UNICORE_B200_SHARD_OPTIMIZER ignored: the EMA update reads the full fp32 master weights")
            return
        stepper = ShardedAdamStepper(self.reducer, self._param_buffers, seed=getattr(optimizer.args, "seed", 0))
        flats = [f for g in optimizer.fp16_params for f in g["params"]]
        if all(stepper.covers(f) for f in flats) and optimizer.enable_sharded_step(stepper):
            self._stepper = stepper
            logger.info("optimizer step sharded over %d ranks and fused with the parameter all-gather", self.world_size)

    def _plan_bucket_slices(self) -> None:
        """Mode 2: every bucket stops after reduce-scatter; tell the stepper which slice of each bucket is ours
        (same formula as the kernels: 16-byte vectors [begin + r*per, begin + (r+1)*per), per = ceil(n / world))."""
        self._scatter_buckets = False
        if self._stepper is None or self._shard_mode != 2 or not self._covers_all_params:
            return
        slices: Dict[int, List] = {}
        for b in self._buckets:
            epv = 16 // b.buffer.tensor.element_size()
            nvec = (b.hi - b.lo) // epv
            per = -(-nvec // self.world_size)
            lo = b.lo + min(nvec, per * self.reducer.rank) * epv
            hi = b.lo + min(nvec, per * (self.reducer.rank + 1)) * epv
            slices.setdefault(b.buffer.tensor.data_ptr(), []).append((lo, hi))
        max_ranges = int(getattr(self.reducer.native, "SYMM_MAX_SHARD_RANGES", 48))
        if any(len(v) > max_ranges for v in slices.values()):
            logger.warning("too many buckets for slice-wise sharding; keeping the full all-reduce (raise --bucket-cap-mb)")
            return
        self._stepper.bucket_slices = slices
        self._scatter_buckets = True

    def attach_optimizer(self, optimizer) -> None:
        """Build buckets over the flat gradient arenas and install gradient-ready hooks."""
        self._maybe_enable_sharded_step(optimizer)
        for h in self._hooks:
            h.remove()
        self._hooks, self._buckets, self._param_bucket = [], [], {}
        by_ptr = {buf.tensor.data_ptr(): buf for buf in self._buffers}
        for group in optimizer.fp16_params:
            for flat in group["params"]:
                buf = by_ptr.get(flat.grad.data_ptr())
                if buf is None:
                    continue
                esz = flat.grad.element_size()
                per_bucket = max(8, (self.bucket_bytes // esz) // 8 * 8)
                total = buf.tensor.numel()
                edges = list(range(0, total, per_bucket)) + [total]
                buckets = [_Bucket(buf, lo, hi) for lo, hi in zip(edges[:-1], edges[1:])]
                self._buckets.extend(buckets)
                base = flat.grad.data_ptr()
                for p in self.module.parameters():
                    if p.grad is None or not p.requires_grad:
                        continue
                    off = (p.grad.data_ptr() - base) // esz
                    if not (0 <= off < flat.grad.numel()) or p.grad.dtype != flat.grad.dtype:
                        continue
                    # every bucket the parameter overlaps has to wait for its gradient
                    last = min(off + p.grad.numel() - 1, total - 1)
                    first_b = off // per_bucket
                    last_b = min(last // per_bucket, len(buckets) - 1)
                    touched = buckets[first_b:last_b + 1]
                    for b in touched:
                        b.total += 1
                    self._param_bucket[p] = touched
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
        self._reset_counters()
        trainable = [p for p in self.module.parameters() if p.requires_grad]
        self._covers_all_params = len(self._buckets) > 0 and all(p in self._param_bucket for p in trainable)
        if self._covers_all_params and hasattr(optimizer, "set_external_grad_sq_norm"):
            optimizer.set_external_grad_sq_norm(self.grad_sq_norm)
        self._plan_bucket_slices()

    def grad_sq_norm(self):
        """Squared L2 norm of the (averaged) gradients of the step just reduced, or None if unavailable."""
        return self._sq.tensor[0] if self._sq_valid else None

    def _reset_counters(self):
        for b in self._buckets:
            b.pending = b.total
            b.launched = False
        self._started = False

    def _on_grad_ready(self, param):
        if self.accumulate_grads:
            return
        for b in self._param_bucket.get(param, ()):
            b.pending -= 1
            if b.pending == 0 and not b.launched:
                self._launch(b)

    def _launch(self, b: _Bucket):
        # the bucket's gradients were produced on the current (compute) stream
        self._comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            if not self._started:  # first bucket of this update: restart the norm accumulator
                self._started = True
                self._sq_valid = False
                self._sq.tensor.zero_()
            self.reducer(b.buffer, b.lo, b.hi - b.lo, scale=1.0 / self.world_size, sq_acc=self._sq.tensor,
                         scatter_only=self._scatter_buckets)
        b.launched = True

    def all_reduce_grads(self):
        """Flush unreduced buckets (unused params, no hooks fired) and join the comm stream."""
        if self.accumulate_grads:
            return
        if not self._buckets:  # optimizer without flat arenas: fall back to NCCL per tensor
            for p in self.module.parameters():
                if p.grad is not None:
                    p.grad.div_(self.world_size)
                    dist.all_reduce(p.grad, group=self.process_group)
            return
        for b in self._buckets:
            if not b.launched:
                self._launch(b)
        with torch.cuda.stream(self._comm_stream):
            # every rank holds the squares of its slices: one 16-byte one-shot reduction gives all of them the total
            self.reducer(self._sq, 0, 4, scale=1.0, algo=1)
        self._sq_valid = self._covers_all_params
        torch.cuda.current_stream().wait_stream(self._comm_stream)
        self._reset_counters()
