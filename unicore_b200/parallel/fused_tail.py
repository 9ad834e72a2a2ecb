"""Plan + launcher of the fused optimizer tail (``csrc/comm/fused_step.cu``).

The gradient arenas of the mixed-precision optimizer (one flat 16-bit buffer per weight-decay group,
``unicore/optim/fp16_optimizer.py``) are cut into buckets; after a bucket's reduce-scatter rank ``r`` holds the reduced
values of ITS slice of the bucket (16-byte vectors ``[b + r*per, b + (r+1)*per)``, ``per = ceil(n / world)`` - the
formula of the kernels).  The union of a rank's slices is its SHARD of the group: the fp32 master weights and the Adam
moments exist only for the shard, stored compactly (slices back to back, in bucket order).  ``FusedTail`` owns that
geometry, the device-side range table and the scratch of the kernel, and issues the ONE launch that follows backward.

Replaces the reference tail ``clip_grad_norm -> step(scale) -> _sync_fp32_params_to_fp16 -> zero_grad``
(``unicore/optim/fp16_optimizer.py:258-308``), the DDP all-reduce of the last bucket
(``unicore/models/distributed_unicore_model.py:37-46``), the logging all-reduce and the grad-norm all-gather
(``unicore/trainer.py:1011-1084``) and the EMA pass (``unicore/ema.py:44-60``).
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .comm import TAG_TAIL, SymmComm

_DTYPE_TAG = {torch.float16: 1, torch.bfloat16: 2}


class TailBucket:
    __slots__ = ("group", "lo", "hi", "own_lo", "own_hi", "compact_off", "index")

    def __init__(self, group, lo, hi):
        self.group, self.lo, self.hi = group, lo, hi
        self.own_lo = self.own_hi = self.compact_off = self.index = 0


def plan_buckets(numels: Sequence[int], elem_size: int, bucket_bytes: int, max_buckets: int) -> List[Tuple[int, int, int]]:
    """Cut the group arenas (lengths in elements, multiples of 8) into buckets ``(group, lo, hi)`` ordered the way
    backward produces gradients: parameters were laid out in registration order, so the END of every arena becomes
    ready first.  The bucket size grows until the plan fits ``max_buckets``."""
    per = max(8, (bucket_bytes // elem_size) // 8 * 8)
    while True:
        plan = []
        for g, n in enumerate(numels):
            edges = list(range(0, n, per)) + [n]
            plan += [(g, lo, hi) for lo, hi in zip(edges[:-1], edges[1:]) if hi > lo]
        if len(plan) <= max_buckets:
            break
        per *= 2
    # descending relative position; ties (several groups) broken by group index for a rank-independent order
    plan.sort(key=lambda b: (-(b[2] / float(numels[b[0]])), b[0]))
    return plan


class FusedTail:
    def __init__(self, comm: SymmComm, grad_buffers, param_buffers, bucket_bytes: int, seed: int = 0):
        """``grad_buffers[g]`` / ``param_buffers[g]``: the symmetric arenas (``SymmBuffer``) of flat group ``g``."""
        assert len(grad_buffers) == len(param_buffers) >= 1
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        self.grad_buffers, self.param_buffers = list(grad_buffers), list(param_buffers)
        native = comm.native
        if len(self.grad_buffers) > int(native.SYMM_MAX_TAIL_GROUPS):
            raise ValueError("too many flat parameter groups for the fused tail")
        self.dtype = self.grad_buffers[0].tensor.dtype
        if any(b.tensor.dtype != self.dtype for b in self.grad_buffers + self.param_buffers) or self.dtype not in _DTYPE_TAG:
            raise ValueError("the fused tail needs fp16 or bf16 arenas of one dtype")
        self.esz = self.grad_buffers[0].tensor.element_size()
        self.numels = [b.tensor.numel() for b in self.grad_buffers]
        if any(n % 8 for n in self.numels) or any(p.tensor.numel() != n for p, n in zip(self.param_buffers, self.numels)):
            raise ValueError("arenas must be whole 16-byte vectors and pairwise equally long")
        plan = plan_buckets(self.numels, self.esz, bucket_bytes, int(native.SYMM_MAX_TAIL_RANGES))
        self.buckets: List[TailBucket] = []
        compact = [0] * len(self.numels)
        for index, (g, lo, hi) in enumerate(plan):
            b = TailBucket(g, lo, hi)
            b.index = index
            b.own_lo, b.own_hi = comm.slice_of(lo, hi, self.esz)
            b.compact_off = compact[g]
            compact[g] += b.own_hi - b.own_lo
            self.buckets.append(b)
        self.compact_numels = compact
        table = [[b.lo, b.hi, b.own_lo, b.own_hi, b.compact_off, b.group] for b in self.buckets]
        dev = comm.device
        self.ranges = torch.tensor(table, dtype=torch.int64, device=dev)
        self.bucket_sq = torch.zeros(len(self.buckets) * comm.max_blocks, dtype=torch.float32, device=dev)
        self.blocks = int(native.symm_tail_max_blocks())
        self.block_sq = torch.zeros(self.blocks, dtype=torch.float32, device=dev)
        self.grid_sync = torch.zeros(4, dtype=torch.int32, device=dev)
        self.state = torch.zeros(4, dtype=torch.float32, device=dev)  # {grad_norm, multiplier, overflow, sum sq}
        self.stats_dst = torch.zeros(comm.max_stats, dtype=torch.float64, device=dev)
        self.seed = int(seed)
        self.calls = 0

    # -- geometry ---------------------------------------------------------------------------------------------------
    def owned_ranges(self, group: int, rank: Optional[int] = None) -> List[Tuple[int, int]]:
        """This rank's (or ``rank``'s) slices of ``group``, in compact order."""
        out = []
        for b in self.buckets:
            if b.group != group:
                continue
            lo, hi = (b.own_lo, b.own_hi) if rank is None or rank == self.rank else self.comm.slice_of(b.lo, b.hi, self.esz, rank)
            if hi > lo:
                out.append((lo, hi))
        return out

    def sq_slots(self, bucket_index: int) -> torch.Tensor:
        n = self.comm.max_blocks
        return self.bucket_sq[bucket_index * n:(bucket_index + 1) * n]

    # -- shard <-> full conversions (cold paths: build, checkpoints) ---------------------------------------------------
    @torch.no_grad()
    def to_compact(self, full: torch.Tensor, group: int) -> torch.Tensor:
        """The owned elements of a full-length per-group tensor (shorter than the padded arena is fine)."""
        n = self.numels[group]
        if full.numel() < n:
            full = torch.cat([full.reshape(-1), full.new_zeros(n - full.numel())])
        parts = [full[lo:hi] for lo, hi in self.owned_ranges(group)]
        out = torch.cat(parts) if parts else full.new_zeros(0)
        return out.contiguous().clone()

    @torch.no_grad()
    def to_full(self, compact: torch.Tensor, group: int, numel: Optional[int] = None) -> torch.Tensor:
        """COLLECTIVE: every rank's compact shard -> the full-length tensor on every rank (one NCCL all-reduce of a
        zero-filled buffer; checkpoints only)."""
        import torch.distributed as dist

        n = self.numels[group]
        full = torch.zeros(n, dtype=compact.dtype, device=compact.device)
        off = 0
        for lo, hi in self.owned_ranges(group):
            full[lo:hi].copy_(compact[off:off + hi - lo])
            off += hi - lo
        dist.all_reduce(full, group=self.comm.group)
        return full[:numel] if numel is not None else full

    @torch.no_grad()
    def scatter_owned_(self, full: torch.Tensor, group: int) -> None:
        """COLLECTIVE, in place: every rank keeps only ITS slices of the full-length tensor current; afterwards all
        ranks hold the union (EMA arena before a checkpoint / validation)."""
        import torch.distributed as dist

        keep = torch.zeros_like(full)
        for lo, hi in self.owned_ranges(group):
            hi = min(hi, full.numel())
            if hi > lo:
                keep[lo:hi].copy_(full[lo:hi])
        dist.all_reduce(keep, group=self.comm.group)
        full.copy_(keep)

    # -- the launch ---------------------------------------------------------------------------------------------------
    def launch(self, *, masters, exp_avgs, exp_avg_sqs, hypers, pending: Sequence[int], factor: float, max_norm: float,
               clip_eps: float, emas=None, ema_decay: float = 0.0, stats_src: Optional[torch.Tensor] = None,
               denom_index: int = -1, stochastic_rounding: bool = False) -> torch.Tensor:
        """``hypers[g] = (beta1, beta2, eps, step_size, decay_mul)``; ``pending``: indices of the buckets that were NOT
        reduce-scattered during backward.  Returns the device state vector {grad_norm, multiplier, overflow, sum sq}
        (valid once the kernel has run; the statistics sums land in ``self.stats_dst``)."""
        ng = len(self.numels)
        mask = 0
        for i in pending:
            mask |= 1 << int(i)
        self.calls += 1
        emas = list(emas) if emas is not None else [None] * ng
        self.comm.native.symm_fused_tail(
            self.comm.xchg_tail.ptrs, self.comm.flags.ptrs, self.comm.err_dev, self.rank, TAG_TAIL,
            [b.ptrs for b in self.grad_buffers], [b.multicast_ptr for b in self.grad_buffers],
            [b.ptrs for b in self.param_buffers], [b.multicast_ptr for b in self.param_buffers],
            list(masters), list(exp_avgs), list(exp_avg_sqs), emas, list(self.numels),
            [_DTYPE_TAG[self.dtype]] * ng, [list(map(float, h)) for h in hypers],
            self.ranges, mask, self.bucket_sq, self.block_sq, self.grid_sync,
            stats_src, self.stats_dst[:stats_src.numel()] if stats_src is not None else None,
            self.calls & 1, int(denom_index), float(factor), float(max_norm), float(clip_eps), 1.0 / self.world,
            float(ema_decay), bool(stochastic_rounding), self.seed, self.calls, self.state, self.blocks,
        )
        return self.state


def adam_hyper(lr: float, beta1: float, beta2: float, eps: float, step: int, bias_correction: bool, weight_decay: float):
    """(beta1, beta2, eps, step_size, decay_mul) exactly as ``csrc/optim/multi_tensor.cu`` derives them."""
    step_size = lr
    if bias_correction:
        step_size = lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    return (beta1, beta2, eps, step_size, 1.0 - step_size * weight_decay)
