"""The fused optimizer tail, spelled out in PyTorch + ``torch.distributed`` collectives.

This is the executable specification of ``csrc/comm/fused_step.cu``: same inputs (``FusedTail`` geometry, compact fp32
shards, hyper-parameters, statistics vector), same outputs (state vector, statistics sums, parameters on every rank,
zeroed gradients), every formula written the obvious way.  It serves

* the CPU / gloo test-suite (``ReferenceTailEngine`` stands in for ``SymmDataParallel`` so that the optimizer's and the
  trainer's tail plumbing - compact state, checkpoint gather / reshard, EMA on the shard, in-tail statistics, deferred
  overflow - is exercised without a GPU), and
* the multi-GPU parity tests, which run the kernel and this reference side by side on identical inputs.
"""
import contextlib
from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from .fused_tail import FusedTail


class _PlainBuffer:
    """Duck type of ``symm_mem.SymmBuffer`` over ordinary memory (no peer mappings)."""

    def __init__(self, tensor, rank, world):
        self.tensor, self.rank, self.world = tensor, rank, world
        self.ptrs, self.multicast_ptr, self.provider = [0] * world, 0, "plain"
        self.in_use = True


class _Constants:
    SYMM_MAX_TAIL_GROUPS, SYMM_MAX_TAIL_RANGES = 4, 64

    @staticmethod
    def symm_tail_max_blocks():
        return 4


class PlainComm:
    """What ``FusedTail`` needs from a communicator, without symmetric memory."""

    max_blocks, max_stats = 64, 64
    native = _Constants()
    provider = "plain"

    def __init__(self, group, device):
        self.group, self.device = group, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def slice_of(self, lo, hi, elem_size, rank=None):
        rank = self.rank if rank is None else rank
        epv = 16 // elem_size
        nvec = (hi - lo) // epv
        per = -(-nvec // self.world)
        return lo + min(nvec, per * rank) * epv, lo + min(nvec, per * (rank + 1)) * epv

    def check_health(self):
        pass


@torch.no_grad()
def reference_tail(tail: FusedTail, *, masters, exp_avgs, exp_avg_sqs, hypers, factor, max_norm, clip_eps, emas=None,
                   ema_decay=0.0, stats_src=None, denom_index=-1, stochastic_rounding=False, pending=None):
    """One update.  ``tail.grad_buffers[g].tensor`` hold the LOCAL (unreduced) gradients of every rank on entry."""
    group, world = tail.comm.group, tail.world
    ng = len(tail.numels)
    # reduce-scatter (+ 1/world), squares of the owned slices
    local_sq = torch.zeros((), dtype=torch.float64, device=tail.comm.device)
    reduced = []
    for g in range(ng):
        full = tail.grad_buffers[g].tensor.float()
        dist.all_reduce(full, group=group)
        full = (full * (1.0 / world)).to(tail.dtype).float()  # the kernels round the scaled sum to 16 bits
        reduced.append(full)
        for lo, hi in tail.owned_ranges(g):
            local_sq += full[lo:hi].double().pow(2).sum()
    # norm + statistics exchange
    k = 0 if stats_src is None else int(stats_src.numel())
    row = torch.zeros(1 + k, dtype=torch.float64, device=tail.comm.device)
    row[0] = local_sq
    if k:
        row[1:] = stats_src.to(torch.float64)
    dist.all_reduce(row, group=group)
    total_sq = row[0]
    denom = 1.0
    if 0 <= denom_index < k and float(row[1 + denom_index]) > 0:
        denom = float(row[1 + denom_index])
    gmul0 = float(factor) / denom
    norm = float(total_sq.float().sqrt()) * gmul0
    overflow = not (norm == norm and abs(norm) != float("inf"))
    coef = 1.0
    if max_norm > 0 and norm > max_norm:
        coef = max_norm / (norm + clip_eps)
    gmul = gmul0 * coef
    tail.state.copy_(torch.tensor([norm, gmul, 1.0 if overflow else 0.0, float(total_sq)], dtype=torch.float32))
    if k:
        tail.stats_dst[:k].copy_(row[1:])
    # Adam (+EMA) on the shard, parameters to everybody
    emas = list(emas) if emas is not None else [None] * ng
    for g in range(ng):
        if not overflow:
            beta1, beta2, eps, step_size, decay_mul = hypers[g]
            params = tail.param_buffers[g].tensor
            new_params = torch.zeros_like(params, dtype=torch.float32)
            off = 0
            for lo, hi in tail.owned_ranges(g):
                n = hi - lo
                grad = reduced[g][lo:hi] * gmul
                p, m, v = masters[g][off:off + n], exp_avgs[g][off:off + n], exp_avg_sqs[g][off:off + n]
                m.mul_(beta1).add_(grad, alpha=1.0 - beta1)
                v.mul_(beta2).addcmul_(grad, grad, value=1.0 - beta2)
                p.mul_(decay_mul).addcdiv_(m, v.sqrt() + eps, value=-step_size)
                if emas[g] is not None:
                    e = emas[g].data
                    top = min(hi, e.numel())
                    if top > lo:
                        e[lo:top].sub_(e[lo:top] - p[:top - lo], alpha=1.0 - ema_decay)
                new_params[lo:hi] = p
                off += n
            new_params = new_params.to(tail.dtype).float()  # owners round; zero elsewhere: the sum is the all-gather
            dist.all_reduce(new_params, group=group)
            params.copy_(new_params.to(tail.dtype))
        tail.grad_buffers[g].tensor.zero_()
    return tail.state


class ReferenceTailEngine(nn.Module):
    """Drop-in for ``SymmDataParallel`` whose tail is ``reference_tail`` (any device, any backend)."""

    want_fused_tail = True

    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: int = 25):
        super().__init__()
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        device = next(module.parameters()).device
        self.comm = PlainComm(self.process_group, device)
        self.bucket_bytes = max(1, int(bucket_cap_mb * 1024 * 1024))
        self.accumulate_grads = False
        self._grad_buffers: List[_PlainBuffer] = []
        self._param_buffers: List[_PlainBuffer] = []
        self.tail: Optional[FusedTail] = None
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

    def begin_optimizer_build(self):
        self._grad_buffers, self._param_buffers = [], []

    def _alloc(self, pool, numel, dtype, device):
        buf = _PlainBuffer(torch.zeros(-(-numel // 8) * 8, dtype=dtype, device=device), self.comm.rank, self.comm.world)
        pool.append(buf)
        return buf.tensor[:numel]

    def alloc_grad_buffer(self, numel, dtype, device):
        return self._alloc(self._grad_buffers, numel, dtype, device)

    def alloc_param_buffer(self, numel, dtype, device):
        if dtype not in (torch.float16, torch.bfloat16):
            return None
        return self._alloc(self._param_buffers, numel, dtype, device)

    def attach_optimizer(self, optimizer, params=None):
        self.tail = None
        flats = [f for g in getattr(optimizer, "fp16_params", []) for f in g["params"]]

        def find(pool, t):
            return next((b for b in pool if b.tensor.data_ptr() == t.data_ptr()), None)

        gbufs = [find(self._grad_buffers, f.grad) for f in flats]
        pbufs = [find(self._param_buffers, f.data) for f in flats]
        if not flats or any(b is None for b in gbufs + pbufs) or not hasattr(optimizer, "enable_fused_tail"):
            return
        tail = FusedTail(self.comm, gbufs, pbufs, self.bucket_bytes, seed=getattr(optimizer.args, "seed", 0))
        if optimizer.enable_fused_tail(self, tail):
            self.tail = tail

    def all_reduce_grads(self):
        if self.tail is None and not self.accumulate_grads:
            for p in self.module.parameters():
                if p.grad is not None:
                    p.grad.div_(self.world_size)
                    dist.all_reduce(p.grad, group=self.process_group)

    def run_tail(self, **kwargs):
        return reference_tail(self.tail, **kwargs)

    def reduce_stats(self, values):
        out = values.detach().to(torch.float64).clone()
        dist.all_reduce(out, group=self.process_group)
        return out

    def check_health(self):
        pass
