"""Explicit (hook-free, non-overlapped) data parallelism.

The wrapper does nothing during backward; the trainer calls ``all_reduce_grads()`` afterwards,
which averages gradients over the group by packing them into a bounded flat buffer per
(device, dtype) and all-reducing it.  Selected by ``--ddp-backend no_c10d`` (alias
``legacy_ddp``); also the engine used for the CPU/gloo plumbing configuration.
Parity: reference ``unicore/distributed/legacy_distributed_data_parallel.py:27-166`` (buffer of at
most ``buffer_size`` elements, pre-division by the world size, ``no_sync`` context, params flagged
``expert`` are skipped, ``all_reduce_params`` for the fp32-master-grad path).
"""
from collections import OrderedDict
from contextlib import contextmanager

import torch
from torch import nn

from . import utils as dist_utils


class LegacyDistributedDataParallel(nn.Module):
    def __init__(self, module, process_group, buffer_size=2 ** 28):
        super().__init__()
        self.module = module
        self.process_group = process_group
        self.world_size = dist_utils.get_world_size(process_group)
        # never allocate more than the model itself needs
        self.buffer_size = min(buffer_size, sum(p.numel() for p in module.parameters()))
        self.accumulate_grads = False
        self._scratch = {}
        groups = OrderedDict()
        for p in module.parameters():
            groups.setdefault(p.device, []).append(p)
        self.per_device_params = list(groups.values())

    @contextmanager
    def no_sync(self):
        """Skip the reduction for backward passes run inside the block (grad accumulation)."""
        previous, self.accumulate_grads = self.accumulate_grads, True
        try:
            yield
        finally:
            self.accumulate_grads = previous

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)

    # -- reduction ----------------------------------------------------------------------------
    def _scratch_for(self, like: torch.Tensor, numel: int) -> torch.Tensor:
        key = (like.device, like.dtype)
        buf = self._scratch.get(key)
        if buf is None or buf.numel() < numel:
            buf = self._scratch[key] = like.new_empty(max(numel, min(self.buffer_size, 1 << 20)))
        return buf[:numel]

    def _reduce_chunk(self, params):
        """Average the grads of ``params`` (same device + dtype) over the group, in place."""
        if len(params) == 1 and params[0].grad is not None:
            g = params[0].grad.data
            if self.world_size > 1:
                g.div_(self.world_size)
                dist_utils.all_reduce(g, group=self.process_group)
            return
        total = sum(p.numel() for p in params)
        flat = self._scratch_for(params[0], total)
        offset = 0
        for p in params:
            n = p.numel()
            if p.grad is not None:
                flat[offset:offset + n].copy_(p.grad.data.view(-1))
            else:
                flat[offset:offset + n].zero_()
            offset += n
        if self.world_size > 1:
            flat.div_(self.world_size)
            dist_utils.all_reduce(flat, group=self.process_group)
        offset = 0
        for p in params:
            n = p.numel()
            if p.grad is not None:
                p.grad.data.copy_(flat[offset:offset + n].view_as(p))
            else:
                p.grad = flat[offset:offset + n].view_as(p).clone()
            offset += n

    def all_reduce_params(self, params):
        """Average grads of an explicit param list (also used for the flat fp32 master grads)."""
        if self.accumulate_grads:
            return
        pending = OrderedDict()  # dtype -> (params, numel)
        for p in params:
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            if getattr(p, "expert", False):
                continue  # expert-parallel params are not replicated
            n = p.numel()
            if n > self.buffer_size:
                self._reduce_chunk([p])
                continue
            bucket, used = pending.get(p.dtype, ([], 0))
            if used + n > self.buffer_size:
                self._reduce_chunk(bucket)
                bucket, used = [], 0
            bucket.append(p)
            pending[p.dtype] = (bucket, used + n)
        for bucket, _ in pending.values():
            if bucket:
                self._reduce_chunk(bucket)

    def all_reduce_grads(self):
        """Called by the trainer after the last micro-batch's backward."""
        for params in self.per_device_params:
            self.all_reduce_params(params)
