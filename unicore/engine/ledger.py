"""One update's (or one validation batch's) logging statistics as a flat fp64 vector, and its sum over the ranks.

The reference sums scalar logging outputs with a per-step NCCL all-reduce of a freshly built fp64 buffer
(``unicore/trainer.py:1011-1049``) and falls back to pickling when a loss says its outputs cannot be summed
(``:985-1009``).  Here the scalars are packed once (``utils.stack_scalars``: device scalars stacked, host numbers in one
pinned transfer) and the sum travels

* inside the fused optimizer tail (training, ``--ddp-backend b200``): no launch and no collective of its own,
* on the 64-thread peer-memory kernel (``SymmDataParallel.reduce_stats``) where that engine is present,
* through ``distributed_utils.all_reduce`` (NCCL / gloo) otherwise.

Entries keep their position; values are read back as 0-dim views of the summed vector, i.e. without further launches
and without a host sync.
"""
from itertools import chain
from typing import Any, Dict, List, Optional

import torch

from unicore import utils
from unicore.distributed import utils as distributed_utils


class StatLedger:
    def __init__(self):
        self._keys: List[str] = []
        self._values: List[Any] = []
        self._pos: Optional[Dict[str, int]] = None  # key -> position in the packed vector
        self.sums: Optional[torch.Tensor] = None     # fp64 vector of rank sums (packed order)

    def add(self, key: str, value) -> None:
        """Register ``value`` (python number or 0-dim tensor) under ``key``."""
        assert self._pos is None, "the ledger is already packed"
        self._keys.append(key)
        self._values.append(value)

    def __len__(self):
        return len(self._keys)

    def pack(self, device) -> torch.Tensor:
        """The entries as one fp64 vector: device-resident scalars first (one stack per dtype), host numbers behind
        them (one pinned transfer) - the same order on every rank because the keys and their kinds are."""
        on_device = [i for i, v in enumerate(self._values) if torch.is_tensor(v) and v.is_cuda]
        on_host = [i for i in range(len(self._values)) if i not in set(on_device)]
        by_dtype = {}
        for i in on_device:  # (the packing stacks device scalars per dtype: keep that grouping so no reorder is needed)
            by_dtype.setdefault(self._values[i].dtype, []).append(i)
        order = [i for group in by_dtype.values() for i in group] + on_host
        self._pos = {self._keys[i]: pos for pos, i in enumerate(order)}
        return utils.stack_scalars([self._values[i] for i in order], device=device)

    def position(self, key: str) -> int:
        return self._pos[key]

    def adopt(self, sums: torch.Tensor) -> None:
        """``sums``: the rank sums in packed order (may still be in flight on the device)."""
        self.sums = sums

    def value(self, key: str):
        if self.sums is None:
            return self._values[self._keys.index(key)]
        return self.sums[self._pos[key]]

    # -- lists of logging-output dicts ------------------------------------------------------------------------
    def add_logging_outputs(self, logging_outputs: List[Dict[str, Any]], ignore: bool = False) -> List[str]:
        """Sum the per-micro-batch dicts locally and register one entry per key; returns the keys."""
        if not logging_outputs:
            return []
        keys = list(logging_outputs[0].keys())
        for k in keys:
            if ignore:  # a dummy batch contributes zeros but must keep the vector layout of the other ranks
                first = logging_outputs[0][k]
                total = torch.zeros_like(first) if torch.is_tensor(first) else 0
            else:
                total = sum(log[k] for log in logging_outputs if k in log)
            self.add("log:" + k, total)
        return keys

    def logging_output(self, keys: List[str]) -> List[Dict[str, Any]]:
        return [{k: self.value("log:" + k) for k in keys}] if keys else []


def reduce_ledger(ledger: StatLedger, device, group, engine=None) -> None:
    """Sum the ledger over the ranks now (validation, or training without the fused tail)."""
    if len(ledger) == 0:
        return
    if engine is not None and hasattr(engine, "reduce_stats") and device.type == "cuda" and len(ledger) <= 64:
        ledger.adopt(engine.reduce_stats(ledger.pack(device)))
        return
    comm_device = distributed_utils._backend_device()  # noqa: SLF001
    buf = ledger.pack(comm_device)
    distributed_utils.all_reduce(buf, group=group)
    ledger.adopt(buf.to(device))


def gather_objects(logging_outputs, extras, group, max_size: int, ignore: bool = False):
    """Logging outputs that cannot be summed: gather the pickled dicts of all ranks and add up the extras."""
    if ignore:
        logging_outputs = []
    rows = distributed_utils.all_gather_list([logging_outputs] + list(extras), max_size=max_size, group=group)
    columns = list(zip(*rows))
    merged = list(chain.from_iterable(columns[0]))
    return merged, [sum(col) for col in columns[1:]]
