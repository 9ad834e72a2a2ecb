"""Value holders behind the metrics system: running averages, rates, stopwatches.

API parity with reference ``unicore/logging/meters.py`` (``AverageMeter:68``, ``TimeMeter:113``,
``StopwatchMeter:166``, ``MetersDict:222``, ``safe_round:57``): every meter has ``reset``,
``update``, ``state_dict``/``load_state_dict`` and a ``smoothed_value`` used for display, and the
class names are what checkpoints record (``extra_state.metrics`` stores ``(priority, key,
class-name, state)``).  Values may be tensors; they are only converted to Python numbers at
display time so logging never forces a device sync inside the step.
"""
import bisect
import time
from collections import OrderedDict
from typing import Dict, Optional

try:
    import torch

    def _is_tensor(x):
        return torch.is_tensor(x)
except ImportError:  # pragma: no cover
    torch = None

    def _is_tensor(x):
        return False

try:
    import numpy as np
except ImportError:  # pragma: no cover
    np = None


def _scalar_type(value, n):
    """Cast ``n`` to something multiplicable with ``value`` without leaving its device."""
    # a python number multiplies a tensor in place on its device; materialising it with
    # ``new_tensor`` would be a blocking pageable host->device copy per logged scalar
    return n


def safe_round(number, ndigits):
    if hasattr(number, "__round__"):
        return round(number, ndigits)
    if torch is not None and _is_tensor(number) and number.numel() == 1:
        return safe_round(number.item(), ndigits)
    if np is not None and np.ndim(number) == 0 and hasattr(number, "item"):
        return safe_round(number.item(), ndigits)
    return number


class Meter(object):
    """Interface all meters implement."""

    def reset(self):
        raise NotImplementedError

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass

    @property
    def smoothed_value(self) -> float:
        raise NotImplementedError


class AverageMeter(Meter):
    """Weighted running mean; ``val`` is the last sample."""

    def __init__(self, round: Optional[int] = None):
        self.round = round
        self.reset()

    def reset(self):
        self.val = None
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        if val is None:
            return
        self.val = val
        if _is_tensor(n) or n > 0:  # device-resident weights are not inspected (no host sync)
            self.sum = self.sum + val * _scalar_type(val, n)
            self.count = self.count + n

    @property
    def avg(self):
        return self.sum / self.count if self.count > 0 else self.val

    @property
    def smoothed_value(self):
        value = self.avg
        if self.round is not None and value is not None:
            value = safe_round(value, self.round)
        return value

    def state_dict(self):
        return {"val": self.val, "sum": self.sum, "count": self.count, "round": self.round}

    def load_state_dict(self, state_dict):
        self.val = state_dict["val"]
        self.sum = state_dict["sum"]
        self.count = state_dict["count"]
        self.round = state_dict.get("round", None)


class TimeMeter(Meter):
    """Events per second since ``reset`` (e.g. updates/s)."""

    def __init__(self, init: int = 0, n: int = 0, round: Optional[int] = None):
        self.round = round
        self.reset(init, n)

    def reset(self, init=0, n=0):
        self.init = init
        self.start = time.perf_counter()
        self.n = n
        self.i = 0

    def update(self, val=1):
        self.n = self.n + val
        self.i += 1

    @property
    def elapsed_time(self):
        return self.init + (time.perf_counter() - self.start)

    @property
    def avg(self):
        return self.n / self.elapsed_time

    @property
    def smoothed_value(self):
        value = self.avg
        if self.round is not None and value is not None:
            value = safe_round(value, self.round)
        return value

    def state_dict(self):
        return {"init": self.elapsed_time, "n": self.n, "round": self.round}

    def load_state_dict(self, state_dict):
        if "start" in state_dict:  # very old checkpoints stored an absolute start time
            self.reset(init=state_dict["init"])
        else:
            self.reset(init=state_dict["init"], n=state_dict["n"])
            self.round = state_dict.get("round", None)


class StopwatchMeter(Meter):
    """Accumulated duration of start/stop intervals (e.g. ``train_wall``)."""

    def __init__(self, round: Optional[int] = None):
        self.round = round
        self.sum = 0
        self.n = 0
        self.start_time = None

    def start(self):
        self.start_time = time.perf_counter()

    def stop(self, n=1, prehook=None):
        if self.start_time is None:
            return
        if prehook is not None:
            prehook()
        self.sum = self.sum + (time.perf_counter() - self.start_time)
        self.n = self.n + n

    def reset(self):
        self.sum = 0
        self.n = 0
        self.start()

    @property
    def avg(self):
        return self.sum / self.n if self.n > 0 else self.sum

    @property
    def elapsed_time(self):
        return 0.0 if self.start_time is None else time.perf_counter() - self.start_time

    @property
    def smoothed_value(self):
        value = self.avg if self.sum > 0 else self.elapsed_time
        if self.round is not None and value is not None:
            value = safe_round(value, self.round)
        return value

    def state_dict(self):
        return {"sum": self.sum, "n": self.n, "round": self.round}

    def load_state_dict(self, state_dict):
        self.sum = state_dict["sum"]
        self.n = state_dict["n"]
        self.start_time = None
        self.round = state_dict.get("round", None)


class MetersDict(OrderedDict):
    """Ordered dict of meters sorted by ``(priority, insertion order)``; serialisable."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.priorities = []

    def __setitem__(self, key, value):
        if key in self:
            raise KeyError("MetersDict doesn't support reassignment")
        priority, meter = value
        rank = (priority, len(self.priorities), key)
        bisect.insort(self.priorities, rank)
        super().__setitem__(key, meter)
        # re-thread the OrderedDict so iteration follows priority order
        for _, _, k in self.priorities:
            self.move_to_end(k)

    def add_meter(self, key, meter, priority):
        self[key] = (priority, meter)

    def state_dict(self):
        return [
            (pri, key, self[key].__class__.__name__, self[key].state_dict())
            for pri, _, key in self.priorities
            if not isinstance(self[key], MetersDict._DerivedMeter)  # derived meters hold lambdas
        ]

    def load_state_dict(self, state_dict):
        self.clear()
        self.priorities.clear()
        for pri, key, class_name, state in state_dict:
            meter = globals()[class_name]()
            meter.load_state_dict(state)
            self.add_meter(key, meter, pri)

    def get_smoothed_value(self, key: str) -> float:
        meter = self[key]
        if isinstance(meter, MetersDict._DerivedMeter):
            return meter.fn(self)
        return meter.smoothed_value

    def localize(self) -> None:
        """Bring device-resident meter state to the host with ONE transfer per dtype.

        Every ``smoothed_value`` of a meter holding CUDA scalars ends in ``.item()``: a dozen
        blocking device reads per ``get_smoothed_values`` call, i.e. per training step.  Here all
        scalar CUDA tensors held by the meters (val / sum / count) are concatenated, read once and
        written back as python numbers; accumulation simply continues from those.
        """
        if torch is None:
            return
        slots = []
        for meter in self.values():
            for attr in ("val", "sum", "count"):
                v = getattr(meter, attr, None)
                if _is_tensor(v) and v.is_cuda and v.numel() == 1:
                    slots.append((meter, attr, v))
        if not slots:
            return
        try:
            from unicore.utils import tolist
        except ImportError:  # pragma: no cover
            def tolist(t):
                return t.tolist()
        for is_float in (True, False):
            group = [sl for sl in slots if sl[2].is_floating_point() == is_float]
            if not group:
                continue
            flat = torch.cat([v.detach().reshape(1) for _, _, v in group])
            if is_float:
                flat = flat.double()
            for (meter, attr, _), x in zip(group, tolist(flat)):
                setattr(meter, attr, x)

    def get_smoothed_values(self) -> Dict[str, float]:
        self.localize()
        return OrderedDict(
            (key, self.get_smoothed_value(key)) for key in self.keys() if not key.startswith("_")
        )

    def reset(self):
        for meter in self.values():
            if isinstance(meter, MetersDict._DerivedMeter):
                continue
            meter.reset()

    class _DerivedMeter(Meter):
        """A meter whose value is a function of the other meters in the dict."""

        def __init__(self, fn):
            self.fn = fn

        def reset(self):
            pass
