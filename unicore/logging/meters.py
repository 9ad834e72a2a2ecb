"""Value holders behind the metrics system.

Three kinds of meter - a weighted mean, a rate and a stopwatch - plus a priority-ordered collection.  The names and the
serialised state (``extra_state.metrics`` stores ``(priority, key, class name, state)`` rows) are the reference's
(``unicore/logging/meters.py``: ``AverageMeter:68``, ``TimeMeter:113``, ``StopwatchMeter:166``, ``MetersDict:222``,
``safe_round:57``) because checkpoints and log scrapers depend on them.  What is different here: samples may be device
scalars and stay on the device while they accumulate - a meter never forces a host read inside the training step;
``MetersDict.localize`` brings everything to the host in one transfer when values are about to be displayed.
"""
import time
from collections import OrderedDict
from typing import Dict, Optional

try:
    import torch
except ImportError:  # pragma: no cover
    torch = None

try:
    import numpy as np
except ImportError:  # pragma: no cover
    np = None


def _is_tensor(x) -> bool:
    return torch is not None and torch.is_tensor(x)


def _clock() -> float:
    return time.perf_counter()


def safe_round(number, ndigits):
    """``round`` that also accepts one-element tensors and numpy scalars; anything else passes through."""
    if hasattr(number, "__round__"):
        return round(number, ndigits)
    if _is_tensor(number) and number.numel() == 1:
        return safe_round(number.item(), ndigits)
    if np is not None and np.ndim(number) == 0 and hasattr(number, "item"):
        return safe_round(number.item(), ndigits)
    return number


class Meter(object):
    """What the collection expects of a meter."""

    round: Optional[int] = None

    def reset(self):
        raise NotImplementedError

    def state_dict(self):
        return {}

    def load_state_dict(self, state_dict):
        pass

    @property
    def smoothed_value(self) -> float:
        raise NotImplementedError

    def _for_display(self, value):
        return value if (self.round is None or value is None) else safe_round(value, self.round)


class AverageMeter(Meter):
    """Weighted running mean (``sum`` / ``count``); ``val`` keeps the latest sample."""

    def __init__(self, round: Optional[int] = None):
        self.round = round
        self.reset()

    def reset(self):
        self.val, self.sum, self.count = None, 0, 0

    def update(self, val, n=1):
        if val is None:
            return
        self.val = val
        # a device-resident weight is taken as given: looking at its sign would be a host read
        if _is_tensor(n) or n > 0:
            self.sum = self.sum + val * n   # python number x tensor stays on the tensor's device
            self.count = self.count + n

    @property
    def avg(self):
        return self.val if not self.count > 0 else self.sum / self.count

    @property
    def smoothed_value(self):
        return self._for_display(self.avg)

    def state_dict(self):
        return dict(val=self.val, sum=self.sum, count=self.count, round=self.round)

    def load_state_dict(self, state_dict):
        self.val, self.sum, self.count = state_dict["val"], state_dict["sum"], state_dict["count"]
        self.round = state_dict.get("round", None)


class TimeMeter(Meter):
    """Events per second since the last ``reset`` (updates/s, words/s); ``init`` carries time from before a resume."""

    def __init__(self, init: int = 0, n: int = 0, round: Optional[int] = None):
        self.round = round
        self.reset(init, n)

    def reset(self, init=0, n=0):
        self.init, self.n, self.i = init, n, 0
        self.start = _clock()

    def update(self, val=1):
        self.n = self.n + val
        self.i += 1

    @property
    def elapsed_time(self):
        return self.init + (_clock() - self.start)

    @property
    def avg(self):
        return self.n / self.elapsed_time

    @property
    def smoothed_value(self):
        return self._for_display(self.avg)

    def state_dict(self):
        return dict(init=self.elapsed_time, n=self.n, round=self.round)

    def load_state_dict(self, state_dict):
        legacy = "start" in state_dict   # very old checkpoints stored an absolute start time and no count
        self.reset(init=state_dict["init"], n=0 if legacy else state_dict["n"])
        if not legacy:
            self.round = state_dict.get("round", None)


class StopwatchMeter(Meter):
    """Total length of ``start()`` ... ``stop()`` intervals (``train_wall``); shows the running interval until one ends."""

    def __init__(self, round: Optional[int] = None):
        self.round = round
        self.sum, self.n = 0, 0
        self.start_time = None

    def start(self):
        self.start_time = _clock()

    def stop(self, n=1, prehook=None):
        if self.start_time is None:
            return
        if prehook is not None:
            prehook()
        self.sum = self.sum + (_clock() - self.start_time)
        self.n = self.n + n

    def reset(self):
        self.sum, self.n = 0, 0
        self.start()

    @property
    def avg(self):
        return self.sum if not self.n > 0 else self.sum / self.n

    @property
    def elapsed_time(self):
        return 0.0 if self.start_time is None else _clock() - self.start_time

    @property
    def smoothed_value(self):
        return self._for_display(self.avg if self.sum > 0 else self.elapsed_time)

    def state_dict(self):
        return dict(sum=self.sum, n=self.n, round=self.round)

    def load_state_dict(self, state_dict):
        self.sum, self.n = state_dict["sum"], state_dict["n"]
        self.start_time = None
        self.round = state_dict.get("round", None)


class MetersDict(OrderedDict):
    """Meters by name, iterated in ``(priority, insertion)`` order.  Items are set as ``(priority, meter)`` pairs (or
    through ``add_meter``) and read back as the meter; a name cannot be assigned twice."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.priorities = []   # sorted (priority, arrival number, key) - also the serialisation order

    def __setitem__(self, key, value):
        if key in self:
            raise KeyError("MetersDict doesn't support reassignment")
        priority, meter = value
        self.priorities.append((priority, len(self.priorities), key))
        self.priorities.sort()
        super().__setitem__(key, meter)
        for _, _, name in self.priorities:   # keep the underlying order equal to the priority order
            self.move_to_end(name)

    def add_meter(self, key, meter, priority):
        self[key] = (priority, meter)

    # -- (de)serialisation -------------------------------------------------------------------------------------------
    def state_dict(self):
        rows = []
        for priority, _, key in self.priorities:
            meter = self[key]
            if isinstance(meter, MetersDict._DerivedMeter):
                continue   # a function of the other meters: nothing to store
            rows.append((priority, key, type(meter).__name__, meter.state_dict()))
        return rows

    def load_state_dict(self, state_dict):
        self.clear()
        self.priorities.clear()
        kinds = {cls.__name__: cls for cls in (AverageMeter, TimeMeter, StopwatchMeter)}
        for priority, key, class_name, state in state_dict:
            meter = kinds[class_name]() if class_name in kinds else globals()[class_name]()
            meter.load_state_dict(state)
            self.add_meter(key, meter, priority)

    # -- display -----------------------------------------------------------------------------------------------------
    def get_smoothed_value(self, key: str) -> float:
        meter = self[key]
        return meter.fn(self) if isinstance(meter, MetersDict._DerivedMeter) else meter.smoothed_value

    def localize(self) -> None:
        """Replace device scalars held by the meters (``val`` / ``sum`` / ``count``) by python numbers with ONE
        device-to-host transfer per number class, instead of one blocking ``.item()`` per displayed value - a dozen
        reads per ``get_smoothed_values`` call, i.e. per training step, otherwise.  Accumulation continues from the
        host copies."""
        if torch is None:
            return
        floats, ints = [], []
        for meter in self.values():
            for field in ("val", "sum", "count"):
                held = getattr(meter, field, None)
                if _is_tensor(held) and held.is_cuda and held.numel() == 1:
                    (floats if held.is_floating_point() else ints).append((meter, field, held))
        if not floats and not ints:
            return
        try:
            from unicore.utils import tolist
        except ImportError:  # pragma: no cover
            def tolist(t):
                return t.tolist()
        for group, widen in ((floats, True), (ints, False)):
            if not group:
                continue
            packed = torch.cat([held.detach().reshape(1) for _, _, held in group])
            numbers = tolist(packed.double() if widen else packed)
            for (meter, field, _), number in zip(group, numbers):
                setattr(meter, field, number)

    def get_smoothed_values(self) -> Dict[str, float]:
        self.localize()
        return OrderedDict((key, self.get_smoothed_value(key)) for key in self.keys() if not key.startswith("_"))

    def reset(self):
        for meter in self.values():
            if not isinstance(meter, MetersDict._DerivedMeter):
                meter.reset()

    class _DerivedMeter(Meter):
        """Value computed from the other meters of the collection (``fn(meters_dict)``)."""

        def __init__(self, fn):
            self.fn = fn

        def reset(self):
            pass
