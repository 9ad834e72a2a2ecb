"""Process-global, nestable metric aggregation.

Any code may call ``metrics.log_scalar("loss", x)``; the sample lands in every aggregator that is *open* at that
moment.  ``with metrics.aggregate("train"):`` (also a decorator) opens the named aggregator for the block, the
``"default"`` aggregator is always open, ``aggregate(new_root=True)`` hides the enclosing ones for the block (validation
must not pollute the training meters) and ``aggregate()`` without a name opens a throw-away one.  Everything named is
checkpointable (``state_dict`` / ``load_state_dict``).

The function names and their defaults are the reference's (``unicore/logging/metrics.py``: ``aggregate:46``,
``log_scalar:112``, ``log_derived:135``, ``log_speed:149``, ``log_start_time:171``, ``log_stop_time:187``,
``log_custom:205``, ``reset_*:236-250``, ``get_*:253-279``, ``state_dict:281``); the bookkeeping is one small registry
object: named aggregators plus the stack of scopes that are currently open.
"""
import contextlib
import uuid
from collections import OrderedDict
from typing import Callable, Dict, List, Optional

from .meters import AverageMeter, Meter, MetersDict, StopwatchMeter, TimeMeter

_DEFAULT = "default"


class _Registry:
    def __init__(self):
        self.named: "OrderedDict[str, MetersDict]" = OrderedDict()
        self.scopes: List[List] = []   # open scopes, outermost first: [name, aggregator, hides_outer]
        self.named[_DEFAULT] = MetersDict()

    def open_aggregators(self) -> List[MetersDict]:
        """Aggregators that receive samples now: ``default`` (unless a root scope hides it) + the open scopes inwards of
        the innermost root scope, each once."""
        start = 0
        for index, (_, _, hides_outer) in enumerate(self.scopes):
            if hides_outer:
                start = index
        rooted = any(scope[2] for scope in self.scopes)
        seen, out = set(), []
        if not rooted:
            out.append(self.named[_DEFAULT])
            seen.add(id(out[0]))
        for _, agg, _ in self.scopes[start:]:
            if id(agg) not in seen:
                seen.add(id(agg))
                out.append(agg)
        return out

    def rebind(self, name, agg):
        self.named[name] = agg
        for scope in self.scopes:
            if scope[0] == name:
                scope[1] = agg


_registry = _Registry()


def reset() -> None:
    """Forget every aggregator (and open scope) and start over with an empty ``default``."""
    global _registry
    _registry = _Registry()


class aggregate(contextlib.ContextDecorator):
    """Open the aggregator ``name`` for a block / a function call; yields the ``MetersDict``."""

    def __init__(self, name: Optional[str] = None, new_root: bool = False):
        if name == _DEFAULT:
            raise ValueError("'default' is reserved")
        self.name, self.new_root = name, new_root
        self._opened = []

    def __enter__(self) -> MetersDict:
        if self.name is None:   # anonymous: lives for this block only
            label, agg = str(uuid.uuid4()), MetersDict()
        else:
            label, agg = self.name, _registry.named.setdefault(self.name, MetersDict())
        scope = [label, agg, self.new_root]
        _registry.scopes.append(scope)
        self._opened.append(scope)
        return agg

    def __exit__(self, *exc):
        scope = self._opened.pop()
        for index in range(len(_registry.scopes) - 1, -1, -1):   # identity, not equality: remove THIS scope
            if _registry.scopes[index] is scope:
                del _registry.scopes[index]
                break
        return False


def get_active_aggregators() -> List[MetersDict]:
    return _registry.open_aggregators()


def _meters_named(key: str, make: Callable[[], Meter], priority: int):
    """The meter ``key`` of every open aggregator (created on first use); yields ``(meter, created)``."""
    for agg in _registry.open_aggregators():
        created = key not in agg
        if created:
            agg.add_meter(key, make(), priority)
        yield agg[key], created


def log_scalar(key: str, value, weight: float = 1, priority: int = 10, round: Optional[int] = None):
    """One sample of a weighted average."""
    for meter, _ in _meters_named(key, lambda: AverageMeter(round=round), priority):
        meter.update(value, weight)


def log_derived(key: str, fn: Callable[[MetersDict], float], priority: int = 20):
    """A value computed from the other meters of the same aggregator at display time."""
    for _ in _meters_named(key, lambda: MetersDict._DerivedMeter(fn), priority):
        pass


def log_speed(key: str, value, priority: int = 30, round: Optional[int] = None):
    """Events per second; the first call of a key only starts its clock."""
    for meter, created in _meters_named(key, lambda: TimeMeter(round=round), priority):
        if created:
            meter.reset()
        else:
            meter.update(value)


def log_start_time(key: str, priority: int = 40, round: Optional[int] = None):
    for meter, _ in _meters_named(key, lambda: StopwatchMeter(round=round), priority):
        meter.start()


def log_stop_time(key: str, weight: float = 0.0, prehook=None):
    for agg in _registry.open_aggregators():
        if key in agg:
            agg[key].stop(weight, prehook)


def log_custom(new_meter_fn: Callable[[], Meter], key: str, *args, priority: int = 50, **kwargs):
    for meter, _ in _meters_named(key, new_meter_fn, priority):
        meter.update(*args, **kwargs)


# ---- access by aggregator name ------------------------------------------------------------------------------------------
def get_meters(name: str) -> Optional[MetersDict]:
    return _registry.named.get(name, None)


def get_meter(name: str, key: str) -> Optional[Meter]:
    agg = get_meters(name)
    return None if agg is None else agg.get(key, None)


def reset_meter(name: str, key: str) -> None:
    meter = get_meter(name, key)
    if meter is not None:
        meter.reset()


def reset_meters(name: str) -> None:
    agg = get_meters(name)
    if agg is not None:
        agg.reset()


def get_smoothed_value(name: str, key: str) -> float:
    return _registry.named[name].get_smoothed_value(key)


def get_smoothed_values(name: str) -> Dict[str, float]:
    return _registry.named[name].get_smoothed_values()


# ---- checkpointing ----------------------------------------------------------------------------------------------------------
def state_dict():
    return OrderedDict((name, agg.state_dict()) for name, agg in _registry.named.items())


def load_state_dict(state):
    for name, rows in state.items():
        restored = MetersDict()
        restored.load_state_dict(rows)
        _registry.rebind(name, restored)   # scopes that are open right now continue on the restored meters
