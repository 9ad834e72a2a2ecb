"""Process-global, nestable metric aggregation.

Any code may call ``metrics.log_scalar("loss", x)``; the value lands in every *active*
aggregator.  Aggregators are activated with ``with metrics.aggregate("train"):`` (also usable as
a decorator); the ``"default"`` aggregator is always active.  ``new_root=True`` temporarily
hides all outer aggregators (used for validation so that validation stats do not pollute the
training meters).  The whole state is checkpointable (``state_dict``/``load_state_dict``).

Parity: reference ``unicore/logging/metrics.py`` (``aggregate:46``, ``log_scalar:112``,
``log_derived:135``, ``log_speed:149``, ``log_start_time:171``, ``log_stop_time:187``,
``log_custom:205``, ``reset_*:236-250``, ``get_*:253-279``, ``state_dict:281``).
"""
import contextlib
import uuid
from collections import OrderedDict, defaultdict
from typing import Callable, Dict, List, Optional

from .meters import AverageMeter, Meter, MetersDict, StopwatchMeter, TimeMeter

_aggregators: "OrderedDict[str, MetersDict]" = OrderedDict()
_active: "OrderedDict[str, MetersDict]" = OrderedDict()
_refcount: Dict[str, int] = defaultdict(int)


def reset() -> None:
    """Forget every aggregator and re-create the always-on ``default`` one."""
    _aggregators.clear()
    _active.clear()
    _refcount.clear()
    _aggregators["default"] = MetersDict()
    _active["default"] = _aggregators["default"]
    _refcount["default"] = 1


reset()


@contextlib.contextmanager
def aggregate(name: Optional[str] = None, new_root: bool = False):
    """Activate the aggregator ``name`` (anonymous + temporary when ``None``) for the block."""
    temporary = name is None
    if temporary:
        name = str(uuid.uuid4())
        agg = MetersDict()
    else:
        if name == "default":
            raise ValueError("'default' is reserved")
        agg = _aggregators.setdefault(name, MetersDict())

    saved_active = saved_refs = None
    if new_root:
        saved_active = _active.copy()
        saved_refs = dict(_refcount)
        _active.clear()
        _refcount.clear()

    _active[name] = agg
    _refcount[name] += 1
    try:
        yield agg
    finally:
        _refcount[name] -= 1
        if _refcount[name] == 0 and name in _active:
            del _active[name]
        if new_root:
            _active.clear()
            _active.update(saved_active)
            _refcount.clear()
            _refcount.update(saved_refs)


def get_active_aggregators() -> List[MetersDict]:
    return list(_active.values())


def log_scalar(key: str, value, weight: float = 1, priority: int = 10, round: Optional[int] = None):
    """Record ``value`` with ``weight`` into a weighted-average meter named ``key``."""
    for agg in get_active_aggregators():
        if key not in agg:
            agg.add_meter(key, AverageMeter(round=round), priority)
        agg[key].update(value, weight)


def log_derived(key: str, fn: Callable[[MetersDict], float], priority: int = 20):
    """Register a value computed from the other meters of the same aggregator."""
    for agg in get_active_aggregators():
        if key not in agg:
            agg.add_meter(key, MetersDict._DerivedMeter(fn), priority)


def log_speed(key: str, value, priority: int = 30, round: Optional[int] = None):
    """Record a rate (events / second)."""
    for agg in get_active_aggregators():
        if key not in agg:
            agg.add_meter(key, TimeMeter(round=round), priority)
            agg[key].reset()  # the first call only starts the clock
        else:
            agg[key].update(value)


def log_start_time(key: str, priority: int = 40, round: Optional[int] = None):
    for agg in get_active_aggregators():
        if key not in agg:
            agg.add_meter(key, StopwatchMeter(round=round), priority)
        agg[key].start()


def log_stop_time(key: str, weight: float = 0.0, prehook=None):
    for agg in get_active_aggregators():
        if key in agg:
            agg[key].stop(weight, prehook)


def log_custom(new_meter_fn: Callable[[], Meter], key: str, *args, priority: int = 50, **kwargs):
    for agg in get_active_aggregators():
        if key not in agg:
            agg.add_meter(key, new_meter_fn(), priority)
        agg[key].update(*args, **kwargs)


def reset_meter(name: str, key: str) -> None:
    meter = get_meter(name, key)
    if meter is not None:
        meter.reset()


def reset_meters(name: str) -> None:
    meters = get_meters(name)
    if meters is not None:
        meters.reset()


def get_meter(name: str, key: str) -> Optional[Meter]:
    if name not in _aggregators:
        return None
    return _aggregators[name].get(key, None)


def get_meters(name: str) -> Optional[MetersDict]:
    return _aggregators.get(name, None)


def get_smoothed_value(name: str, key: str) -> float:
    return _aggregators[name].get_smoothed_value(key)


def get_smoothed_values(name: str) -> Dict[str, float]:
    return _aggregators[name].get_smoothed_values()


def state_dict():
    return OrderedDict((name, agg.state_dict()) for name, agg in _aggregators.items())


def load_state_dict(state):
    for name, agg_state in state.items():
        _aggregators[name] = MetersDict()
        _aggregators[name].load_state_dict(agg_state)
        if name in _active:  # re-point live contexts at the restored meters
            _active[name] = _aggregators[name]
    # keep ``default`` wired as the always-active aggregator after a reload
    if "default" in _aggregators:
        _active["default"] = _aggregators["default"]
        _refcount["default"] = max(_refcount["default"], 1)
