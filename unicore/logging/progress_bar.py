"""Progress reporting: one iterable wrapper (:class:`ProgressReporter`) + pluggable *sinks*.

The training loop iterates over the reporter and calls ``log(stats)`` after every update and ``print(stats)`` at the end
of an epoch / validation pass.  The reporter only keeps the position; what happens with the numbers is decided by its
sinks:

    ConsoleSink("json" | "simple")   a line every ``log_interval`` updates and a summary line
    TqdmSink                         live bar on a terminal
    BoardSink                        TensorBoard scalars (one writer per tag) and, optionally, Weights & Biases

``--log-format`` picks the console sink, ``--tensorboard-logdir`` / ``--wandb-project`` add the board sink.  The emitted
lines are byte-compatible with the reference's bars (``unicore/logging/progress_bar.py:138-300`` there: json payload with
``epoch`` / ``update`` first, ``"prefix:  i / n k=v, ..."`` and ``"prefix | k v | ..."``) because log scrapers depend on
them; the class names of the reference (``JsonProgressBar`` ...) remain available as pre-configured reporters.
"""
import atexit
import json
import logging
import os
import sys
from collections import OrderedDict
from contextlib import contextmanager
from numbers import Number
from typing import Optional

import torch

from .meters import AverageMeter, StopwatchMeter, TimeMeter

logger = logging.getLogger(__name__)


# ---- rendering ------------------------------------------------------------------------------------------------------
def format_stat(stat):
    """Scalar / meter / tensor -> what goes into a log line."""
    if isinstance(stat, Number):
        return "{:g}".format(stat)
    if isinstance(stat, AverageMeter):
        return "{:.3f}".format(stat.avg)
    if isinstance(stat, TimeMeter):
        return "{:g}".format(round(stat.avg))
    if isinstance(stat, StopwatchMeter):
        return "{:g}".format(round(stat.sum))
    return stat.tolist() if torch.is_tensor(stat) else stat


def _as_text(stats):
    return OrderedDict((name, str(format_stat(value)).strip()) for name, value in stats.items())


@contextmanager
def rename_logger(log, new_name):
    """Attribute the lines written inside the block to ``new_name`` (train_inner, valid, ...)."""
    original = log.name
    log.name = new_name if new_name is not None else original
    try:
        yield log
    finally:
        log.name = original


class _Where:
    """Position of the run when a sink is called."""

    __slots__ = ("epoch", "prefix", "index", "total", "tag", "step")

    def __init__(self, epoch, prefix, index, total, tag, step):
        self.epoch, self.prefix, self.index, self.total, self.tag, self.step = epoch, prefix, index, total, tag, step

    @property
    def fractional_epoch(self):
        if self.epoch is None or not self.total or self.index is None:
            return None
        return self.epoch - 1 + (self.index + 1) / float(self.total)


# ---- sinks ----------------------------------------------------------------------------------------------------------
class Sink:
    def wrap(self, iterable, prefix):
        """Chance to substitute the iterable (tqdm does)."""
        return iterable

    def update(self, where, stats):
        pass

    def summary(self, where, stats):
        pass

    def configure(self, config):
        pass


class ConsoleSink(Sink):
    def __init__(self, style, log_interval):
        self.style = style
        self.log_interval = log_interval

    def _due(self, where):
        step = where.step or where.index or 0
        return step > 0 and self.log_interval is not None and step % self.log_interval == 0

    @staticmethod
    def _json_line(where, stats, with_update):
        payload = OrderedDict()
        if where.epoch is not None:
            payload["epoch"] = where.epoch
        if with_update and where.fractional_epoch is not None:
            payload["update"] = round(where.fractional_epoch, 3)
        for name, value in stats.items():
            payload[name] = format_stat(value)
        return json.dumps(payload)

    def update(self, where, stats):
        if not self._due(where):
            return
        if self.style == "json":
            line = self._json_line(where, stats, with_update=True)
        else:
            pairs = ", ".join("{}={}".format(k, v) for k, v in _as_text(stats).items())
            line = "{}:  {:5d} / {:d} {}".format(where.prefix, where.index + 1, where.total, pairs)
        with rename_logger(logger, where.tag):
            logger.info(line)

    def summary(self, where, stats):
        if self.style == "json":
            if where.tag is not None:
                stats = OrderedDict((where.tag + "_" + k, v) for k, v in stats.items())
            line = self._json_line(where, stats, with_update=False)
        else:
            line = "{} | {}".format(where.prefix, " | ".join("{} {}".format(k, v) for k, v in _as_text(stats).items()))
        with rename_logger(logger, where.tag):
            logger.info(line)


class TqdmSink(ConsoleSink):
    def __init__(self):
        super().__init__("simple", None)
        self.bar = None

    def wrap(self, iterable, prefix):
        from tqdm import tqdm

        self.bar = tqdm(iterable, prefix, leave=False, disable=(logger.getEffectiveLevel() > logging.INFO))
        return self.bar

    def update(self, where, stats):
        self.bar.set_postfix(_as_text(stats), refresh=False)


_open_writers = {}


@atexit.register
def _close_writers():
    for writer in _open_writers.values():
        writer.close()


def _summary_writer_class():
    for module in ("torch.utils.tensorboard", "tensorboardX"):
        try:
            return getattr(__import__(module, fromlist=["SummaryWriter"]), "SummaryWriter")
        except Exception:  # noqa: BLE001  (missing package, broken protobuf, ...)
            continue
    return None


class BoardSink(Sink):
    """Scalars of every ``log`` / ``print`` -> TensorBoard (directory per tag) and optionally W&B."""

    def __init__(self, logdir, wandb_project=None, wandb_name=None, args=None):
        self.logdir = logdir
        self.writer_cls = _summary_writer_class()
        if self.writer_cls is None:
            logger.warning("tensorboard not found, please install with: pip install tensorboard")
        self.wandb = None
        if wandb_project:
            try:
                import wandb

                if wandb.run is None:
                    wandb.init(project=wandb_project, name=wandb_name or None, config=vars(args) if args else None)
                self.wandb = wandb
            except Exception:  # noqa: BLE001
                logger.warning("wandb not available; --wandb-project ignored")

    def _writer(self, tag):
        if self.writer_cls is None:
            return None
        if tag not in _open_writers:
            _open_writers[tag] = self.writer_cls(os.path.join(self.logdir, tag))
            _open_writers[tag].add_text("sys.argv", " ".join(sys.argv))
        return _open_writers[tag]

    @staticmethod
    def _scalars(stats):
        out = {}
        for name, value in stats.items():
            if name == "num_updates":
                continue
            if isinstance(value, AverageMeter):
                value = value.val
            elif torch.is_tensor(value) and value.numel() == 1:
                value = value.item()
            if isinstance(value, Number):
                out[name] = value
        return out

    def update(self, where, stats):
        step = where.step if where.step is not None else stats.get("num_updates", None)
        scalars = self._scalars(stats)
        writer = self._writer(where.tag or "")
        if writer is not None:
            for name, value in scalars.items():
                writer.add_scalar(name, value, step)
            writer.flush()
        if self.wandb is not None and scalars:
            scope = (where.tag + "/") if where.tag else ""
            self.wandb.log({scope + k: v for k, v in scalars.items()}, step=step)

    summary = update

    def configure(self, config):
        if self.wandb is not None:
            self.wandb.config.update(config, allow_val_change=True)


# ---- the iterable wrapper ---------------------------------------------------------------------------------------------
class ProgressReporter:
    def __init__(self, iterable, epoch=None, prefix=None, sinks=()):
        self.iterable = iterable
        self.epoch = epoch
        self.prefix = " | ".join(
            part for part in ("epoch {:03d}".format(epoch) if epoch is not None else None, prefix) if part is not None
        )
        self.sinks = list(sinks)
        self.n = getattr(iterable, "n", 0)   # resumed iterators start in the middle of an epoch
        self.i = None
        self.size = None
        self._source = iterable
        for sink in self.sinks:
            self._source = sink.wrap(self._source, self.prefix)

    def __len__(self):
        return len(self.iterable)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        self.size = len(self.iterable)
        for self.i, item in enumerate(self._source, start=self.n):
            yield item

    def _where(self, tag, step):
        return _Where(self.epoch, self.prefix, self.i, self.size, tag, step)

    def log(self, stats, tag=None, step=None):
        """Statistics of the update that just finished."""
        where = self._where(tag, step)
        for sink in reversed(self.sinks):   # boards first, console last (order of the reference's wrapper)
            sink.update(where, stats)

    def print(self, stats, tag=None, step=None):
        """End-of-epoch / end-of-validation summary."""
        where = self._where(tag, step)
        for sink in reversed(self.sinks):
            sink.summary(where, stats)

    def update_config(self, config):
        for sink in self.sinks:
            sink.configure(config)


def _console_sinks(fmt, log_interval):
    if fmt == "json" or fmt == "simple":
        return [ConsoleSink(fmt, log_interval)]
    if fmt == "tqdm":
        return [TqdmSink()]
    if fmt == "none":
        return []
    raise ValueError("Unknown log format: {}".format(fmt))


def progress_bar(
    iterator,
    log_format: Optional[str] = None,
    log_interval: int = 100,
    epoch: Optional[int] = None,
    prefix: Optional[str] = None,
    tensorboard_logdir: Optional[str] = None,
    wandb_project: Optional[str] = None,
    wandb_name: Optional[str] = None,
    default_log_format: str = "tqdm",
    args=None,
):
    fmt = log_format if log_format is not None else default_log_format
    if fmt == "tqdm" and not sys.stderr.isatty():
        fmt = "simple"
    sinks = _console_sinks(fmt, log_interval)
    if tensorboard_logdir:
        sinks.append(BoardSink(tensorboard_logdir, wandb_project, wandb_name, args))
    return ProgressReporter(iterator, epoch, prefix, sinks)


# ---- reference class names (pre-configured reporters) ----------------------------------------------------------------
BaseProgressBar = ProgressReporter


class JsonProgressBar(ProgressReporter):
    def __init__(self, iterable, epoch=None, prefix=None, log_interval=1000):
        super().__init__(iterable, epoch, prefix, _console_sinks("json", log_interval))


class SimpleProgressBar(ProgressReporter):
    def __init__(self, iterable, epoch=None, prefix=None, log_interval=1000):
        super().__init__(iterable, epoch, prefix, _console_sinks("simple", log_interval))


class NoopProgressBar(ProgressReporter):
    def __init__(self, iterable, epoch=None, prefix=None):
        super().__init__(iterable, epoch, prefix, [])


class TqdmProgressBar(ProgressReporter):
    def __init__(self, iterable, epoch=None, prefix=None):
        super().__init__(iterable, epoch, prefix, _console_sinks("tqdm", None))


class TensorboardProgressBarWrapper(ProgressReporter):
    def __init__(self, wrapped_bar, tensorboard_logdir, wandb_project=None, wandb_name=None, args=None):
        super().__init__(wrapped_bar.iterable, wrapped_bar.epoch, None,
                         wrapped_bar.sinks + [BoardSink(tensorboard_logdir, wandb_project, wandb_name, args)])
        self.prefix = wrapped_bar.prefix
        self._source = wrapped_bar._source
