"""Log sinks: wrap an iterable and print/forward statistics.

Formats ``json`` / ``none`` / ``simple`` / ``tqdm`` plus a TensorBoard (and optional W&B)
forwarding wrapper; line formats match the reference so that log scrapers keep working
(reference ``unicore/logging/progress_bar.py``: factory ``:29``, ``JsonProgressBar:138``,
``NoopProgressBar:189``, ``SimpleProgressBar:208``, ``TqdmProgressBar:243``,
``TensorboardProgressBarWrapper:302``).  TensorBoard / tensorboardX / wandb are imported lazily.
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
    builders = {
        "json": lambda: JsonProgressBar(iterator, epoch, prefix, log_interval),
        "none": lambda: NoopProgressBar(iterator, epoch, prefix),
        "simple": lambda: SimpleProgressBar(iterator, epoch, prefix, log_interval),
        "tqdm": lambda: TqdmProgressBar(iterator, epoch, prefix),
    }
    if fmt not in builders:
        raise ValueError("Unknown log format: {}".format(fmt))
    bar = builders[fmt]()
    if tensorboard_logdir:
        bar = TensorboardProgressBarWrapper(bar, tensorboard_logdir, wandb_project, wandb_name, args)
    return bar


def format_stat(stat):
    if isinstance(stat, Number):
        return "{:g}".format(stat)
    if isinstance(stat, AverageMeter):
        return "{:.3f}".format(stat.avg)
    if isinstance(stat, TimeMeter):
        return "{:g}".format(round(stat.avg))
    if isinstance(stat, StopwatchMeter):
        return "{:g}".format(round(stat.sum))
    if torch.is_tensor(stat):
        return stat.tolist()
    return stat


class BaseProgressBar(object):
    """Iterable wrapper with ``log`` (mid-epoch) and ``print`` (end-of-epoch) hooks."""

    def __init__(self, iterable, epoch=None, prefix=None):
        self.iterable = iterable
        self.n = getattr(iterable, "n", 0)
        self.epoch = epoch
        parts = []
        if epoch is not None:
            parts.append("epoch {:03d}".format(epoch))
        if prefix is not None:
            parts.append(prefix)
        self.prefix = " | ".join(parts)

    def __len__(self):
        return len(self.iterable)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __iter__(self):
        raise NotImplementedError

    def log(self, stats, tag=None, step=None):
        raise NotImplementedError

    def print(self, stats, tag=None, step=None):
        raise NotImplementedError

    def update_config(self, config):
        """Hook for sinks that record the run configuration (console bars ignore it)."""

    def _str_commas(self, stats):
        return ", ".join("{}={}".format(k, v.strip()) for k, v in stats.items())

    def _str_pipes(self, stats):
        return " | ".join("{} {}".format(k, v.strip()) for k, v in stats.items())

    def _format_stats(self, stats):
        return OrderedDict((k, str(format_stat(v))) for k, v in stats.items())


@contextmanager
def rename_logger(log, new_name):
    """Temporarily rename a logger so lines are attributed to the tag (train_inner/valid...)."""
    old = log.name
    if new_name is not None:
        log.name = new_name
    try:
        yield log
    finally:
        log.name = old


class JsonProgressBar(BaseProgressBar):
    def __init__(self, iterable, epoch=None, prefix=None, log_interval=1000):
        super().__init__(iterable, epoch, prefix)
        self.log_interval = log_interval
        self.i = None
        self.size = None

    def __iter__(self):
        self.size = len(self.iterable)
        for i, obj in enumerate(self.iterable, start=self.n):
            self.i = i
            yield obj

    def _payload(self, stats, epoch=None, update=None):
        out = OrderedDict()
        if epoch is not None:
            out["epoch"] = epoch
        if update is not None:
            out["update"] = round(update, 3)
        for k, v in stats.items():
            out[k] = format_stat(v)
        return out

    def log(self, stats, tag=None, step=None):
        step = step or self.i or 0
        if step > 0 and self.log_interval is not None and step % self.log_interval == 0:
            update = (
                self.epoch - 1 + (self.i + 1) / float(self.size)
                if self.epoch is not None and self.size
                else None
            )
            with rename_logger(logger, tag):
                logger.info(json.dumps(self._payload(stats, epoch=self.epoch, update=update)))

    def print(self, stats, tag=None, step=None):
        self.stats = stats
        if tag is not None:
            self.stats = OrderedDict((tag + "_" + k, v) for k, v in self.stats.items())
        with rename_logger(logger, tag):
            logger.info(json.dumps(self._payload(self.stats, epoch=self.epoch)))


class NoopProgressBar(BaseProgressBar):
    def __iter__(self):
        for obj in self.iterable:
            yield obj

    def log(self, stats, tag=None, step=None):
        pass

    def print(self, stats, tag=None, step=None):
        pass


class SimpleProgressBar(BaseProgressBar):
    """One plain log line every ``log_interval`` steps."""

    def __init__(self, iterable, epoch=None, prefix=None, log_interval=1000):
        super().__init__(iterable, epoch, prefix)
        self.log_interval = log_interval
        self.i = None
        self.size = None

    def __iter__(self):
        self.size = len(self.iterable)
        for i, obj in enumerate(self.iterable, start=self.n):
            self.i = i
            yield obj

    def log(self, stats, tag=None, step=None):
        step = step or self.i or 0
        if step > 0 and self.log_interval is not None and step % self.log_interval == 0:
            text = self._str_commas(self._format_stats(stats))
            with rename_logger(logger, tag):
                logger.info("{}:  {:5d} / {:d} {}".format(self.prefix, self.i + 1, self.size, text))

    def print(self, stats, tag=None, step=None):
        text = self._str_pipes(self._format_stats(stats))
        with rename_logger(logger, tag):
            logger.info("{} | {}".format(self.prefix, text))


class TqdmProgressBar(BaseProgressBar):
    def __init__(self, iterable, epoch=None, prefix=None):
        super().__init__(iterable, epoch, prefix)
        from tqdm import tqdm

        self.tqdm = tqdm(
            iterable,
            self.prefix,
            leave=False,
            disable=(logger.getEffectiveLevel() > logging.INFO),
        )

    def __iter__(self):
        return iter(self.tqdm)

    def log(self, stats, tag=None, step=None):
        self.tqdm.set_postfix(self._format_stats(stats), refresh=False)

    def print(self, stats, tag=None, step=None):
        text = self._str_pipes(self._format_stats(stats))
        with rename_logger(logger, tag):
            logger.info("{} | {}".format(self.prefix, text))


_writers = {}


def _close_writers():
    for w in _writers.values():
        w.close()


atexit.register(_close_writers)


def _summary_writer_class():
    try:
        from torch.utils.tensorboard import SummaryWriter

        return SummaryWriter
    except Exception:  # noqa: BLE001
        try:
            from tensorboardX import SummaryWriter

            return SummaryWriter
        except Exception:  # noqa: BLE001
            return None


class TensorboardProgressBarWrapper(BaseProgressBar):
    """Forward every ``log``/``print`` to TensorBoard (one writer per tag) and optionally W&B."""

    def __init__(self, wrapped_bar, tensorboard_logdir, wandb_project=None, wandb_name=None, args=None):
        self.wrapped_bar = wrapped_bar
        self.tensorboard_logdir = tensorboard_logdir
        self._writer_cls = _summary_writer_class()
        if self._writer_cls is None:
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

    def _writer(self, key):
        if self._writer_cls is None:
            return None
        if key not in _writers:
            _writers[key] = self._writer_cls(os.path.join(self.tensorboard_logdir, key))
            _writers[key].add_text("sys.argv", " ".join(sys.argv))
        return _writers[key]

    def __len__(self):
        return len(self.wrapped_bar)

    def __iter__(self):
        return iter(self.wrapped_bar)

    def log(self, stats, tag=None, step=None):
        self._forward(stats, tag, step)
        self.wrapped_bar.log(stats, tag=tag, step=step)

    def print(self, stats, tag=None, step=None):
        self._forward(stats, tag, step)
        self.wrapped_bar.print(stats, tag=tag, step=step)

    def update_config(self, config):
        if self.wandb is not None:
            self.wandb.config.update(config, allow_val_change=True)
        self.wrapped_bar.update_config(config)

    def _forward(self, stats, tag=None, step=None):
        writer = self._writer(tag or "")
        if step is None:
            step = stats.get("num_updates", None)
        scalars = {}
        for key in stats.keys() - {"num_updates"}:
            val = stats[key]
            if isinstance(val, AverageMeter):
                val = val.val
            elif torch.is_tensor(val) and val.numel() == 1:
                val = val.item()
            if isinstance(val, Number):
                scalars[key] = val
        if writer is not None:
            for key, val in scalars.items():
                writer.add_scalar(key, val, step)
            writer.flush()
        if self.wandb is not None and scalars:
            prefix = (tag + "/") if tag else ""
            self.wandb.log({prefix + k: v for k, v in scalars.items()}, step=step)
