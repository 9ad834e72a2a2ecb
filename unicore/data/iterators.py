"""Epoch/batch iteration: counting, sharding, grouping, background prefetch, resumable epochs.

Parity: reference ``unicore/data/iterators.py`` (``CountingIterator:28``,
``EpochBatchIterator:151`` incl. ``state_dict:311`` / ``load_state_dict:326`` with position
re-scaling, ``GroupedIterator:406``, ``ShardedIterator:438``, ``BufferedIterator:496``).

B200 design difference: the reference's "GPU-CPU overlapping" loader only overlaps *collation*
(SURVEY D19 - batches stay pageable so the H2D copy is synchronous).  ``BufferedIterator`` here
can page-lock each prefetched batch in its producer thread (``pin_memory=True``) so that the
trainer's ``non_blocking`` copy is a real async DMA, and ``DevicePrefetcher`` stages the next
batch on a side CUDA stream while the current step computes.
"""
import itertools
import logging
import math
import os
import queue
import threading
import time

import numpy as np
import torch

from . import data_utils

logger = logging.getLogger(__name__)

_END = object()  # sentinel placed in the prefetch queue when the source is exhausted


class CountingIterator(object):
    """Iterator wrapper that knows how many items it has yielded (``n``) and its total length."""

    def __init__(self, iterable, start=None, total=None):
        self.iterable = iterable
        self._itr = iter(self)
        self.n = start if start is not None else getattr(iterable, "n", 0)
        self.total = total if total is not None else self.n + len(iterable)

    def __len__(self):
        return self.total

    def __iter__(self):
        for item in self.iterable:
            if self.n >= self.total:
                raise RuntimeError(
                    "Mismatch between actual and expected iterable length. This may be caused by resuming "
                    "training from a checkpoint using a different number of GPUs, in which case you can try "
                    "the --reset-dataloader option. Alternatively you may have a train or validation set that "
                    "is smaller than the number of GPUs. If none of these apply, please report this."
                )
            self.n += 1
            yield item

    def __next__(self):
        return next(self._itr)

    def has_next(self):
        return self.n < len(self)

    def skip(self, num_to_skip):
        next(itertools.islice(self._itr, num_to_skip, num_to_skip), None)
        return self

    def take(self, n):
        """Truncate to at most ``n`` items in total."""
        self.total = min(self.total, n)
        propagated = max(n - self.n, 0)
        if hasattr(self.iterable, "take"):
            self.iterable.take(propagated)
        else:
            self.iterable = itertools.islice(self.iterable, propagated)
        return self


class EpochBatchIterating(object):
    def __len__(self) -> int:
        raise NotImplementedError

    @property
    def next_epoch_idx(self):
        raise NotImplementedError

    def next_epoch_itr(self, shuffle=True, fix_batches_to_gpus=False, set_dataset_epoch=True):
        raise NotImplementedError

    def end_of_epoch(self) -> bool:
        raise NotImplementedError

    @property
    def iterations_in_epoch(self) -> int:
        raise NotImplementedError

    def state_dict(self):
        raise NotImplementedError

    def load_state_dict(self, state_dict):
        raise NotImplementedError

    @property
    def first_batch(self):
        return "DUMMY"


class EpochBatchIterator(EpochBatchIterating):
    """Multi-epoch iterator over a fixed ("frozen") list of batches.

    Each epoch: shuffle the batch list with ``seed + epoch``, give rank ``shard_id`` every
    ``num_shards``-th batch (short shards are padded with empty batches), feed the index lists to
    a ``torch.utils.data.DataLoader`` and wrap the result for prefetch + counting.  The position
    is checkpointable and can be restored into a run with a different world size.
    """

    def __init__(
        self,
        dataset,
        collate_fn,
        batch_sampler,
        seed=1,
        num_shards=1,
        shard_id=0,
        num_workers=0,
        epoch=1,
        buffer_size=0,
        timeout=0,
        disable_shuffling=False,
        pin_memory=False,
    ):
        if not isinstance(dataset, torch.utils.data.Dataset):
            raise TypeError("dataset must be a torch Dataset")
        self.dataset = dataset
        self.collate_fn = collate_fn
        self.batch_sampler = batch_sampler
        self._frozen_batches = None if callable(batch_sampler) else tuple(batch_sampler)
        self.seed = seed
        self.num_shards = num_shards
        self.shard_id = shard_id
        self.num_workers = num_workers
        self.buffer_size = min(buffer_size, 32)  # politeness cap on shared hosts
        self.timeout = timeout
        self.disable_shuffling = disable_shuffling
        self.pin_memory = pin_memory
        self.epoch = max(epoch, 1)  # epochs are 1-based
        self.shuffle = not disable_shuffling
        self._cur_epoch_itr = None
        self._next_epoch_itr = None
        self._supports_prefetch = getattr(dataset, "supports_prefetch", False)

    # -- batches ------------------------------------------------------------------------------
    @property
    def frozen_batches(self):
        if self._frozen_batches is None:
            self._frozen_batches = tuple(self.batch_sampler(self.dataset, self.epoch))
        return self._frozen_batches

    @property
    def first_batch(self):
        if len(self.frozen_batches) == 0:
            raise Exception(
                "The dataset is empty. This could indicate that all elements in the dataset have been skipped. "
                "Try increasing the max number of allowed tokens or using a larger dataset."
            )
        if getattr(self.dataset, "supports_fetch_outside_dataloader", True):
            return self.collate_fn([self.dataset[i] for i in self.frozen_batches[0]])
        return "DUMMY"

    def __len__(self):
        return int(math.ceil(len(self.frozen_batches) / float(self.num_shards)))

    @property
    def n(self):
        return self.iterations_in_epoch

    # -- epoch control ------------------------------------------------------------------------
    @property
    def next_epoch_idx(self):
        if self._next_epoch_itr is not None:
            return self.epoch
        if self._cur_epoch_itr is not None and self.end_of_epoch():
            return self.epoch + 1
        return self.epoch

    def next_epoch_itr(self, shuffle=True, fix_batches_to_gpus=False, set_dataset_epoch=True):
        if self.disable_shuffling:
            shuffle = False
        self.epoch = self.next_epoch_idx
        if set_dataset_epoch and hasattr(self.dataset, "set_epoch"):
            self.dataset.set_epoch(self.epoch)
        if self._next_epoch_itr is not None:  # prepared by load_state_dict
            self._cur_epoch_itr, self._next_epoch_itr = self._next_epoch_itr, None
        else:
            if callable(self.batch_sampler):
                self._frozen_batches = None  # re-sample for the new epoch
            self._cur_epoch_itr = self._get_iterator_for_epoch(
                self.epoch, shuffle, fix_batches_to_gpus=fix_batches_to_gpus
            )
        self.shuffle = shuffle
        return self._cur_epoch_itr

    def end_of_epoch(self) -> bool:
        return not self._cur_epoch_itr.has_next()

    @property
    def iterations_in_epoch(self):
        for itr in (self._cur_epoch_itr, self._next_epoch_itr):
            if itr is not None:
                return itr.n
        return 0

    # -- checkpointing ------------------------------------------------------------------------
    def state_dict(self):
        finished = self.end_of_epoch()
        return {
            "epoch": self.epoch + 1 if finished else self.epoch,
            "iterations_in_epoch": 0 if finished else self.iterations_in_epoch,
            "shuffle": self.shuffle,
            "len": len(self),
        }

    def load_state_dict(self, state_dict):
        self.epoch = state_dict["epoch"]
        pos = state_dict.get("iterations_in_epoch", 0)
        if pos <= 0:
            self._next_epoch_itr = None
            return
        saved_len = state_dict.get("len", None)
        if saved_len is not None and saved_len != len(self):
            rescaled = int(pos * len(self) / saved_len)
            logger.info(
                "Iterator size changed ({} -> {}; different world size or update_freq?). "
                "Position rescaled from {} to {}.".format(saved_len, len(self), pos, rescaled)
            )
            pos = rescaled
        self._next_epoch_itr = self._get_iterator_for_epoch(
            self.epoch, shuffle=state_dict.get("shuffle", True), offset=pos
        )
        if self._next_epoch_itr is None:
            raise RuntimeError(
                "Cannot resume training due to dataloader mismatch. You can relaunch "
                "training with `--reset-dataloader` and it should work."
            )

    # -- construction of one epoch's iterator ----------------------------------------------------
    @staticmethod
    def _shuffled(batches, seed):
        batches = list(batches)
        with data_utils.numpy_seed(seed):
            np.random.shuffle(batches)
        return batches

    def _shard(self, batches):
        return list(ShardedIterator(batches, self.num_shards, self.shard_id, fill_value=[]))

    def _get_iterator_for_epoch(self, epoch, shuffle, fix_batches_to_gpus=False, offset=0):
        batches = self.frozen_batches
        if self._supports_prefetch:
            if shuffle and not fix_batches_to_gpus:
                batches = self._shuffled(batches, self.seed + epoch)
            batches = self._shard(batches)
            self.dataset.prefetch([i for b in batches for i in b])
            if shuffle and fix_batches_to_gpus:
                batches = self._shuffled(batches, self.seed + epoch + self.shard_id)
        else:
            if shuffle:
                batches = self._shuffled(batches, self.seed + epoch)
            batches = self._shard(batches)

        if offset > 0 and offset >= len(batches):
            return None
        if self.num_workers > 0:
            os.environ["PYTHONWARNINGS"] = "ignore:semaphore_tracker:UserWarning"

        itr = torch.utils.data.DataLoader(
            self.dataset,
            collate_fn=self.collate_fn,
            batch_sampler=batches[offset:],
            num_workers=self.num_workers,
            timeout=self.timeout,
        )
        if self.buffer_size > 0:
            itr = BufferedIterator(self.buffer_size, itr, pin_memory=self.pin_memory)
        return CountingIterator(itr, start=offset)


class GroupedIterator(CountingIterator):
    """Yield lists of ``chunk_size`` consecutive items (the micro-batches of one update)."""

    def __init__(self, iterable, chunk_size):
        def chunks():
            it = iter(iterable)
            while True:
                group = list(itertools.islice(it, chunk_size))
                if not group:
                    return
                yield group

        super().__init__(
            chunks(),
            start=int(math.ceil(getattr(iterable, "n", 0) / float(chunk_size))),
            total=int(math.ceil(len(iterable) / float(chunk_size))),
        )
        self.chunk_size = chunk_size


class ShardedIterator(CountingIterator):
    """Every ``num_shards``-th item starting at ``shard_id``; all shards have equal length
    (``ceil(len/num_shards)``), short ones are completed with ``fill_value``."""

    def __init__(self, iterable, num_shards, shard_id, fill_value=None):
        if not 0 <= shard_id < num_shards:
            raise ValueError("shard_id must be between 0 and num_shards")
        sharded_len = int(math.ceil(len(iterable) / float(num_shards)))
        picked = itertools.islice(iterable, shard_id, len(iterable), num_shards)
        padded = itertools.chain(picked, itertools.repeat(fill_value))
        super().__init__(
            itertools.islice(padded, sharded_len),
            start=int(math.ceil(getattr(iterable, "n", 0) / float(num_shards))),
            total=sharded_len,
        )


class BackgroundConsumer(threading.Thread):
    """Daemon thread: pulls items from ``source`` into ``out_queue`` (optionally pinning them)."""

    def __init__(self, out_queue, source, max_len, pin_memory=False):
        super().__init__(daemon=True)
        self._queue = out_queue
        self._source = source
        self._max_len = max_len
        self._pin = pin_memory and torch.cuda.is_available()
        self.count = 0

    def run(self):
        from unicore import utils

        try:
            for item in self._source:
                if self._pin:
                    item = utils.pin_sample(item)
                self._queue.put(item)
                self.count += 1
                if self._max_len is not None and self.count >= self._max_len:
                    break
            self._queue.put(_END)
        except Exception as exc:  # noqa: BLE001 - hand the failure to the consumer thread
            self._queue.put(exc)


class BufferedIterator(object):
    """Prefetch up to ``size`` collated batches in a background thread."""

    def __init__(self, size, iterable, pin_memory=False):
        self._queue = queue.Queue(size)
        self._iterable = iterable
        self._consumer = None
        self._pin_memory = pin_memory
        self.start_time = time.time()
        self.warning_time = None
        self.total = len(iterable)

    def _start(self):
        self._consumer = BackgroundConsumer(self._queue, self._iterable, self.total, self._pin_memory)
        self._consumer.start()

    def __iter__(self):
        return self

    def __len__(self):
        return self.total

    def take(self, n):
        self.total = min(self.total, n)
        if hasattr(self._iterable, "take"):
            self._iterable.take(n)
        return self

    def __next__(self):
        if self._consumer is None:
            self._start()
        # tell the user (at most every 15 min) when the input pipeline is the bottleneck
        if self._queue.qsize() < min(2, max(1, self._queue.maxsize // 2)):
            now = time.time()
            if now - self.start_time > 5 * 60 and (self.warning_time is None or now - self.warning_time > 15 * 60):
                logger.debug(
                    "Data loading buffer is empty or nearly empty. This may indicate a data loading "
                    "bottleneck, and increasing the number of workers (--num-workers) may help."
                )
                self.warning_time = now
        item = self._queue.get(True)
        if isinstance(item, Exception):
            raise item
        if item is _END:
            raise StopIteration()
        return item


class DevicePrefetcher(object):
    """Stage batch *i+1* on the device (side stream) while batch *i* is being consumed.

    Wraps any iterator of (nested) CPU samples; yields samples whose tensors already live on
    ``device``.  With pinned sources the copy is a true async DMA; the consumer stream waits on
    a CUDA event, never on the host.
    """

    def __init__(self, iterable, device=None):
        self._iterable = iterable
        self._device = device if device is not None else torch.cuda.current_device()
        self._stream = torch.cuda.Stream(device=self._device)

    def __len__(self):
        return len(self._iterable)

    def _stage(self, sample):
        from unicore import utils

        with torch.cuda.stream(self._stream):
            moved = utils.move_to_cuda(sample, device=self._device)
        event = torch.cuda.Event()
        event.record(self._stream)
        return moved, event

    def __iter__(self):
        from unicore import utils

        it = iter(self._iterable)
        try:
            staged = self._stage(next(it))
        except StopIteration:
            return
        for nxt in it:
            ready, event = staged
            staged = self._stage(nxt)
            torch.cuda.current_stream().wait_event(event)
            utils.apply_to_sample(lambda t: t.record_stream(torch.cuda.current_stream()), ready)
            yield ready
        ready, event = staged
        torch.cuda.current_stream().wait_event(event)
        utils.apply_to_sample(lambda t: t.record_stream(torch.cuda.current_stream()), ready)
        yield ready
