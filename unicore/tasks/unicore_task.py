"""Base class of all tasks: which datasets exist, how they are batched, and what one training / validation step is.

The public surface (method names, signatures, return conventions) is the reference's ``unicore/tasks/unicore_task.py``
(``UnicoreTask:45``, ``get_batch_iterator:138``, ``train_step:253``, ``valid_step:286``, ``optimizer_step:292``,
``reduce_metrics:308``, ``state_dict:320``) because user plug-ins subclass it.  Internals differ: task state lives in a
per-instance ``StatefulContainer`` (the reference shares class-level dicts between all tasks - SURVEY D12), iterator
reuse is an explicit small cache, and the batch iterator can page-lock prefetched batches (``--pin-memory``).
"""
import logging
import warnings
from argparse import Namespace
from typing import Any, Callable, Dict, List

import torch

from unicore import metrics, utils
from unicore.data import UnicoreDataset, data_utils, iterators

logger = logging.getLogger(__name__)


class StatefulContainer(object):
    """Checkpointed task state, attribute style: ``state.foo`` is built by its registered factory on first use and is
    restored from ``task_state`` when a checkpoint is loaded."""

    _RESERVED = ("_values", "_makers")

    def __init__(self):
        object.__setattr__(self, "_values", {})
        object.__setattr__(self, "_makers", {})

    def add_factory(self, name: str, factory: Callable[[], Any]) -> None:
        self._makers[name] = factory

    def merge_state_dict(self, state_dict: Dict[str, Any]) -> None:
        self._values.update(state_dict)

    @property
    def state_dict(self) -> Dict[str, Any]:
        return self._values

    def __getattr__(self, name):
        if name in StatefulContainer._RESERVED:
            raise AttributeError(name)
        values = object.__getattribute__(self, "_values")
        if name in values:
            return values[name]
        makers = object.__getattribute__(self, "_makers")
        if name not in makers:
            raise AttributeError("Task state has no factory for attribute {}".format(name))
        values[name] = makers[name]()
        return values[name]


class UnicoreTask(object):
    # ---- construction / registration -------------------------------------------------------------------------
    @classmethod
    def add_args(cls, parser):
        """Add task-specific arguments to the parser."""

    @classmethod
    def setup_task(cls, args: Namespace, **kwargs):
        return cls(args, **kwargs)

    def __init__(self, args: Namespace, **kwargs):
        self.args = args
        self.datasets: Dict[str, UnicoreDataset] = {}
        self.dataset_to_epoch_iter: Dict[UnicoreDataset, Any] = {}
        self.state = StatefulContainer()

    def build_model(self, args: Namespace):
        from unicore import models

        return models.build_model(args, self)

    def build_loss(self, args: Namespace):
        from unicore import losses

        return losses.build_loss(args, self)

    # ---- datasets ---------------------------------------------------------------------------------------------
    def load_dataset(self, split: str, combine: bool = False, **kwargs):
        """Subclasses fill ``self.datasets[split]``."""
        raise NotImplementedError

    def has_sharded_data(self, split):
        return False

    def dataset(self, split):
        try:
            found = self.datasets[split]
        except KeyError:
            raise KeyError("Dataset not loaded: " + split) from None
        if not isinstance(found, UnicoreDataset):
            raise TypeError("Datasets are expected to be of type UnicoreDataset")
        return found

    def build_dataset_for_inference(self, src_tokens: List[torch.Tensor], src_lengths: List[int], **kwargs):
        raise NotImplementedError

    def disable_shuffling(self) -> bool:
        return False

    def can_reuse_epoch_itr(self, dataset):
        return getattr(dataset, "can_reuse_epoch_itr_across_epochs", False)

    def get_batch_iterator(self, dataset, batch_size=None, ignore_invalid_inputs=False, required_batch_size_multiple=1,
                           seed=1, num_shards=1, shard_id=0, num_workers=0, epoch=1, data_buffer_size=0,
                           disable_iterator_cache=False):
        """The sharded, resumable ``EpochBatchIterator`` over ``dataset`` (cached per dataset when the dataset says its
        batches do not depend on the epoch)."""
        cacheable = self.can_reuse_epoch_itr(dataset) and not disable_iterator_cache
        cached = self.dataset_to_epoch_iter.get(dataset) if cacheable else None
        if cached is not None:
            logger.info("reusing EpochBatchIterator for epoch {}".format(epoch))
            return cached
        if not isinstance(dataset, UnicoreDataset):
            raise TypeError("dataset must be a UnicoreDataset")
        logger.info("get EpochBatchIterator for epoch {}".format(epoch))
        dataset.set_epoch(epoch)
        with data_utils.numpy_seed(seed):  # the dataset's own ordering (e.g. length sorting with random ties)
            order = dataset.ordered_indices()
        sampler = dataset.batch_by_size(order, batch_size=batch_size,
                                        required_batch_size_multiple=required_batch_size_multiple)
        itr = iterators.EpochBatchIterator(
            dataset=dataset, collate_fn=dataset.collater, batch_sampler=sampler, seed=seed, num_shards=num_shards,
            shard_id=shard_id, num_workers=num_workers, epoch=epoch, buffer_size=data_buffer_size,
            disable_shuffling=self.disable_shuffling(), pin_memory=getattr(self.args, "pin_memory", False),
        )
        if cacheable:
            self.dataset_to_epoch_iter[dataset] = itr
        return itr

    # ---- the step contract ----------------------------------------------------------------------------------------
    def train_step(self, sample, model, loss, optimizer, update_num, ignore_grad=False):
        """Forward + backward of one micro-batch -> ``(loss, sample_size, logging_output)``.

        ``ignore_grad`` marks the dummy batch of a rank whose shard is exhausted: the backward still runs (the ranks must
        issue identical collectives) but on a loss of zero."""
        if not model.training:  # (Module.train() walks the whole tree; the trainer has usually done it already)
            model.train()
        model.set_num_updates(update_num)
        with torch.autograd.profiler.record_function("forward"):
            value, sample_size, logging_output = loss(model, sample)
        if ignore_grad:
            value = value * 0
        # single process: the statistics need no cross-rank reduction, so their copy to the host can start right here,
        # behind the forward pass (``utils.stage_logging_output``); with several ranks they travel in the optimizer tail
        utils.stage_logging_output(logging_output, enabled=getattr(self.args, "distributed_world_size", 1) == 1)
        with torch.autograd.profiler.record_function("backward"):
            optimizer.backward(value)
        return value, sample_size, logging_output

    def valid_step(self, sample, model, loss, test=False):
        model.eval()
        with torch.no_grad():
            return loss(model, sample)

    def optimizer_step(self, optimizer, model, update_num):
        optimizer.step()

    def begin_epoch(self, epoch, model):
        """Called at the start of every training epoch."""

    def begin_valid_epoch(self, epoch, model):
        """Called at the start of every validation pass."""

    # ---- logging ------------------------------------------------------------------------------------------------
    @staticmethod
    def logging_outputs_can_be_summed(loss, is_train) -> bool:
        """True: the per-rank logging outputs are scalars that may simply be added (they travel as one small vector);
        False: they are gathered as pickled objects."""
        return loss.logging_outputs_can_be_summed(is_train)

    def reduce_metrics(self, logging_outputs, loss, split="train"):
        """Feed the (already cross-rank) logging outputs into the metrics system."""
        if any("bsz" in log for log in logging_outputs):
            metrics.log_scalar("bsz", sum(log.get("bsz", 0) for log in logging_outputs), priority=190, round=1)
        else:
            warnings.warn("bsz not found in Loss logging outputs, cannot log bsz")
        type(loss).reduce_metrics(logging_outputs, split)

    # ---- checkpointed state -------------------------------------------------------------------------------------
    def state_dict(self):
        return {} if self.state is None else dict(self.state.state_dict)

    def load_state_dict(self, state_dict: Dict[str, Any]):
        if self.state is not None:
            self.state.merge_state_dict(state_dict)
