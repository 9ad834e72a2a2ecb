"""Base task: dataset bookkeeping, batch-iterator construction, the train/valid step contract.

Parity: reference ``unicore/tasks/unicore_task.py`` (``StatefulContainer:20``, ``UnicoreTask:45``
with ``get_batch_iterator:138``, ``train_step:253``, ``valid_step:286``, ``optimizer_step:292``,
``reduce_metrics:308``, ``state_dict:320``).  Differences: the state container keeps per-instance
dicts (the reference shares class-level dicts between all tasks - SURVEY D12) and the batch
iterator can page-lock prefetched batches (``--pin-memory``).
"""
import logging
import warnings
from argparse import Namespace
from typing import Any, Callable, Dict, List

import torch

from unicore import metrics
from unicore.data import UnicoreDataset, data_utils, iterators

logger = logging.getLogger(__name__)


class StatefulContainer(object):
    """Named pieces of task state created lazily by factories and saved in checkpoints."""

    def __init__(self):
        self._state: Dict[str, Any] = {}
        self._factories: Dict[str, Callable[[], Any]] = {}

    def add_factory(self, name, factory: Callable[[], Any]):
        self._factories[name] = factory

    def merge_state_dict(self, state_dict: Dict[str, Any]):
        self._state.update(state_dict)

    @property
    def state_dict(self) -> Dict[str, Any]:
        return self._state

    def __getattr__(self, name):
        if name in ("_state", "_factories"):
            raise AttributeError(name)
        if name not in self._state:
            if name not in self._factories:
                raise AttributeError("Task state has no factory for attribute {}".format(name))
            self._state[name] = self._factories[name]()
        return self._state[name]


class UnicoreTask(object):
    @classmethod
    def add_args(cls, parser):
        """Add task-specific arguments to the parser."""
        pass

    @staticmethod
    def logging_outputs_can_be_summed(loss, is_train) -> bool:
        """Whether per-rank logging outputs can be reduced with a plain sum (fast path: one tiny
        all-reduce) instead of being gathered as pickled objects."""
        return loss.logging_outputs_can_be_summed(is_train)

    def __init__(self, args: Namespace, **kwargs):
        self.args = args
        self.datasets = dict()
        self.dataset_to_epoch_iter = dict()
        self.state = StatefulContainer()

    @classmethod
    def setup_task(cls, args: Namespace, **kwargs):
        return cls(args, **kwargs)

    def has_sharded_data(self, split):
        return False

    def load_dataset(self, split: str, combine: bool = False, **kwargs):
        """Populate ``self.datasets[split]``."""
        raise NotImplementedError

    def dataset(self, split):
        if split not in self.datasets:
            raise KeyError("Dataset not loaded: " + split)
        ds = self.datasets[split]
        if not isinstance(ds, UnicoreDataset):
            raise TypeError("Datasets are expected to be of type UnicoreDataset")
        return ds

    def can_reuse_epoch_itr(self, dataset):
        return getattr(dataset, "can_reuse_epoch_itr_across_epochs", False)

    def get_batch_iterator(
        self,
        dataset,
        batch_size=None,
        ignore_invalid_inputs=False,
        required_batch_size_multiple=1,
        seed=1,
        num_shards=1,
        shard_id=0,
        num_workers=0,
        epoch=1,
        data_buffer_size=0,
        disable_iterator_cache=False,
    ):
        """Build (or reuse) the sharded, resumable ``EpochBatchIterator`` for ``dataset``."""
        reusable = not disable_iterator_cache and self.can_reuse_epoch_itr(dataset)
        if reusable and dataset in self.dataset_to_epoch_iter:
            logger.info("reusing EpochBatchIterator for epoch {}".format(epoch))
            return self.dataset_to_epoch_iter[dataset]
        logger.info("get EpochBatchIterator for epoch {}".format(epoch))
        if not isinstance(dataset, UnicoreDataset):
            raise TypeError("dataset must be a UnicoreDataset")

        dataset.set_epoch(epoch)
        with data_utils.numpy_seed(seed):
            indices = dataset.ordered_indices()
        batches = dataset.batch_by_size(
            indices, batch_size=batch_size, required_batch_size_multiple=required_batch_size_multiple
        )
        epoch_iter = iterators.EpochBatchIterator(
            dataset=dataset,
            collate_fn=dataset.collater,
            batch_sampler=batches,
            seed=seed,
            num_shards=num_shards,
            shard_id=shard_id,
            num_workers=num_workers,
            epoch=epoch,
            buffer_size=data_buffer_size,
            disable_shuffling=self.disable_shuffling(),
            pin_memory=getattr(self.args, "pin_memory", False),
        )
        if reusable:
            self.dataset_to_epoch_iter[dataset] = epoch_iter
        return epoch_iter

    def build_model(self, args: Namespace):
        from unicore import models

        return models.build_model(args, self)

    def build_loss(self, args: Namespace):
        from unicore import losses

        return losses.build_loss(args, self)

    def train_step(self, sample, model, loss, optimizer, update_num, ignore_grad=False):
        """Forward + backward for one micro-batch.

        Returns ``(loss, sample_size, logging_output)``.  ``ignore_grad`` (dummy batch on an
        exhausted shard) multiplies the loss by 0 so the collective schedule stays identical on
        all ranks.
        """
        if not model.training:  # Module.train() walks the whole tree; the trainer has usually done it
            model.train()
        model.set_num_updates(update_num)
        with torch.autograd.profiler.record_function("forward"):
            loss_val, sample_size, logging_output = loss(model, sample)
        if ignore_grad:
            loss_val = loss_val * 0
        with torch.autograd.profiler.record_function("backward"):
            optimizer.backward(loss_val)
        return loss_val, sample_size, logging_output

    def valid_step(self, sample, model, loss, test=False):
        model.eval()
        with torch.no_grad():
            loss_val, sample_size, logging_output = loss(model, sample)
        return loss_val, sample_size, logging_output

    def optimizer_step(self, optimizer, model, update_num):
        optimizer.step()

    def build_dataset_for_inference(self, src_tokens: List[torch.Tensor], src_lengths: List[int], **kwargs):
        raise NotImplementedError

    def begin_epoch(self, epoch, model):
        """Hook at the start of every epoch."""
        pass

    def begin_valid_epoch(self, epoch, model):
        """Hook at the start of every validation pass."""
        pass

    def reduce_metrics(self, logging_outputs, loss, split="train"):
        """Aggregate logging outputs from data-parallel workers into the metrics system."""
        if not any("bsz" in log for log in logging_outputs):
            warnings.warn("bsz not found in Loss logging outputs, cannot log bsz")
        else:
            bsz = sum(log.get("bsz", 0) for log in logging_outputs)
            metrics.log_scalar("bsz", bsz, priority=190, round=1)
        loss.__class__.reduce_metrics(logging_outputs, split)

    def state_dict(self):
        return dict(self.state.state_dict) if self.state is not None else {}

    def load_state_dict(self, state_dict: Dict[str, Any]):
        if self.state is not None:
            self.state.merge_state_dict(state_dict)

    def disable_shuffling(self) -> bool:
        return False
