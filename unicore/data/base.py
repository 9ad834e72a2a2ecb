"""Dataset base classes (reference ``data/unicore_dataset.py:14-91`` and ``data/base_wrapper_dataset.py:12-61``).

``UnicoreDataset`` is a ``torch.utils.data.Dataset`` plus the hooks the batching / iteration machinery calls:
``collater``, ``ordered_indices``, ``batch_by_size``, epoch notifications and optional prefetching.
``BaseWrapperDataset`` answers every one of those hooks by asking the dataset it wraps, so a concrete wrapper overrides
only what it changes.
"""
import numpy as np
import torch.utils.data

from . import data_utils


class EpochListening:
    """Mixin for objects that are told when an epoch begins."""

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        """True when the set of batches does not depend on the epoch, so that the task may keep one
        ``EpochBatchIterator`` alive across epochs.  A dataset whose ``set_epoch`` changes sizes or order must leave
        this False."""
        return False

    def set_epoch(self, epoch):
        pass


class UnicoreDataset(torch.utils.data.Dataset, EpochListening):
    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def collater(self, samples):
        """List of items -> mini-batch (dict / tensor)."""
        raise NotImplementedError

    def ordered_indices(self):
        """The order in which items are grouped into batches; identity unless overridden."""
        return np.arange(len(self), dtype=np.int64)

    def batch_by_size(self, indices, batch_size=None, required_batch_size_multiple=1):
        """``indices`` cut into batches of ``batch_size`` items."""
        return data_utils.batch_by_size(
            indices, batch_size=batch_size, required_batch_size_multiple=required_batch_size_multiple
        )

    def attr(self, attr: str, index: int):
        return getattr(self, attr, None)

    # -- optional capabilities ------------------------------------------------------------------------------------
    supports_prefetch = False
    supports_fetch_outside_dataloader = True

    def prefetch(self, indices):
        raise NotImplementedError


class BaseWrapperDataset(UnicoreDataset):
    """Every hook is answered by ``self.dataset``."""

    def __init__(self, dataset):
        super().__init__()
        self.dataset = dataset

    def __getitem__(self, index):
        return self.dataset[index]

    def __len__(self):
        return len(self.dataset)

    def collater(self, samples):
        inner = getattr(self.dataset, "collater", None)
        return torch.utils.data.default_collate(samples) if inner is None else inner(samples)

    def ordered_indices(self):
        return self.dataset.ordered_indices()

    def batch_by_size(self, indices, batch_size=None, required_batch_size_multiple=1):
        return self.dataset.batch_by_size(
            indices, batch_size=batch_size, required_batch_size_multiple=required_batch_size_multiple
        )

    def attr(self, attr: str, index: int):
        return self.dataset.attr(attr, index)

    def prefetch(self, indices):
        self.dataset.prefetch(indices)

    def set_epoch(self, epoch):
        super().set_epoch(epoch)
        notify = getattr(self.dataset, "set_epoch", None)
        if notify is not None:
            notify(epoch)

    # capabilities are the wrapped dataset's (with the base class defaults for plain torch datasets)
    supports_prefetch = property(lambda self: getattr(self.dataset, "supports_prefetch", False))
    supports_fetch_outside_dataloader = property(
        lambda self: getattr(self.dataset, "supports_fetch_outside_dataloader", True))
    can_reuse_epoch_itr_across_epochs = property(lambda self: self.dataset.can_reuse_epoch_itr_across_epochs)
