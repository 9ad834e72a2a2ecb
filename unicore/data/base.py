"""Dataset base classes.

``UnicoreDataset`` = ``torch.utils.data.Dataset`` + the hooks the batching/iteration machinery
needs (``collater``, ``ordered_indices``, ``batch_by_size``, epoch notifications, optional
prefetch).  ``BaseWrapperDataset`` forwards all of them to an inner dataset so that wrappers
override only what they change.  Parity: reference ``data/unicore_dataset.py:14-91`` and
``data/base_wrapper_dataset.py:12-61``.
"""
import numpy as np
import torch.utils.data

from . import data_utils


class EpochListening:
    """Mixin: receives ``set_epoch`` at the start of every epoch."""

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        """True when the set of batches does not depend on the epoch, which lets the task keep
        one ``EpochBatchIterator`` alive across epochs. Datasets that override ``set_epoch`` to
        change sizes/order must leave this False."""
        return False

    def set_epoch(self, epoch):
        pass


class UnicoreDataset(torch.utils.data.Dataset, EpochListening):
    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def collater(self, samples):
        """Merge a list of items into a mini-batch (dict / tensor)."""
        raise NotImplementedError

    def ordered_indices(self):
        """Index order used to form batches (identity by default)."""
        return np.arange(len(self), dtype=np.int64)

    @property
    def supports_prefetch(self):
        return False

    def attr(self, attr: str, index: int):
        return getattr(self, attr, None)

    def prefetch(self, indices):
        raise NotImplementedError

    def batch_by_size(self, indices, batch_size=None, required_batch_size_multiple=1):
        """Chunk ``indices`` into batches of ``batch_size`` items."""
        return data_utils.batch_by_size(
            indices, batch_size=batch_size, required_batch_size_multiple=required_batch_size_multiple
        )

    @property
    def supports_fetch_outside_dataloader(self):
        return True


class BaseWrapperDataset(UnicoreDataset):
    """Delegates every hook to ``self.dataset``."""

    def __init__(self, dataset):
        super().__init__()
        self.dataset = dataset

    def __getitem__(self, index):
        return self.dataset[index]

    def __len__(self):
        return len(self.dataset)

    def collater(self, samples):
        if hasattr(self.dataset, "collater"):
            return self.dataset.collater(samples)
        return torch.utils.data.default_collate(samples)

    def ordered_indices(self):
        return self.dataset.ordered_indices()

    @property
    def supports_prefetch(self):
        return getattr(self.dataset, "supports_prefetch", False)

    def attr(self, attr: str, index: int):
        return self.dataset.attr(attr, index)

    def prefetch(self, indices):
        self.dataset.prefetch(indices)

    def batch_by_size(self, indices, batch_size=None, required_batch_size_multiple=1):
        return self.dataset.batch_by_size(
            indices, batch_size=batch_size, required_batch_size_multiple=required_batch_size_multiple
        )

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        return self.dataset.can_reuse_epoch_itr_across_epochs

    def set_epoch(self, epoch):
        super().set_epoch(epoch)
        if hasattr(self.dataset, "set_epoch"):
            self.dataset.set_epoch(epoch)

    @property
    def supports_fetch_outside_dataloader(self):
        return getattr(self.dataset, "supports_fetch_outside_dataloader", True)
