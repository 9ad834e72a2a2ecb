"""BERT-style token masking.

For each item choose ``round_prob(mask_prob * (len-2))`` interior positions; of those, a fraction
``leave_unmasked_prob`` keeps the original token, a fraction ``random_token_prob`` gets a random
non-special token and the rest become ``mask_idx``.  The *target* view holds the original token at
chosen positions and ``pad_idx`` elsewhere.  Noise is a pure function of ``(seed, epoch, index)``
so source and target views agree and runs are reproducible.
Parity: reference ``unicore/data/mask_tokens_dataset.py:16-132`` (same draw order from NumPy's
global RNG, so the masks are identical for identical seeds).
"""
import numpy as np
import torch

from . import data_utils
from .base import BaseWrapperDataset
from .dictionary import Dictionary
from .wrappers import LRUCacheDataset, _Memo


class MaskTokensDataset(BaseWrapperDataset):
    @classmethod
    def apply_mask(cls, dataset, *args, **kwargs):
        """Return ``(source_view, target_view)`` over a shared, cached base dataset."""
        base = LRUCacheDataset(dataset)
        src = LRUCacheDataset(cls(base, *args, **kwargs, return_masked_tokens=False))
        tgt = LRUCacheDataset(cls(base, *args, **kwargs, return_masked_tokens=True))
        return src, tgt

    def __init__(
        self,
        dataset,
        vocab: Dictionary,
        pad_idx: int,
        mask_idx: int,
        return_masked_tokens: bool = False,
        seed: int = 1,
        mask_prob: float = 0.15,
        leave_unmasked_prob: float = 0.1,
        random_token_prob: float = 0.1,
    ):
        super().__init__(dataset)
        if not 0.0 < mask_prob < 1.0:
            raise ValueError("mask_prob must be in (0, 1)")
        if not (0.0 <= random_token_prob <= 1.0 and 0.0 <= leave_unmasked_prob <= 1.0):
            raise ValueError("probabilities must be in [0, 1]")
        if random_token_prob + leave_unmasked_prob > 1.0:
            raise ValueError("random_token_prob + leave_unmasked_prob must be <= 1")
        self.vocab = vocab
        self.pad_idx = pad_idx
        self.mask_idx = mask_idx
        self.return_masked_tokens = return_masked_tokens
        self.seed = seed
        self.mask_prob = mask_prob
        self.leave_unmasked_prob = leave_unmasked_prob
        self.random_token_prob = random_token_prob
        if random_token_prob > 0.0:
            w = np.ones(len(vocab))
            w[vocab.special_index()] = 0
            self.weights = w / w.sum()
        self.epoch = None
        self._memo = _Memo()

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        return True  # only the noise changes with the epoch, not the item sizes

    def set_epoch(self, epoch, **unused):
        super().set_epoch(epoch)
        self.epoch = epoch

    def __getitem__(self, index: int):
        return self._memo.get((self.epoch, index), lambda key: self._noised(*key))

    def _noised(self, epoch, index):
        with data_utils.numpy_seed(self.seed, epoch, index):
            item = self.dataset[index]
            item_np = item.numpy() if torch.is_tensor(item) else np.asarray(item)
            n = len(item_np)
            if n <= 2:
                raise ValueError("cannot mask a sequence of length <= 2")
            if (item_np == self.mask_idx).any():
                raise ValueError("Dataset contains mask_idx (={}), this is not expected!".format(self.mask_idx))

            # draw 1: probabilistic rounding of the number of masked positions
            n_mask = int(self.mask_prob * (n - 2) + np.random.rand())
            # draw 2: which interior positions (never first/last)
            chosen = np.zeros(n, dtype=bool)
            chosen[np.random.choice(n - 2, n_mask, replace=False) + 1] = True

            if self.return_masked_tokens:
                target = np.full(n, self.pad_idx, dtype=item_np.dtype if item_np.dtype.kind == "i" else np.int64)
                target[chosen] = item_np[chosen]
                return torch.from_numpy(target)

            keep = rand = None
            special = self.random_token_prob + self.leave_unmasked_prob
            if special > 0.0:
                # draw 3: which chosen positions are *not* simply replaced by [MASK]
                special_pos = chosen & (np.random.rand(n) < special)
                if self.random_token_prob == 0.0:
                    keep = special_pos
                elif self.leave_unmasked_prob == 0.0:
                    rand = special_pos
                else:
                    # draw 4: split them between "keep original" and "random token"
                    to_keep = np.random.rand(n) < (self.leave_unmasked_prob / special)
                    keep = special_pos & to_keep
                    rand = special_pos & ~to_keep

            to_mask = chosen ^ keep if keep is not None else chosen
            out = np.copy(item_np)
            out[to_mask] = self.mask_idx
            if rand is not None:
                n_rand = int(rand.sum())
                if n_rand > 0:
                    # draw 5: replacement tokens
                    out[rand] = np.random.choice(len(self.vocab), n_rand, p=self.weights)
            return torch.from_numpy(out)
