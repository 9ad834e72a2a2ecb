"""Collation, deterministic seeding and fixed-count batching helpers.

Parity: reference ``unicore/data/data_utils.py`` (``collate_tokens:17``, ``collate_tokens_2d:40``,
``collate_dict:63``, ``str_hash:76``, ``numpy_seed:84``, ``batch_by_size:107``).  The seed folding
in ``numpy_seed`` is a *numerical contract* (it fixes data order and masking noise), so the same
``hash(tuple) % 1e8`` derivation is used to keep runs comparable with the reference.
"""
import contextlib
from typing import Optional, Sequence

import numpy as np
import torch


def _padded_length(longest: int, pad_to_length: Optional[int], pad_to_multiple: int) -> int:
    size = longest if pad_to_length is None else max(longest, pad_to_length)
    if pad_to_multiple > 1:
        size = -(-size // pad_to_multiple) * pad_to_multiple
    return size


def collate_tokens(values: Sequence[torch.Tensor], pad_idx, left_pad=False, pad_to_length=None, pad_to_multiple=1):
    """Stack 1-D tensors of different lengths into a padded ``[B, T]`` tensor."""
    size = _padded_length(max(v.size(0) for v in values), pad_to_length, pad_to_multiple)
    out = values[0].new_full((len(values), size), pad_idx)
    for row, v in zip(out, values):
        n = v.size(0)
        (row[size - n:] if left_pad else row[:n]).copy_(v)
    return out


def collate_tokens_2d(values: Sequence[torch.Tensor], pad_idx, left_pad=False, pad_to_length=None, pad_to_multiple=1):
    """Stack square ``[n_i, n_i]`` pair tensors into a padded ``[B, T, T]`` tensor."""
    size = _padded_length(max(v.size(0) for v in values), pad_to_length, pad_to_multiple)
    out = values[0].new_full((len(values), size, size), pad_idx)
    for plane, v in zip(out, values):
        n = v.size(0)
        (plane[size - n:, size - n:] if left_pad else plane[:n, :n]).copy_(v)
    return out


def collate_dict(values, dim=0):
    """Stack a list of flat dicts of tensors key by key."""
    if len(values) == 0:
        return values
    return {key: torch.stack([v[key] for v in values], dim=dim) for key in values[0].keys()}


def str_hash(text: str) -> int:
    """Small stable (process-independent) string hash."""
    acc = 0
    for ch in text:
        acc = (acc * 281 ^ ord(ch) * 997) & 0xFFFFFFFF
    return acc


def _fold_seed(seed, addl_seeds, key) -> int:
    for s in (seed,) + tuple(addl_seeds):
        if not isinstance(s, (int, np.integer)):
            raise TypeError("seeds must be integers, got {!r}".format(type(s)))
    seed = int(seed)
    if len(addl_seeds) > 0:
        seed = int(hash((seed,) + tuple(int(s) for s in addl_seeds)) % 1e8)
    if key is not None:
        seed = int(hash((seed, str_hash(key))) % 1e8)
    return seed


@contextlib.contextmanager
def numpy_seed(seed, *addl_seeds, key=None):
    """Seed NumPy's global RNG inside the block and restore the previous state afterwards."""
    if seed is None:
        yield
        return
    folded = _fold_seed(seed, addl_seeds, key)
    saved = np.random.get_state()
    np.random.seed(folded)
    try:
        yield
    finally:
        np.random.set_state(saved)


def batch_by_size(indices, batch_size=None, required_batch_size_multiple=1):
    """Split ``indices`` into consecutive batches of a fixed item count.

    The count is ``batch_size`` rounded up to a multiple of ``required_batch_size_multiple``;
    the last batch may be smaller.  (Fixed-count batching is the only mode the reference has.)
    """
    batch_size = 1 if batch_size is None else int(batch_size)
    mult = max(1, int(required_batch_size_multiple))
    step = -(-batch_size // mult) * mult
    if not isinstance(indices, np.ndarray):
        indices = np.fromiter(indices, dtype=np.int64, count=-1)
    if len(indices) == 0:
        return []
    cuts = np.arange(step, len(indices), step)
    return np.split(indices, cuts)
