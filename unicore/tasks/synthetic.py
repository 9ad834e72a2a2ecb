"""Built-in synthetic tasks for benchmarks, smoke tests and CI (no data files, no network).

``synthetic_mlm`` produces BERT-style masked-LM batches of a fixed shape: random token ids
``[B, L]``, ``--mask-prob`` of the positions replaced by ``[MASK]`` with the original id as the
target and padding everywhere else - the exact batch contract of ``MaskedLMLoss``
(``{"net_input": {"src_tokens"}, "target"}``).  It is what ``bench.py`` and the reference-arm
plug-in feed both frameworks with (seq 512, vocab 30,522).
"""
import logging
import os

import numpy as np
import torch

from unicore.data import Dictionary, UnicoreDataset
from unicore.tasks import UnicoreTask, register_task

logger = logging.getLogger(__name__)


def build_synthetic_dictionary(vocab_size: int) -> Dictionary:
    """BERT-like vocabulary of exactly ``vocab_size`` symbols with the usual special-token ids
    (``[PAD]``=0, ``[UNK]``=100, ``[CLS]``=101, ``[SEP]``=102, ``[MASK]``=103)."""
    if vocab_size < 128:
        d = Dictionary()
        for sym in ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"):
            d.add_symbol(sym, is_special=True)
        for i in range(len(d), vocab_size):
            d.add_symbol("tok{}".format(i))
        return d
    specials = {0: "[PAD]", 100: "[UNK]", 101: "[CLS]", 102: "[SEP]", 103: "[MASK]"}
    d = Dictionary()
    for i in range(vocab_size):
        d.add_symbol(specials.get(i, "[unused{}]".format(i) if i < 1000 else "tok{}".format(i)))
    d.specials.add("[MASK]")
    return d


class SyntheticMLMDataset(UnicoreDataset):
    """Deterministic random sentences; item ``i`` depends only on ``(seed, epoch, i)``."""

    def __init__(self, num_samples, seq_len, vocab_size, pad_idx, mask_idx, special_ids, mask_prob=0.15, seed=1):
        super().__init__()
        self.num_samples = num_samples
        self.seq_len = seq_len
        self.vocab_size = vocab_size
        self.pad_idx = pad_idx
        self.mask_idx = mask_idx
        self.mask_prob = mask_prob
        self.seed = seed
        self.epoch = 1
        allowed = np.ones(vocab_size, dtype=bool)
        allowed[list(special_ids)] = False
        self._allowed = np.nonzero(allowed)[0]

    def set_epoch(self, epoch):
        self.epoch = epoch

    @property
    def can_reuse_epoch_itr_across_epochs(self):
        return True

    def __len__(self):
        return self.num_samples

    def __getitem__(self, index):
        rng = np.random.RandomState((self.seed * 1000003 + self.epoch * 7919 + index) % (2 ** 31 - 1))
        tokens = self._allowed[rng.randint(0, len(self._allowed), size=self.seq_len)]
        n_mask = max(1, int(round(self.mask_prob * self.seq_len)))
        pos = rng.choice(self.seq_len, n_mask, replace=False)
        target = np.full(self.seq_len, self.pad_idx, dtype=np.int64)
        target[pos] = tokens[pos]
        src = tokens.copy()
        src[pos] = self.mask_idx
        return {"src": torch.from_numpy(src.astype(np.int64)), "tgt": torch.from_numpy(target)}

    def collater(self, samples):
        if len(samples) == 0:
            return {}
        return {
            "net_input": {"src_tokens": torch.stack([s["src"] for s in samples])},
            "target": torch.stack([s["tgt"] for s in samples]),
        }


@register_task("synthetic_mlm")
class SyntheticMLMTask(UnicoreTask):
    @staticmethod
    def add_args(parser):
        parser.add_argument("data", nargs="?", default=None,
                            help="optional directory containing dict.txt (else a synthetic vocabulary is used)")
        parser.add_argument("--synthetic-vocab-size", default=30522, type=int)
        parser.add_argument("--synthetic-seq-len", default=512, type=int)
        parser.add_argument("--synthetic-num-samples", default=4096, type=int)
        parser.add_argument("--mask-prob", default=0.15, type=float)

    def __init__(self, args, dictionary):
        super().__init__(args)
        self.dictionary = dictionary
        self.seed = args.seed
        self.mask_idx = dictionary.add_symbol("[MASK]", is_special=True)

    @classmethod
    def setup_task(cls, args, **kwargs):
        path = os.path.join(args.data, "dict.txt") if getattr(args, "data", None) else None
        if path is not None and os.path.isfile(path):
            dictionary = Dictionary.load(path)
        else:
            dictionary = build_synthetic_dictionary(args.synthetic_vocab_size)
        logger.info("dictionary: {} types".format(len(dictionary)))
        return cls(args, dictionary)

    def load_dataset(self, split, combine=False, **kwargs):
        n = self.args.synthetic_num_samples
        if split != self.args.train_subset:
            n = max(self.args.batch_size or 1, n // 8)
        self.datasets[split] = SyntheticMLMDataset(
            num_samples=n,
            seq_len=min(self.args.synthetic_seq_len, getattr(self.args, "max_seq_len", 1 << 30)),
            vocab_size=len(self.dictionary),
            pad_idx=self.dictionary.pad(),
            mask_idx=self.mask_idx,
            special_ids=self.dictionary.special_index(),
            mask_prob=self.args.mask_prob,
            seed=self.args.seed + (0 if split == self.args.train_subset else 1),
        )
