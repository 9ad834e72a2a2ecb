"""Symbol <-> index vocabulary with BERT-style special tokens.

Parity: reference ``unicore/data/dictionary.py:12-148`` (specials ``[CLS] [PAD] [SEP] [UNK]``,
``add_symbol`` / ``index`` / ``vec_index`` / ``special_index``, text format ``<symbol> [count]``
with an optional ``#overwrite`` flag).
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)


class Dictionary:
    def __init__(self, *, bos="[CLS]", pad="[PAD]", eos="[SEP]", unk="[UNK]", extra_special_symbols=None):
        self.bos_word, self.pad_word, self.eos_word, self.unk_word = bos, pad, eos, unk
        self.symbols = []
        self.count = []
        self.indices = {}
        self.specials = {bos, unk, pad, eos}
        if extra_special_symbols:
            self.specials.update(extra_special_symbols)

    # -- container protocol ------------------------------------------------------------------
    def __eq__(self, other):
        return isinstance(other, Dictionary) and self.indices == other.indices

    def __getitem__(self, idx):
        return self.symbols[idx] if 0 <= idx < len(self.symbols) else self.unk_word

    def __len__(self):
        return len(self.symbols)

    def __contains__(self, sym):
        return sym in self.indices

    # -- lookup ------------------------------------------------------------------------------
    def index(self, sym):
        if not isinstance(sym, str):
            raise TypeError("symbols are strings")
        hit = self.indices.get(sym)
        return hit if hit is not None else self.indices[self.unk_word]

    def vec_index(self, a):
        return np.vectorize(self.index)(a)

    def special_index(self):
        return [self.index(s) for s in self.specials]

    def bos(self):
        return self.index(self.bos_word)

    def pad(self):
        return self.index(self.pad_word)

    def eos(self):
        return self.index(self.eos_word)

    def unk(self):
        return self.index(self.unk_word)

    # -- mutation ----------------------------------------------------------------------------
    def add_symbol(self, word, n=1, overwrite=False, is_special=False):
        """Add ``word`` (or bump its count) and return its index."""
        if is_special:
            self.specials.add(word)
        if word in self.indices and not overwrite:
            idx = self.indices[word]
            self.count[idx] += n
            return idx
        idx = len(self.symbols)
        self.indices[word] = idx
        self.symbols.append(word)
        self.count.append(n)
        return idx

    # -- I/O ---------------------------------------------------------------------------------
    @classmethod
    def load(cls, f):
        d = cls()
        d.add_from_file(f)
        return d

    def add_from_file(self, f):
        if isinstance(f, str):
            try:
                with open(f, "r", encoding="utf-8") as fd:
                    return self.add_from_file(fd)
            except UnicodeError:
                raise Exception("Incorrect encoding detected in {}, please rebuild the dataset".format(f))
        lines = f.readlines()
        total = len(lines)
        for lineno, raw in enumerate(lines):
            body = raw.rstrip()
            head, sep, tail = body.rpartition(" ")
            if not sep:  # no count column: synthesise a descending pseudo-count
                head, tail = body, str(total - lineno)
            overwrite = False
            if tail == "#overwrite":
                overwrite = True
                head, _, tail = head.rpartition(" ")
            try:
                count = int(tail)
            except ValueError:
                raise ValueError("Incorrect dictionary format, expected '<token> <cnt> [flags]'")
            if head in self and not overwrite:
                logger.info(
                    "Duplicate word found when loading Dictionary: '{}', index is {}.".format(head, self.indices[head])
                )
                continue
            self.add_symbol(head, n=count, overwrite=overwrite)
