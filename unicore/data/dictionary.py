"""Symbol <-> index vocabulary with BERT-style special tokens.

Same contract as the reference's ``unicore/data/dictionary.py:12-148``: the four roles ``bos / pad / eos / unk`` default
to ``[CLS] [PAD] [SEP] [UNK]``; ``add_symbol`` / ``index`` / ``vec_index`` / ``special_index``; a text file with one
``<symbol> [count] [#overwrite]`` entry per line.  Unknown symbols map to the ``unk`` index, out-of-range indices to
the ``unk`` word.
"""
import logging

import numpy as np

logger = logging.getLogger(__name__)

_ROLES = ("bos", "pad", "eos", "unk")


class Dictionary:
    def __init__(self, *, bos="[CLS]", pad="[PAD]", eos="[SEP]", unk="[UNK]", extra_special_symbols=None):
        words = dict(bos=bos, pad=pad, eos=eos, unk=unk)
        for role in _ROLES:
            setattr(self, role + "_word", words[role])
        self.symbols, self.count, self.indices = [], [], {}
        self.specials = set(words.values()) | set(extra_special_symbols or ())

    # ---- reading ------------------------------------------------------------------------------------------------------
    def __len__(self):
        return len(self.symbols)

    def __contains__(self, sym):
        return sym in self.indices

    def __eq__(self, other):
        return isinstance(other, Dictionary) and other.indices == self.indices

    def __getitem__(self, idx):
        in_range = 0 <= idx < len(self.symbols)
        return self.symbols[idx] if in_range else self.unk_word

    def index(self, sym):
        """Index of ``sym``; the ``unk`` index when it is not in the vocabulary."""
        if not isinstance(sym, str):
            raise TypeError("symbols are strings")
        found = self.indices.get(sym)
        return self.indices[self.unk_word] if found is None else found

    def vec_index(self, a):
        return np.vectorize(self.index)(a)

    def special_index(self):
        return [self.index(word) for word in self.specials]

    # ---- writing ------------------------------------------------------------------------------------------------------
    def add_symbol(self, word, n=1, overwrite=False, is_special=False):
        """Register ``word`` with count ``n`` - or add ``n`` to the count of an existing entry - and return its index.
        ``overwrite`` appends a NEW entry for an existing word (the file format's ``#overwrite`` flag)."""
        if is_special:
            self.specials.add(word)
        known = self.indices.get(word)
        if known is not None and not overwrite:
            self.count[known] += n
            return known
        position = len(self.symbols)
        self.symbols.append(word)
        self.count.append(n)
        self.indices[word] = position
        return position

    # ---- files --------------------------------------------------------------------------------------------------------
    @classmethod
    def load(cls, f):
        vocabulary = cls()
        vocabulary.add_from_file(f)
        return vocabulary

    @staticmethod
    def _parse_entry(text, default_count):
        """``"<symbol> [count] [#overwrite]"`` -> (symbol, count, overwrite); a missing count becomes ``default_count``."""
        symbol, space, last = text.rpartition(" ")
        if not space:
            return text, default_count, False
        overwrite = last == "#overwrite"
        if overwrite:
            symbol, _, last = symbol.rpartition(" ")
        try:
            return symbol, int(last), overwrite
        except ValueError:
            raise ValueError("Incorrect dictionary format, expected '<token> <cnt> [flags]'")

    def add_from_file(self, f):
        if isinstance(f, str):
            try:
                with open(f, "r", encoding="utf-8") as handle:
                    return self.add_from_file(handle)
            except UnicodeError:
                raise Exception("Incorrect encoding detected in {}, please rebuild the dataset".format(f))
        entries = f.readlines()
        for position, line in enumerate(entries):
            # without a count column the entries get descending pseudo-counts (file order = frequency order)
            symbol, count, overwrite = self._parse_entry(line.rstrip(), len(entries) - position)
            if symbol in self.indices and not overwrite:
                logger.info("Duplicate word found when loading Dictionary: '{}', index is {}.".format(
                    symbol, self.indices[symbol]))
                continue
            self.add_symbol(symbol, n=count, overwrite=overwrite)


def _role_accessor(role):
    def index_of_role(self):
        return self.index(getattr(self, role + "_word"))

    index_of_role.__name__ = role
    index_of_role.__doc__ = "Index of the ``{}`` symbol.".format(role)
    return index_of_role


for _role in _ROLES:
    setattr(Dictionary, _role, _role_accessor(_role))
del _role
