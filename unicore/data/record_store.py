"""Single-file record store: the on-disk format the data pipeline falls back to when ``lmdb`` is not installed.

The reference keeps every corpus in a single-file LMDB of pickled records keyed by the decimal record index
(``unicore/data/lmdb_dataset.py:16-49``, ``examples/bert/example_data/preprocess.py``).  LMDB is an optional
C dependency; this store gives the same access pattern - O(1) random reads of pickled records from one
memory-mapped file, safe to share between DataLoader workers - with nothing but the standard library::

    offset 0   8 bytes   magic  b"UCRSTOR1"
    offset 8   uint64    number of records n
    offset 16  uint64    byte offset of the index
    offset 24  ...       record payloads, back to back
    index      (n + 1) x uint64   payload i spans [index[i], index[i + 1])

``LMDBDataset`` sniffs the magic and reads either format, so ``<split>.lmdb`` paths on the command line keep
working whichever writer produced the file.
"""
import mmap
import os
import pickle
import struct

import numpy as np

MAGIC = b"UCRSTOR1"
_HEADER = struct.Struct("<8sQQ")


def is_record_store(path) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(len(MAGIC)) == MAGIC
    except OSError:
        return False


class RecordStoreWriter:
    """Append-only writer; ``close()`` (or leaving the ``with`` block) seals the index."""

    def __init__(self, path):
        self.path = path
        self._f = open(path, "wb")
        self._f.write(_HEADER.pack(MAGIC, 0, 0))
        self._offsets = [self._f.tell()]

    def append(self, obj) -> int:
        self._f.write(pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL))
        self._offsets.append(self._f.tell())
        return len(self._offsets) - 2

    def close(self):
        if self._f is None:
            return
        index_at = self._offsets[-1]
        self._f.write(np.asarray(self._offsets, dtype="<u8").tobytes())
        self._f.seek(0)
        self._f.write(_HEADER.pack(MAGIC, len(self._offsets) - 1, index_at))
        self._f.close()
        self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class RecordStoreReader:
    """Memory-mapped reader.  The mapping is opened lazily and dropped on pickling, so instances can be handed
    to DataLoader worker processes (each maps the file itself; pages are shared through the page cache)."""

    def __init__(self, path):
        self.path = path
        if not os.path.isfile(path):
            raise FileNotFoundError("{} not found".format(path))
        with open(path, "rb") as f:
            magic, count, index_at = _HEADER.unpack(f.read(_HEADER.size))
            if magic != MAGIC:
                raise ValueError("{} is not a record store".format(path))
            if count == 0 and index_at == 0:
                raise ValueError("{} was not sealed (writer not closed)".format(path))
            f.seek(index_at)
            self._index = np.frombuffer(f.read(8 * (count + 1)), dtype="<u8").astype(np.int64)
        if len(self._index) != count + 1:
            raise ValueError("{} is truncated".format(path))
        self._map = None

    def __len__(self):
        return len(self._index) - 1

    def _mapping(self):
        if self._map is None:
            with open(self.path, "rb") as f:
                self._map = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        return self._map

    def read_bytes(self, idx) -> memoryview:
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        return memoryview(self._mapping())[self._index[idx]:self._index[idx + 1]]

    def __getitem__(self, idx):
        return pickle.loads(self.read_bytes(idx))

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_map"] = None
        return state
