"""Stub so that the UNMODIFIED reference package can be imported offline: it does
``import lmdb`` at package import time (reference unicore/data/lmdb_dataset.py:5) and the wheel is
not installable here. The benchmark never opens an LMDB."""


def open(*args, **kwargs):  # noqa: A001
    raise RuntimeError("lmdb is not available in this environment (stub)")
