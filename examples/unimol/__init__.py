"""Uni-Mol style plug-in (``--user-dir examples/unimol``): model ``unimol`` (arch ``unimol_base``),
loss ``unimol`` and a synthetic-molecule task ``synthetic_unimol`` for benchmarks (BASELINE config 4)."""
from . import loss, model, task  # noqa: F401
