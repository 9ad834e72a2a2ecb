"""Symmetric-memory data parallelism (hand-written NVLink peer-memory reductions fused with Adam)."""
from .symm_dp import SymmDataParallel, symm_available  # noqa: F401
