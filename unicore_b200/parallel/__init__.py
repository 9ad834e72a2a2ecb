"""Symmetric-memory data parallelism (hand-written NVLink peer-memory reductions fused with Adam)."""
import os

from .symm_dp import SymmDataParallel, symm_available  # noqa: F401


def reference_tail_requested() -> bool:
    """``UNICORE_B200_REFERENCE_TAIL=1``: run ``--ddp-backend b200`` on ANY device / backend (CPU + gloo included) with
    the PyTorch specification of the fused optimizer tail (``reference_tail.py``) - for debugging and for the CPU
    test-suite; the fused-Adam code path (with its PyTorch fallback math) is selected even without a GPU."""
    return os.environ.get("UNICORE_B200_REFERENCE_TAIL", "0") == "1"
