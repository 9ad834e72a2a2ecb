"""Meters, global metric aggregation and progress-bar log sinks."""
from . import meters, metrics, progress_bar  # noqa: F401
