"""``softmax_dropout`` functional (implementation: ``unicore_b200/ops/softmax_ops.py``)."""
from unicore.ops import softmax_dropout  # noqa: F401
