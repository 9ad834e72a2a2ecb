"""Optimizer registry (``--optimizer``, default ``adam``) and the mixed-precision wrapper.
Reference ``unicore/optim/__init__.py:17-37``."""
from unicore import registry
from unicore.optim.unicore_optimizer import UnicoreOptimizer  # noqa: F401
from unicore.optim.fp16_optimizer import FP16Optimizer, separate_decay_params  # noqa: F401

__all__ = ["UnicoreOptimizer", "FP16Optimizer"]

_build_optimizer, register_optimizer, OPTIMIZER_REGISTRY = registry.setup_registry(
    "--optimizer", base_class=UnicoreOptimizer, default="adam"
)


def build_optimizer(args, params, separate=True, *extra_args, **extra_kwargs):
    """``params`` is a list of ``(name, param)`` (or already param groups when ``separate=False``)."""
    if separate:
        params = separate_decay_params(args, params)
    return _build_optimizer(args, params, *extra_args, **extra_kwargs)


from . import adadelta, adagrad, adam, sgd  # noqa: E402,F401
from . import lr_scheduler  # noqa: E402,F401
