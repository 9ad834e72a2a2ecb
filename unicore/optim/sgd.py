"""``--optimizer sgd`` (reference ``unicore/optim/sgd.py:13``); built by :mod:`unicore.optim.torch_wrappers`."""
from .torch_wrappers import SGD  # noqa: F401
