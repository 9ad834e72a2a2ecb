"""``--optimizer adagrad`` (reference ``unicore/optim/adagrad.py:13``); built by :mod:`unicore.optim.torch_wrappers`."""
from .torch_wrappers import Adagrad  # noqa: F401
