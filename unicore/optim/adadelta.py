"""``--optimizer adadelta`` (reference ``unicore/optim/adadelta.py:13``); built by :mod:`unicore.optim.torch_wrappers`."""
from .torch_wrappers import Adadelta  # noqa: F401
