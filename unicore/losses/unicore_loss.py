"""Loss base class.

Contract (reference ``unicore/losses/unicore_loss.py:14-78``): ``forward(model, sample, reduce=True)`` returns
``(loss, sample_size, logging_output)``; the static ``reduce_metrics`` turns the per-worker logging outputs into logged
metrics; ``logging_outputs_can_be_summed`` tells the trainer that the logging outputs are plain sums, which enables its
single-reduction statistics path.
"""
import inspect
from typing import Any, Dict, List

from torch.nn.modules.loss import _Loss


def _constructor_arguments(cls, args, task):
    """Values for the constructor parameters of ``cls``: ``task`` and ``args`` by name, anything else from the namespace
    ``args`` when it has an attribute of that name (parameters with defaults may stay unset)."""
    chosen = {}
    for param in inspect.signature(cls).parameters.values():
        if param.kind in (param.POSITIONAL_ONLY, param.VAR_POSITIONAL, param.VAR_KEYWORD):
            raise NotImplementedError("{} not supported".format(param.kind))
        if param.name in ("task", "args"):
            chosen[param.name] = task if param.name == "task" else args
        elif hasattr(args, param.name):
            chosen[param.name] = getattr(args, param.name)
        elif param.default is param.empty:
            raise NotImplementedError(
                "Unable to infer Loss arguments, please implement {}.build_loss".format(cls.__name__)
            )
    return chosen


class UnicoreLoss(_Loss):
    def __init__(self, task):
        super().__init__()
        self.task = task
        self.args = None if task is None else task.args
        if hasattr(task, "dictionary"):
            self.padding_idx = task.dictionary.pad()

    @classmethod
    def add_args(cls, parser):
        """Losses with command-line flags declare them here."""

    @classmethod
    def build_loss(cls, args, task):
        return cls(**_constructor_arguments(cls, args, task))

    def forward(self, model, sample, reduce=True):
        raise NotImplementedError

    @staticmethod
    def logging_outputs_can_be_summed(is_train: bool) -> bool:
        return False

    @staticmethod
    def reduce_metrics(logging_outputs: List[Dict[str, Any]], split="train") -> None:
        raise NotImplementedError
