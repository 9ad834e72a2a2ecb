"""Loss base class.  Contract: ``forward(model, sample, reduce=True) -> (loss, sample_size,
logging_output)``; ``reduce_metrics`` (static) turns the per-worker logging outputs into metrics.
Parity: reference ``unicore/losses/unicore_loss.py:14-78``.
"""
import inspect
from typing import Any, Dict, List

from torch.nn.modules.loss import _Loss



class UnicoreLoss(_Loss):
    def __init__(self, task):
        super().__init__()
        self.task = task
        self.args = task.args if task is not None else None
        if task is not None and hasattr(task, "dictionary"):
            self.padding_idx = task.dictionary.pad()

    @classmethod
    def add_args(cls, parser):
        pass

    @classmethod
    def build_loss(cls, args, task):
        """Instantiate ``cls`` by matching constructor parameter names against ``task`` / ``args``."""
        kwargs = {}
        for p in inspect.signature(cls).parameters.values():
            if p.kind in (p.POSITIONAL_ONLY, p.VAR_POSITIONAL, p.VAR_KEYWORD):
                raise NotImplementedError("{} not supported".format(p.kind))
            if p.name == "task":
                kwargs["task"] = task
            elif p.name == "args":
                kwargs["args"] = args
            elif hasattr(args, p.name):
                kwargs[p.name] = getattr(args, p.name)
            elif p.default is not p.empty:
                continue
            else:
                raise NotImplementedError(
                    "Unable to infer Loss arguments, please implement {}.build_loss".format(cls.__name__)
                )
        return cls(**kwargs)

    def forward(self, model, sample, reduce=True):
        raise NotImplementedError

    @staticmethod
    def logging_outputs_can_be_summed(is_train: bool) -> bool:
        """True enables the single-all-reduce stat sync in the trainer."""
        return False

    @staticmethod
    def reduce_metrics(logging_outputs: List[Dict[str, Any]], split="train") -> None:
        raise NotImplementedError
