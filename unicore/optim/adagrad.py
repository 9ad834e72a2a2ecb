"""Adagrad wrapper (reference ``unicore/optim/adagrad.py:13-41``)."""
import torch.optim

from . import UnicoreOptimizer, register_optimizer


@register_optimizer("adagrad")
class Adagrad(UnicoreOptimizer):
    def __init__(self, args, params):
        super().__init__(args)
        self._optimizer = torch.optim.Adagrad(params, **self.optimizer_config)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--weight-decay", "--wd", default=0.0, type=float, metavar="WD", help="weight decay")

    @property
    def optimizer_config(self):
        return {"lr": self.args.lr[0], "weight_decay": self.args.weight_decay}

    @property
    def supports_flat_params(self):
        return False
