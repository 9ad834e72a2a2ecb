"""LR scheduler base (reference ``unicore/optim/lr_scheduler/unicore_lr_scheduler.py:12-50``).

Call protocol: ``step_update(num_updates)`` at build time with 0 and after
every successful update; ``step_begin_epoch(epoch)`` at each epoch start; ``step(epoch, val_loss)``
at each epoch end.
"""
from unicore.optim import UnicoreOptimizer


class UnicoreLRScheduler(object):
    def __init__(self, args, optimizer, total_train_steps):
        super().__init__()
        if optimizer is not None and not isinstance(optimizer, UnicoreOptimizer):
            raise ValueError("optimizer must be an instance of UnicoreOptimizer")
        self.args = args
        self.optimizer = optimizer
        self.total_train_steps = total_train_steps
        self.best = None

    @classmethod
    def add_args(cls, parser):
        pass

    def state_dict(self):
        return {"best": self.best}

    def load_state_dict(self, state_dict):
        self.best = state_dict["best"]

    def step_begin_epoch(self, epoch):
        pass

    def step(self, epoch, val_loss=None):
        if val_loss is not None:
            self.best = val_loss if self.best is None else min(self.best, val_loss)

    def step_update(self, num_updates):
        return self.optimizer.get_lr()
