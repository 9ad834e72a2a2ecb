"""Base class of the learning-rate schedules (interface of reference
``unicore/optim/lr_scheduler/unicore_lr_scheduler.py:12-50``).

The trainer drives a schedule through three entry points:

========================  ==========================================================================
``step_update(n)``        once with ``n = 0`` when the schedule is built, then after every successful
                          parameter update; returns the learning rate now in force
``step_begin_epoch(e)``   at the start of epoch ``e``
``step(e, val_loss)``     at the end of epoch ``e``; the base class only tracks the best validation loss
========================  ==========================================================================

``state_dict`` / ``load_state_dict`` carry that best loss through checkpoints.
"""
from unicore.optim import UnicoreOptimizer


def _better(current, candidate):
    """Smaller validation loss wins; ``None`` means "nothing seen yet"."""
    if candidate is None:
        return current
    return candidate if current is None else min(current, candidate)


class UnicoreLRScheduler(object):
    def __init__(self, args, optimizer, total_train_steps):
        if not (optimizer is None or isinstance(optimizer, UnicoreOptimizer)):
            raise ValueError("optimizer must be an instance of UnicoreOptimizer")
        self.args, self.optimizer = args, optimizer
        self.total_train_steps = total_train_steps
        self.best = None

    # -- flags contributed by the concrete schedule ----------------------------------------------------
    @classmethod
    def add_args(cls, parser):
        """Schedules override this to register their own flags."""

    # -- driven by the trainer ---------------------------------------------------------------------------
    def step_update(self, num_updates):
        return self.optimizer.get_lr()

    def step_begin_epoch(self, epoch):
        """Hook for schedules that change per epoch."""

    def step(self, epoch, val_loss=None):
        self.best = _better(self.best, val_loss)

    # -- checkpointing -----------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        self.best = state_dict["best"]

    def state_dict(self):
        return dict(best=self.best)
