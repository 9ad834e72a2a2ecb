"""LR scheduler registry (``--lr-scheduler``, default ``fixed``).
Reference ``unicore/optim/lr_scheduler/__init__.py:17-34``."""
from unicore import registry
from unicore.optim.lr_scheduler.unicore_lr_scheduler import UnicoreLRScheduler  # noqa: F401

build_lr_scheduler_, register_lr_scheduler, LR_SCHEDULER_REGISTRY = registry.setup_registry(
    "--lr-scheduler", base_class=UnicoreLRScheduler, default="fixed"
)


def build_lr_scheduler(args, optimizer, total_train_steps):
    return build_lr_scheduler_(args, optimizer, total_train_steps)


from . import schedules  # noqa: E402,F401  (registers the nine built-in schedules)
