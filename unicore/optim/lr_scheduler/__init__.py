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

# Module paths of the reference layout (one file per schedule) keep working as aliases, so plug-ins
# that do e.g. ``from unicore.optim.lr_scheduler.polynomial_decay_schedule import ...`` import unchanged.
_LEGACY_MODULES = {
    "cosine_lr_scheduler": ["CosineLRSchedule"],
    "exponential_decay_schedule": ["ExponentialDecayLRSchedule"],
    "fixed_schedule": ["FixedLRSchedule"],
    "inverse_square_root_schedule": ["InverseSquareRootSchedule"],
    "pass_through": ["PassThroughScheduleSchedule"],
    "polynomial_decay_schedule": ["PolynomialDecayLRSchedule"],
    "reduce_lr_on_plateau": ["ReduceLROnPlateauLRSchedule"],
    "tri_stage_lr_scheduler": ["TriStageLRSchedule"],
    "triangular_lr_scheduler": ["TriangularLRSchedule"],
}


def _install_legacy_modules():
    import sys
    import types

    here = sys.modules[__name__]
    for mod_name, names in _LEGACY_MODULES.items():
        full = __name__ + "." + mod_name
        if full in sys.modules:
            continue
        alias = types.ModuleType(full, "compatibility alias; see unicore.optim.lr_scheduler.schedules")
        for n in names:
            setattr(alias, n, getattr(schedules, n))
        alias.UnicoreLRScheduler = UnicoreLRScheduler
        alias.register_lr_scheduler = register_lr_scheduler
        sys.modules[full] = alias
        setattr(here, mod_name, alias)


_install_legacy_modules()
