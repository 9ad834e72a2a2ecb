"""Task registry.  Tasks own the data (dictionary, datasets, batch iterators) and define how a
training/validation step uses model + loss.  Built-in tasks are imported explicitly below;
external ones arrive through ``--user-dir``.  Parity: reference ``unicore/tasks/__init__.py:17-86``.
"""
from .unicore_task import StatefulContainer, UnicoreTask  # noqa: F401

TASK_REGISTRY = {}
TASK_CLASS_NAMES = set()


def setup_task(args, **kwargs):
    return TASK_REGISTRY[args.task].setup_task(args, **kwargs)


def register_task(name):
    """Class decorator: ``@register_task("bert")`` makes the task selectable via ``--task bert``."""

    def _register(cls):
        if name in TASK_REGISTRY:
            raise ValueError("Cannot register duplicate task ({})".format(name))
        if not issubclass(cls, UnicoreTask):
            raise ValueError("Task ({}: {}) must extend UnicoreTask".format(name, cls.__name__))
        if cls.__name__ in TASK_CLASS_NAMES:
            raise ValueError("Cannot register task with duplicate class name ({})".format(cls.__name__))
        TASK_REGISTRY[name] = cls
        TASK_CLASS_NAMES.add(cls.__name__)
        return cls

    return _register


def get_task(name):
    return TASK_REGISTRY[name]


from . import synthetic  # noqa: E402,F401  (registers the built-in synthetic benchmark tasks)
