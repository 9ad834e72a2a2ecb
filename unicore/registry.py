"""Generic component registries (optimizer / lr-scheduler / loss ...).

A registry is a small object that owns (a) the name -> class table, (b) the CLI flag that
selects an entry and (c) the rule for turning ``args`` into an instance.  The module level
``REGISTRIES`` dict is what ``unicore.options`` walks to add one ``--<registry>`` flag per
registry and to let the *selected* class contribute its own flags.

Behavioural parity with the reference (``unicore/registry.py:13-81``): ``setup_registry`` returns
``(build_x, register_x, REGISTRY)``; building prefers a ``build_<name>`` classmethod over the
constructor; before building, defaults declared in the class's ``add_args`` are back-filled into
``args`` so components can be constructed without going through the CLI.
"""
import argparse
from typing import Callable, Dict, Optional, Tuple

REGISTRIES: Dict[str, dict] = {}


class _Registry:
    def __init__(self, flag: str, base_class: Optional[type], default: Optional[str]):
        if not flag.startswith("--"):
            raise ValueError("registry flag must look like --name, got %r" % flag)
        self.flag = flag
        self.key = flag[2:].replace("-", "_")
        self.base_class = base_class
        self.default = default
        self.table: Dict[str, type] = {}
        self._class_names = set()

    # -- registration -----------------------------------------------------------------------
    def register(self, name: str) -> Callable[[type], type]:
        def _decorator(cls: type) -> type:
            if name in self.table:
                raise ValueError("Cannot register duplicate %s (%s)" % (self.key, name))
            if cls.__name__ in self._class_names:
                raise ValueError(
                    "Cannot register %s with duplicate class name (%s)" % (self.key, cls.__name__)
                )
            if self.base_class is not None and not issubclass(cls, self.base_class):
                raise ValueError(
                    "%s must extend %s" % (cls.__name__, self.base_class.__name__)
                )
            self.table[name] = cls
            self._class_names.add(cls.__name__)
            return cls

        return _decorator

    # -- construction -----------------------------------------------------------------------
    def build(self, args, *extra_args, **extra_kwargs):
        choice = getattr(args, self.key, None)
        if choice is None:
            return None
        cls = self.table[choice]
        factory = getattr(cls, "build_" + self.key, cls)
        set_defaults(args, cls)
        return factory(args, *extra_args, **extra_kwargs)


def setup_registry(
    registry_name: str, base_class: Optional[type] = None, default: Optional[str] = None
) -> Tuple[Callable, Callable, Dict[str, type]]:
    """Create a registry; see module docstring. ``registry_name`` is the CLI flag (``--optimizer``)."""
    reg = _Registry(registry_name, base_class, default)
    if reg.key in REGISTRIES:
        raise ValueError("registry %s already exists" % reg.key)
    REGISTRIES[reg.key] = {"registry": reg.table, "default": default, "object": reg}
    return reg.build, reg.register, reg.table


def set_defaults(args, cls) -> None:
    """Fill ``args`` with the defaults a class declares in ``add_args`` (only missing attrs)."""
    add_args = getattr(cls, "add_args", None)
    if add_args is None:
        return
    probe = argparse.ArgumentParser(
        argument_default=argparse.SUPPRESS, allow_abbrev=False, add_help=False
    )
    add_args(probe)
    found = argparse.Namespace()
    for action in probe._actions:  # noqa: SLF001 - argparse has no public iterator
        if action.dest is argparse.SUPPRESS or action.dest == "help":
            continue
        if hasattr(found, action.dest):
            continue
        if action.default is not argparse.SUPPRESS:
            setattr(found, action.dest, action.default)
    for key, value in vars(found).items():
        if not hasattr(args, key):
            setattr(args, key, value)
