"""Model and architecture registries (reference ``unicore/models/__init__.py:17-120``).

``@register_model("bert")`` registers a model class; ``@register_model_architecture("bert", "bert_large")`` registers a
function that fills an architecture's hyper-parameters into ``args``; ``--arch`` names an architecture and
``build_model`` instantiates the model class it belongs to.
"""
from .unicore_model import BaseUnicoreModel  # noqa: F401
from .distributed_unicore_model import DistributedUnicoreModel  # noqa: F401

MODEL_REGISTRY = {}            # model name   -> class
ARCH_MODEL_REGISTRY = {}       # architecture -> class
ARCH_MODEL_INV_REGISTRY = {}   # model name   -> [architectures]
ARCH_CONFIG_REGISTRY = {}      # architecture -> function(args)

__all__ = [
    "BaseUnicoreModel", "DistributedUnicoreModel", "build_model", "register_model",
    "register_model_architecture", "MODEL_REGISTRY", "ARCH_MODEL_REGISTRY",
    "ARCH_MODEL_INV_REGISTRY", "ARCH_CONFIG_REGISTRY",
]


def build_model(args, task):
    return ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)


def _require(condition, message, *fields):
    if not condition:
        raise ValueError(message.format(*fields))


def register_model(name):
    def decorator(cls):
        _require(name not in MODEL_REGISTRY, "Cannot register duplicate model ({})", name)
        _require(issubclass(cls, BaseUnicoreModel), "Model ({}: {}) must extend BaseUnicoreModel", name, cls.__name__)
        MODEL_REGISTRY[name] = cls
        return cls

    return decorator


def register_model_architecture(model_name, arch_name):
    def decorator(fill_defaults):
        _require(model_name in MODEL_REGISTRY,
                 "Cannot register model architecture for unknown model type ({})", model_name)
        _require(arch_name not in ARCH_MODEL_REGISTRY, "Cannot register duplicate model architecture ({})", arch_name)
        _require(callable(fill_defaults), "Model architecture must be callable ({})", arch_name)
        ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
        ARCH_CONFIG_REGISTRY[arch_name] = fill_defaults
        ARCH_MODEL_INV_REGISTRY.setdefault(model_name, []).append(arch_name)
        return fill_defaults

    return decorator
