"""Model + architecture registries.

``@register_model("bert")`` registers a model class; ``@register_model_architecture("bert",
"bert_large")`` registers a function that fills architecture hyper-parameters into ``args``.
``--arch`` selects an architecture name; ``build_model`` instantiates its model class.
Parity: reference ``unicore/models/__init__.py:17-120``.
"""
from .unicore_model import BaseUnicoreModel  # noqa: F401
from .distributed_unicore_model import DistributedUnicoreModel  # noqa: F401

MODEL_REGISTRY = {}
ARCH_MODEL_REGISTRY = {}
ARCH_MODEL_INV_REGISTRY = {}
ARCH_CONFIG_REGISTRY = {}

__all__ = [
    "BaseUnicoreModel", "DistributedUnicoreModel", "build_model", "register_model",
    "register_model_architecture", "MODEL_REGISTRY", "ARCH_MODEL_REGISTRY",
    "ARCH_MODEL_INV_REGISTRY", "ARCH_CONFIG_REGISTRY",
]


def build_model(args, task):
    return ARCH_MODEL_REGISTRY[args.arch].build_model(args, task)


def register_model(name):
    def _register(cls):
        if name in MODEL_REGISTRY:
            raise ValueError("Cannot register duplicate model ({})".format(name))
        if not issubclass(cls, BaseUnicoreModel):
            raise ValueError("Model ({}: {}) must extend BaseUnicoreModel".format(name, cls.__name__))
        MODEL_REGISTRY[name] = cls
        return cls

    return _register


def register_model_architecture(model_name, arch_name):
    def _register(fn):
        if model_name not in MODEL_REGISTRY:
            raise ValueError("Cannot register model architecture for unknown model type ({})".format(model_name))
        if arch_name in ARCH_MODEL_REGISTRY:
            raise ValueError("Cannot register duplicate model architecture ({})".format(arch_name))
        if not callable(fn):
            raise ValueError("Model architecture must be callable ({})".format(arch_name))
        ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
        ARCH_MODEL_INV_REGISTRY.setdefault(model_name, []).append(arch_name)
        ARCH_CONFIG_REGISTRY[arch_name] = fn
        return fn

    return _register
