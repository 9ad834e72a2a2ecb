"""Make a DDP-style wrapper look like the module it wraps.

Attribute lookups fall through wrapper -> DDP object -> inner module; ``state_dict`` /
``load_state_dict`` address the inner module so checkpoints carry no ``module.`` prefixes
(reference ``unicore/distributed/module_proxy_wrapper.py:10-62``).
"""
from torch import nn


class ModuleProxyWrapper(nn.Module):
    def __init__(self, module: nn.Module):
        super().__init__()
        if not hasattr(module, "module"):
            raise TypeError("ModuleProxyWrapper expects a wrapped module to have a 'module' attribute")
        self.module = module

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)  # parameters / buffers / sub-modules of the wrapper
        except AttributeError:
            pass
        wrapper = super().__getattr__("module")
        try:
            return getattr(wrapper, name)  # e.g. DDP.no_sync
        except AttributeError:
            return getattr(wrapper.module, name)  # the user's model

    def state_dict(self, *args, **kwargs):
        return self.module.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.module.load_state_dict(*args, **kwargs)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def half(self):
        self.module.module.half()
        return self

    def bfloat16(self):
        self.module.module.bfloat16()
        return self
