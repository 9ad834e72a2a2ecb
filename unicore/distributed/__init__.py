"""Distributed runtime: launcher, collective helpers and data-parallel engines."""
from .legacy_distributed_data_parallel import LegacyDistributedDataParallel
from .module_proxy_wrapper import ModuleProxyWrapper

__all__ = ["LegacyDistributedDataParallel", "ModuleProxyWrapper"]
