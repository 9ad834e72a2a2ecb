"""One data-parallel replica: the task's model and loss, cast / moved / re-tied, and wrapped on demand."""
import logging

import torch

from unicore import models, utils

logger = logging.getLogger(__name__)


def _walk_parameters(module, prefix=""):
    """Yield ``(dotted path, Parameter)`` for every registration site, i.e. shared parameters appear once per site."""
    for name, param in module._parameters.items():
        if param is not None:
            yield (prefix + name, param)
    for name, child in module._modules.items():
        if child is not None:
            yield from _walk_parameters(child, prefix + name + ".")


def _tied_sites(module):
    """Groups of paths that hold the same Parameter object."""
    sites = {}
    for path, param in _walk_parameters(module):
        sites.setdefault(id(param), []).append(path)
    return [paths for paths in sites.values() if len(paths) > 1]


def _resolve(module, path):
    *parents, leaf = path.split(".")
    for name in parents:
        module = getattr(module, name)
    return module, leaf


class Replica:
    """Holds the bare modules (``.model`` / ``.loss``) and builds their data-parallel wrappers lazily."""

    def __init__(self, args, model, loss, device, distributed: bool):
        self.args, self.device, self.distributed = args, device, distributed
        ties = _tied_sites(model)
        if args.fp16:
            model, loss = model.half(), loss.half()
        elif args.bf16:
            model, loss = model.bfloat16(), loss.bfloat16()
        if not distributed:  # the data-parallel wrapper moves its module itself
            model, loss = model.to(device=device), loss.to(device=device)
        # casting / moving re-creates Parameter objects: point every tied site at the first one again
        for paths in ties:
            owner, leaf = _resolve(model, paths[0])
            shared = getattr(owner, leaf)
            for other in paths[1:]:
                logger.info("detected shared parameter: {} <- {}".format(paths[0], other))
                owner, leaf = _resolve(model, other)
                setattr(owner, leaf, shared)
        self.model, self.loss = model, loss
        self.wrapped_model = None
        self.wrapped_loss = None

    def drop_wrappers(self):
        self.wrapped_model = self.wrapped_loss = None

    def _wrap(self, module, process_group):
        return models.DistributedUnicoreModel(self.args, module, process_group=process_group, device=self.device)

    def get_model(self, process_group):
        if self.wrapped_model is None:
            self.wrapped_model = self._wrap(self.model, process_group) if self.distributed else self.model
        return self.wrapped_model

    def get_loss(self, process_group):
        if self.wrapped_loss is None:
            wrap = self.distributed and utils.has_parameters(self.loss) and not self._loss_rides_with_model()
            self.wrapped_loss = self._wrap(self.loss, process_group) if wrap else self.loss
        return self.wrapped_loss

    def _loss_rides_with_model(self) -> bool:
        """Under ``--ddp-backend b200`` the loss parameters sit in the model engine's flat arenas (they are part of the
        same optimizer groups), are reduced by its buckets and counted in its norm: a second wrapper would only add a
        redundant per-tensor NCCL pass over already averaged gradients."""
        return self.engine is not None and hasattr(self.engine, "attach_optimizer")

    @property
    def engine(self):
        """The innermost data-parallel engine object (``None`` when not distributed / not yet wrapped)."""
        w = self.wrapped_model
        if w is None or w is self.model:
            return None
        return getattr(w, "module", None)
