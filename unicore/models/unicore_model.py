"""``BaseUnicoreModel``: what the trainer, the checkpoint code and the task registry expect from a model
(interface of reference ``unicore/models/unicore_model.py:18-58``).

=====================  ==============================================================================
``add_args(parser)``   classmethod; the model's own command-line flags (two-pass option parsing)
``build_model``        classmethod factory ``(args, task) -> model``; every registered model provides it
``extract_features``   forward pass that stops before the task head (default: the full forward)
``load_state_dict``    accepts the extra ``model_args`` the checkpoint loader passes along
``set_num_updates``    fan-out of the update counter to sub-modules that schedule on it
=====================  ==============================================================================
"""
import logging

import torch.nn as nn

logger = logging.getLogger(__name__)


class BaseUnicoreModel(nn.Module):
    # -- construction ------------------------------------------------------------------------------------
    @classmethod
    def build_model(cls, args, task):
        raise NotImplementedError("{} does not implement build_model(args, task)".format(cls.__name__))

    @classmethod
    def add_args(cls, parser):
        """Models override this to contribute flags; the base model has none."""

    # -- inference helpers -------------------------------------------------------------------------------
    def extract_features(self, *args, **kwargs):
        """Forward pass that stops before the task head; models with a head override it."""
        return self(*args, **kwargs)

    # -- training-loop hooks -----------------------------------------------------------------------------
    def set_num_updates(self, num_updates):
        """Every sub-module that defines ``set_num_updates`` (e.g. in-module schedules) learns the update count."""
        listeners = getattr(self, "_num_updates_listeners", None)
        n_modules = sum(1 for _ in self.modules())
        if listeners is None or listeners[0] != n_modules:  # (re)discover when the module tree changed size
            found = [m for m in self.modules() if m is not self and hasattr(m, "set_num_updates")]
            listeners = (n_modules, found)
            object.__setattr__(self, "_num_updates_listeners", listeners)
        for module in listeners[1]:
            module.set_num_updates(num_updates)

    # -- checkpoints -------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, model_args=None):
        # ``model_args`` (the args stored in the checkpoint) lets plug-in models upgrade old state; unused here
        return nn.Module.load_state_dict(self, state_dict, strict)
