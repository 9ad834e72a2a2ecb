"""Base class every registered model extends (reference ``unicore/models/unicore_model.py:18-58``)."""
import logging

import torch
import torch.nn as nn

logger = logging.getLogger(__name__)


class BaseUnicoreModel(nn.Module):
    def __init__(self):
        super().__init__()

    @classmethod
    def add_args(cls, parser):
        """Add model-specific arguments to the parser."""
        pass

    @classmethod
    def build_model(cls, args, task):
        raise NotImplementedError("Model must implement the build_model method")

    def extract_features(self, *args, **kwargs):
        """Like ``forward`` but returns features instead of task outputs."""
        return self(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, model_args=None):
        """``model_args`` is accepted for plug-in compatibility (models may upgrade old state)."""
        return super().load_state_dict(state_dict, strict)

    def set_num_updates(self, num_updates):
        """Tell every sub-module that cares (schedules inside modules) the current update count."""

        def visit(m):
            if m is not self and hasattr(m, "set_num_updates"):
                m.set_num_updates(num_updates)

        self.apply(visit)
