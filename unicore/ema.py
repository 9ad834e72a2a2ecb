"""Exponential moving average of the model weights.

Parity: reference ``unicore/ema.py:6-65``: a deep copy of the model holds the average; for
mixed-precision runs (``is_flattened``) the shadow parameters are views of ONE flat fp32 buffer per
weight-decay group, laid out exactly like the optimizer's fp32 master so that group *i* of the EMA
pairs with ``optimizer.fp32_params[i]``; ``state_dict()`` = ``{"params": shadow.state_dict(),
"decay": d}``.  The update ``ema -= (1-d) * (ema - w)`` is one fused pass (``ops.ema_update_``)
instead of three element-wise kernels and a model-sized temporary.
"""
from copy import deepcopy

import torch

from unicore import ops
from unicore.optim.fp16_optimizer import flatten_parameters_fp32, separate_decay_params


class ExponentialMovingAverageModel:
    def __init__(self, args, model, decay, is_flattened=False):
        self.args = args
        self.decay = decay
        self.is_flattened = is_flattened
        source = model.module.module if hasattr(model, "module") and hasattr(model.module, "module") else model
        self.model_ema = deepcopy(source).float()
        self.model_ema.requires_grad_(False)
        # requires_grad is what separate_decay_params filters on: mirror the live model's flags
        for (_, p_src), (_, p_ema) in zip(source.named_parameters(), self.model_ema.named_parameters()):
            p_ema._ema_trainable = p_src.requires_grad  # noqa: SLF001
        if is_flattened:
            self.flatten_params = self.flatten_parameters()
        else:
            self.name2param = self.get_name2param()

    def get_name2param(self):
        return {n: p for n, p in self.model_ema.named_parameters() if getattr(p, "_ema_trainable", True)}

    def flatten_parameters(self):
        named = []
        for n, p in self.model_ema.named_parameters():
            if getattr(p, "_ema_trainable", True):
                p.requires_grad_(True)  # only so that separate_decay_params keeps it
                named.append((n, p))
        groups = separate_decay_params(self.args, named)
        flats = []
        for group in groups:
            flat = flatten_parameters_fp32(group["params"], set_to_param=True, set_grad=False)
            flat.requires_grad_(False)
            flats.append(flat)
        for _, p in named:
            p.requires_grad_(False)
        return flats

    @torch.no_grad()
    def update_one_param(self, ema_param, new_param):
        ops.ema_update_(ema_param.data, new_param.data, self.decay)

    @torch.no_grad()
    def update(self, new_param):
        """``new_param``: the optimizer's ``fp32_params`` groups (flattened) or ``named_parameters()``."""
        if self.is_flattened:
            for flat, group in zip(self.flatten_params, new_param):
                self.update_one_param(flat, group["params"][0])
        else:
            for name, p in new_param:
                name = name[len("module."):] if name.startswith("module.") and name not in self.name2param else name
                if name in self.name2param:
                    self.update_one_param(self.name2param[name], p)

    @torch.no_grad()
    def reset_from(self, model) -> None:
        """Restart the average from ``model``'s current weights (a checkpoint without EMA state was loaded)."""
        source = dict(model.state_dict())
        self.model_ema.load_state_dict({k: v.float() if torch.is_floating_point(v) else v for k, v in source.items()})

    def load_state_dict(self, state_dict):
        self.model_ema.load_state_dict(state_dict["params"])
        self.decay = state_dict["decay"] if "decay" in state_dict else self.decay

    def state_dict(self):
        return {"params": self.model_ema.state_dict(), "decay": self.decay}
