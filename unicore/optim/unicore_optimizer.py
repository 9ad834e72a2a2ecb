"""Optimizer wrapper base: a thin, uniform facade over a ``torch.optim.Optimizer``.

Parity: reference ``unicore/optim/unicore_optimizer.py:10-191`` (``params``, ``get_lr/set_lr``,
``state_dict/load_state_dict`` with overrides, ``backward``, ``all_reduce_grads``,
``multiply_grads``, ``clip_grad_norm``, ``per_sample_clip_grad_norm``, ``step(scale=...)``,
``zero_grad``, capability properties).  ``multiply_grads`` is one multi-tensor launch on GPU.
"""
import torch

from unicore import utils


class UnicoreOptimizer(object):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self._grad_buffer = None
        self._need_sync_grad_buf = False

    @classmethod
    def add_args(cls, parser):
        pass

    # -- wrapped optimizer ----------------------------------------------------------------------
    def _checked(self):
        if not hasattr(self, "_optimizer"):
            raise NotImplementedError
        if not isinstance(self._optimizer, torch.optim.Optimizer):
            raise ValueError("_optimizer must be an instance of torch.optim.Optimizer")
        return self._optimizer

    @property
    def optimizer(self):
        return self._checked()

    @optimizer.setter
    def optimizer(self, optimizer):
        self._checked()
        self._optimizer = optimizer

    @property
    def optimizer_config(self):
        """kwargs that override values stored in a checkpoint's param groups on resume."""
        raise NotImplementedError

    @property
    def params(self):
        for group in self.param_groups:
            for p in group["params"]:
                yield p

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    def __getstate__(self):
        return self._optimizer.__getstate__()

    def get_lr(self):
        return self.param_groups[0]["lr"]

    def set_lr(self, lr):
        for group in self.param_groups:
            group["lr"] = lr

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict, optimizer_overrides=None):
        """Load state, then let current CLI values (lr etc.) and explicit overrides win."""
        self.optimizer.load_state_dict(state_dict)
        if optimizer_overrides:
            for group in self.param_groups:
                group.update(optimizer_overrides)

    # -- gradient plumbing ------------------------------------------------------------------------
    def backward(self, loss):
        loss.backward()

    def all_reduce_grads(self, module):
        """Explicit gradient sync for engines without autograd hooks (legacy DDP)."""
        if hasattr(module, "all_reduce_grads"):
            module.all_reduce_grads()

    def multiply_grads(self, c):
        from unicore import ops

        grads = [p.grad.data for p in self.params if p.grad is not None]
        if grads:
            ops.multi_tensor_scale_(grads, c)

    def per_sample_clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        """Clip this micro-batch's grads, move them into an accumulation buffer, clear ``.grad``."""
        if max_norm <= 0.0:
            return 0.0
        params = list(self.params)
        if self._grad_buffer is None:
            self._grad_buffer = [torch.zeros_like(p) for p in params]
        gnorm = utils.clip_grad_norm_(params, max_norm, aggregate_norm_fn)
        for buf, p in zip(self._grad_buffer, params):
            if p.grad is None:
                continue
            buf.add_(p.grad)
            p.grad = None
        self._need_sync_grad_buf = True
        return gnorm

    def _restore_grads_from_buffer(self):
        if not self._need_sync_grad_buf:
            return
        for buf, p in zip(self._grad_buffer, self.params):
            p.grad = buf
        self._need_sync_grad_buf = False

    def clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        self._restore_grads_from_buffer()
        return utils.clip_grad_norm_(list(self.params), max_norm, aggregate_norm_fn)

    def step(self, closure=None, scale=1.0, groups=None):
        """One update.  ``scale`` divides the gradients: inside the kernel when the wrapped optimizer can do that
        (``supports_step_with_scale``), by a multi-tensor pre-pass otherwise; ``groups`` is forwarded to optimizers
        that update a subset of their parameter groups (``supports_groups``)."""
        extra = {}
        if self.supports_step_with_scale:
            extra["scale"] = scale
        elif scale != 1.0:
            self.multiply_grads(1.0 / scale)
        if self.supports_groups:
            extra["groups"] = groups
        self.optimizer.step(closure, **extra)

    def zero_grad(self):
        for p in self.params:
            p.grad = None
        self.optimizer.zero_grad()
        if self._grad_buffer is not None:
            for buf in self._grad_buffer:
                buf.zero_()



def _capability(name, doc):
    """Capability flags are read off the wrapped ``torch.optim.Optimizer`` (absent means False)."""
    return property(lambda self: getattr(self.optimizer, name, False), doc=doc)


for _name, _doc in (
    ("supports_memory_efficient_fp16", "the update can run on half-precision parameters directly"),
    ("supports_step_with_scale", "``step(scale=...)`` divides the gradients inside the update kernel"),
    ("supports_groups", "``step(groups=...)`` restricts the update to some parameter groups"),
    ("supports_flat_params", "the update is correct when all parameters are views of one flat tensor"),
):
    setattr(UnicoreOptimizer, _name, _capability(_name, _doc))
del _name, _doc
