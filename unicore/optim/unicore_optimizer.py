"""Base of the optimizer wrappers: a uniform facade over one ``torch.optim.Optimizer``.

Subclasses (``adam.py``, ``torch_wrappers.py`` ...) build ``self._optimizer`` and describe their CLI flags; the trainer
and ``FP16Optimizer`` only talk to the facade.  The method names and semantics are the reference's
(``unicore/optim/unicore_optimizer.py:10-191``: ``params``, ``get_lr`` / ``set_lr``, ``state_dict`` /
``load_state_dict`` with overrides, ``backward``, ``all_reduce_grads``, ``multiply_grads``, ``clip_grad_norm``,
``per_sample_clip_grad_norm``, ``step(scale=...)``, ``zero_grad`` and the capability properties); gradient scaling is one
multi-tensor launch on the GPU.
"""
import torch

from unicore import utils


class _ClippedAccumulator:
    """Gradients of micro-batches that were clipped one by one (``--per-sample-clip-norm``): each clipped ``.grad`` is
    added to a private buffer and cleared; before the real clip / update the buffers become the ``.grad`` again."""

    def __init__(self):
        self.buffers = None
        self.holds_grads = False

    def absorb(self, params):
        if self.buffers is None:
            self.buffers = [torch.zeros_like(p) for p in params]
        for buffer, p in zip(self.buffers, params):
            if p.grad is not None:
                buffer.add_(p.grad)
                p.grad = None
        self.holds_grads = True

    def hand_back(self, params):
        if not self.holds_grads:
            return
        for buffer, p in zip(self.buffers, params):
            p.grad = buffer
        self.holds_grads = False

    def clear(self):
        for buffer in self.buffers or ():
            buffer.zero_()


class UnicoreOptimizer(object):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self._clipped = _ClippedAccumulator()

    @classmethod
    def add_args(cls, parser):
        """Subclasses declare their command-line flags here."""

    # ---- the wrapped torch optimizer --------------------------------------------------------------------------------
    def _wrapped(self):
        inner = getattr(self, "_optimizer", None)
        if inner is None:
            raise NotImplementedError
        if not isinstance(inner, torch.optim.Optimizer):
            raise ValueError("_optimizer must be an instance of torch.optim.Optimizer")
        return inner

    @property
    def optimizer(self):
        return self._wrapped()

    @optimizer.setter
    def optimizer(self, optimizer):
        self._wrapped()   # only an already configured wrapper may swap its optimizer
        self._optimizer = optimizer

    @property
    def optimizer_config(self):
        """Keyword arguments of the wrapped optimizer as given on the command line; on resume they override what the
        checkpoint's parameter groups recorded."""
        raise NotImplementedError

    @property
    def param_groups(self):
        return self.optimizer.param_groups

    @property
    def params(self):
        """All parameters, group by group."""
        return (p for group in self.param_groups for p in group["params"])

    def __getstate__(self):
        return self._optimizer.__getstate__()

    # ---- learning rate / state ----------------------------------------------------------------------------------------
    def get_lr(self):
        return self.param_groups[0]["lr"]

    def set_lr(self, lr):
        for group in self.param_groups:
            group["lr"] = lr

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict, optimizer_overrides=None):
        """Restore the optimizer; ``optimizer_overrides`` (current CLI values: lr, betas ...) then win over the stored
        parameter-group entries."""
        self.optimizer.load_state_dict(state_dict)
        for group in self.param_groups if optimizer_overrides else ():
            group.update(optimizer_overrides)

    # ---- gradients ------------------------------------------------------------------------------------------------------
    def backward(self, loss):
        loss.backward()

    def all_reduce_grads(self, module):
        """Engines without autograd hooks (legacy DDP, the b200 engine's explicit path) reduce here."""
        reduce = getattr(module, "all_reduce_grads", None)
        if reduce is not None:
            reduce()

    def multiply_grads(self, c):
        from unicore import ops

        present = [p.grad.data for p in self.params if p.grad is not None]
        if present:
            ops.multi_tensor_scale_(present, c)

    def per_sample_clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        """Clip the gradients of the micro-batch that was just back-propagated and park them (see
        ``_ClippedAccumulator``).  Returns their norm (0.0 when clipping is off)."""
        if max_norm <= 0.0:
            return 0.0
        params = list(self.params)
        norm = utils.clip_grad_norm_(params, max_norm, aggregate_norm_fn)
        self._clipped.absorb(params)
        return norm

    def clip_grad_norm(self, max_norm, aggregate_norm_fn=None):
        params = list(self.params)
        self._clipped.hand_back(params)
        return utils.clip_grad_norm_(params, max_norm, aggregate_norm_fn)

    def step(self, closure=None, scale=1.0, groups=None):
        """One update.  ``scale`` divides the gradients - inside the update kernel when the wrapped optimizer can
        (``supports_step_with_scale``), by a multi-tensor pre-pass otherwise; ``groups`` goes to optimizers that can
        update a subset of their parameter groups (``supports_groups``)."""
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
        self._clipped.clear()


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
