"""Locate the first module that produces NaN/Inf (forward or backward) during a replayed step.

Used by the trainer when the gradient norm is non-finite or inconsistent across ranks
(reference ``unicore/nan_detector.py:13-109``, call site ``trainer.py:727-748``).
"""
import logging

import torch

logger = logging.getLogger(__name__)


class NanDetector:
    def __init__(self, model, forward=True, backward=True):
        self.named_parameters = list(model.named_parameters())
        self.fhooks, self.bhooks = [], []
        self.forward, self.backward = forward, backward
        self._reported = {"forward": False, "backward": False}
        for name, mod in model.named_modules():
            mod.__dict__["_nan_detector_name"] = name
            self.add_hooks(mod)

    def add_hooks(self, module):
        """Attach the scanners to one more module (reference ``nan_detector.py:52``)."""
        if self.forward:
            self.fhooks.append(module.register_forward_hook(self.fhook_fn))
        if self.backward:
            self.bhooks.append(module.register_full_backward_hook(self.bhook_fn))

    def reset(self):
        """Report again on the next non-finite tensor of either pass."""
        self._reported = {"forward": False, "backward": False}

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        # dump per-parameter gradient norms if any is non-finite
        norms = {}
        for name, p in self.named_parameters:
            if p.grad is not None:
                norms[name] = torch.norm(p.grad.data.float(), p=2).item()
        bad = {k: v for k, v in norms.items() if v != v or abs(v) == float("inf")}
        if bad:
            logger.info("Detected nan/inf grad norm, dumping norms...")
            logger.info("norms: {}".format(norms))
            logger.info("gradients: {}".format(sorted(bad)))
        self.close()
        return False

    def _describe(self, module, tensor, kind):
        if tensor is None or not torch.is_tensor(tensor) or tensor.numel() < 1 or not tensor.is_floating_point():
            return None
        with torch.no_grad():
            flat = tensor.detach().float()
            what = "NaN" if torch.isnan(flat).any() else ("Inf" if torch.isinf(flat).any() else None)
            if what is None:
                return None
            finite = flat[torch.isfinite(flat)]
            lo = finite.min().item() if finite.numel() else float("nan")
            hi = finite.max().item() if finite.numel() else float("nan")
        return "{} detected in output of {}, shape: {}, {} pass (finite range {:.4g}..{:.4g})".format(
            what, module.__dict__.get("_nan_detector_name", "?"), tuple(tensor.shape), kind, lo, hi
        )

    def _scan(self, module, payload, kind):
        if self._reported[kind]:
            return
        items = payload if isinstance(payload, (tuple, list)) else [payload]
        for item in items:
            if isinstance(item, dict):
                item = next((v for v in item.values() if torch.is_tensor(v)), None)
            msg = self._describe(module, item, kind)
            if msg is not None:
                logger.warning(msg)
                self._reported[kind] = True
                return

    def fhook_fn(self, module, inp, output):
        self._scan(module, output, "forward")

    def bhook_fn(self, module, grad_in, grad_out):
        self._scan(module, grad_out, "backward")

    def close(self):
        for h in self.fhooks + self.bhooks:
            h.remove()
        self.fhooks, self.bhooks = [], []
