"""Dynamic loss scaling for fp16 training.

Behaviour (that of reference ``unicore/optim/dynamic_loss_scaler.py:8-71``, re-expressed with explicit
"updates since ..." counters instead of absolute iteration indices):

* every ``scale_window`` consecutive clean updates the scale is multiplied by ``scale_factor``;
* an overflow (non-finite total gradient norm) counts towards the overflow fraction since the last rescale; once
  that fraction reaches ``tolerance`` the scale is divided by ``scale_factor`` (never below ``threshold``);
* if the scale would drop to ``min_loss_scale`` the run is aborted with ``FloatingPointError``;
* otherwise ``OverflowError`` tells the trainer to skip the update.
"""
import math


class DynamicLossScaler(object):
    def __init__(self, init_scale=2.0 ** 15, scale_factor=2.0, scale_window=2000, tolerance=0.0,
                 threshold=None, min_loss_scale=1e-4):
        self.loss_scale = init_scale
        self.scale_factor, self.scale_window = scale_factor, scale_window
        self.tolerance, self.threshold, self.min_loss_scale = tolerance, threshold, min_loss_scale
        # counters in units of calls to update()/check_overflow(); the reference's indices start at
        # iter = 0, last_overflow = last_rescale = -1, i.e. both distances start at 1
        self._since_overflow = 1
        self._since_rescale = 1
        self._overflows_since_rescale = 0

    # -- used by the optimizer ---------------------------------------------------------------------------
    def scale(self, outputs):
        return self.loss_scale * outputs

    def update(self):
        """A parameter update went through without overflow."""
        if self._since_overflow % self.scale_window == 0:
            self.loss_scale *= self.scale_factor
            self._since_rescale = 0
        self._tick()

    def check_overflow(self, grad_norm):
        """Raise ``OverflowError`` (skip this update) or ``FloatingPointError`` (give up) on a non-finite norm."""
        if not self.has_overflow(grad_norm):
            return
        before = self.loss_scale
        self._overflows_since_rescale += 1
        overflow_fraction = self._overflows_since_rescale / float(self._since_rescale)
        self._since_overflow = 0
        if overflow_fraction >= self.tolerance:
            self._shrink()
            self._since_rescale = 0
            self._overflows_since_rescale = 0
        if self.loss_scale <= self.min_loss_scale:
            self.loss_scale = before  # keep the last usable scale for the error report / a restart
            raise FloatingPointError(
                "Minimum loss scale reached ({}). Your loss is probably exploding. Try lowering the learning "
                "rate, using gradient clipping or increasing the batch size.".format(self.min_loss_scale)
            )
        self._tick()
        raise OverflowError("setting loss scale to: " + str(self.loss_scale))

    @staticmethod
    def has_overflow(grad_norm) -> bool:
        return not math.isfinite(float(grad_norm))

    # -- internals -----------------------------------------------------------------------------------------
    def _tick(self):
        self._since_overflow += 1
        self._since_rescale += 1

    def _shrink(self):
        smaller = self.loss_scale / self.scale_factor
        self.loss_scale = smaller if self.threshold is None else max(smaller, self.threshold)
