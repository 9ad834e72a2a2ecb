"""Dynamic loss scaling state machine for fp16 training.

Semantics (reference ``unicore/optim/dynamic_loss_scaler.py:8-71``): grow x``scale_factor`` whenever
``scale_window`` updates have passed since the last overflow; on overflow (non-finite total grad
norm) shrink - if the overflow fraction since the last rescale reaches ``tolerance`` - abort with
``FloatingPointError`` once the scale would fall to ``min_loss_scale``, else raise ``OverflowError``
which the trainer turns into a skipped update.
"""
import math


class DynamicLossScaler(object):
    def __init__(self, init_scale=2.0 ** 15, scale_factor=2.0, scale_window=2000, tolerance=0.0,
                 threshold=None, min_loss_scale=1e-4):
        self.loss_scale = init_scale
        self.scale_factor = scale_factor
        self.scale_window = scale_window
        self.tolerance = tolerance
        self.threshold = threshold
        self.min_loss_scale = min_loss_scale
        self._iter = 0
        self._last_overflow_iter = -1
        self._last_rescale_iter = -1
        self._overflows_since_rescale = 0

    def scale(self, outputs):
        return self.loss_scale * outputs

    def update(self):
        """Call after every successful (non-overflowed) optimizer step."""
        if (self._iter - self._last_overflow_iter) % self.scale_window == 0:
            self.loss_scale *= self.scale_factor
            self._last_rescale_iter = self._iter
        self._iter += 1

    def _decrease_loss_scale(self):
        self.loss_scale /= self.scale_factor
        if self.threshold is not None:
            self.loss_scale = max(self.loss_scale, self.threshold)

    @staticmethod
    def has_overflow(grad_norm) -> bool:
        value = float(grad_norm)
        return math.isinf(value) or math.isnan(value)

    def check_overflow(self, grad_norm):
        if not self.has_overflow(grad_norm):
            return
        previous = self.loss_scale
        since_rescale = self._iter - self._last_rescale_iter
        self._last_overflow_iter = self._iter
        self._overflows_since_rescale += 1
        if self._overflows_since_rescale / float(since_rescale) >= self.tolerance:
            self._decrease_loss_scale()
            self._last_rescale_iter = self._iter
            self._overflows_since_rescale = 0
        if self.loss_scale <= self.min_loss_scale:
            self.loss_scale = previous
            raise FloatingPointError(
                "Minimum loss scale reached ({}). Your loss is probably exploding. Try lowering the learning "
                "rate, using gradient clipping or increasing the batch size.".format(self.min_loss_scale)
            )
        self._iter += 1
        raise OverflowError("setting loss scale to: " + str(self.loss_scale))
