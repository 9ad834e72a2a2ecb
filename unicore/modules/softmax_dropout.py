"""``softmax_dropout`` functional (implementation: ``unicore_b200/ops/softmax_ops.py``)."""
from unicore.ops import softmax_dropout, softmax_dropout_with_logits  # noqa: F401


class SoftmaxDropoutFast:
    """Argument-order shim for code written against the reference autograd Function
    (``softmax_dropout.py:18-60``: ``apply(is_training, inputs, mask, bias, dropout_prob)``, in place)."""

    @staticmethod
    def apply(is_training, inputs, mask, bias, dropout_prob):
        return softmax_dropout(inputs, dropout_prob, is_training, mask=mask, bias=bias, inplace=True)
