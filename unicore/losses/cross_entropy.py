"""``--loss cross_entropy``: sentence-level classification loss (interface and numerics of reference
``unicore/losses/cross_entropy.py:14-65``): fp32 log-softmax over the model output, NLL summed over the batch,
``sample_size`` = number of sentences; the logged ``loss`` is in bits (base 2) per sentence.

On CUDA the fp32 log-probabilities are never materialised: ``ops.softmax_cross_entropy`` (one fused kernel per
direction) produces the summed NLL directly.
"""
import math

import torch.nn.functional as F

from unicore import metrics, ops
from unicore.losses import UnicoreLoss, register_loss

_LN2 = math.log(2)


@register_loss("cross_entropy")
class CrossEntropyLoss(UnicoreLoss):
    def forward(self, model, sample, reduce=True):
        target = sample["target"]
        n_sentences = target.size(0)
        loss = self.compute_loss(model, model(**sample["net_input"]), sample, reduce=reduce)
        return loss, n_sentences, dict(loss=loss.data, bsz=n_sentences, sample_size=n_sentences)

    def compute_loss(self, model, net_output, sample, reduce=True):
        logits = net_output.reshape(-1, net_output.size(-1))
        target = sample["target"].reshape(-1)
        if reduce and ops.use_native(logits, target):
            return ops.softmax_cross_entropy(logits, target)
        return F.nll_loss(F.log_softmax(logits.float(), dim=-1), target, reduction="sum" if reduce else "none")

    @staticmethod
    def logging_outputs_can_be_summed(is_train) -> bool:
        return True

    @staticmethod
    def reduce_metrics(logging_outputs, split="valid") -> None:
        totals = {key: sum(entry.get(key, 0) for entry in logging_outputs) for key in ("loss", "sample_size")}
        metrics.log_scalar("loss", totals["loss"] / totals["sample_size"] / _LN2, totals["sample_size"], round=3)
