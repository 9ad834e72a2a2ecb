"""Masked-LM loss: only positions whose target is not padding contribute; the model is asked to
project only those positions (``masked_tokens``).  ``sample_size`` is the (device-resident) count
of masked tokens, which is what gradients are normalised by.
Parity: reference ``unicore/losses/masked_lm.py:13-67``.

B200 path: the fp32 log-softmax + NLL over the ``[n_masked, vocab]`` logits is one fused kernel
(``unicore.ops.softmax_cross_entropy``) that never materialises the fp32 log-probabilities.
"""
import math

from unicore import metrics, ops, utils
from unicore.losses import UnicoreLoss, register_loss


@register_loss("masked_lm")
class MaskedLMLoss(UnicoreLoss):
    def __init__(self, task):
        super().__init__(task)
        self.padding_idx = task.dictionary.pad()

    def forward(self, model, sample, reduce=True):
        target = sample["target"]
        masked_tokens = target.ne(self.padding_idx)
        sample_size = masked_tokens.int().sum()
        # start resolving the masked positions NOW (asynchronous count read, consumed by the LM head and the target
        # gather below): the copy is queued ahead of the encoder, so nobody waits for it
        utils.request_mask_index(masked_tokens)
        logits = model(**sample["net_input"], masked_tokens=masked_tokens)
        target = target.reshape(-1).index_select(0, utils.mask_to_index(masked_tokens))
        loss = ops.softmax_cross_entropy(logits, target, ignore_index=self.padding_idx)
        logging_output = {
            "loss": loss.data,
            "bsz": sample["target"].size(0),
            "sample_size": sample_size,
            "seq_len": sample["target"].size(1) * sample["target"].size(0),
        }
        return loss, sample_size, logging_output

    @staticmethod
    def reduce_metrics(logging_outputs, split="valid") -> None:
        loss_sum = sum(log.get("loss", 0) for log in logging_outputs)
        bsz = sum(log.get("bsz", 0) for log in logging_outputs)
        sample_size = sum(log.get("sample_size", 0) for log in logging_outputs)
        seq_len = sum(log.get("seq_len", 0) for log in logging_outputs)
        metrics.log_scalar("loss", loss_sum / sample_size / math.log(2), sample_size, round=3)
        metrics.log_scalar("seq_len", seq_len / bsz, 1, round=3)

    @staticmethod
    def logging_outputs_can_be_summed(is_train) -> bool:
        return True
