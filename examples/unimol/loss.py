"""Uni-Mol pre-training loss: masked-atom cross entropy + coordinate smooth-L1 + pair-distance
smooth-L1 + representation norm regularisers (weights from the model flags)."""
import torch
import torch.nn.functional as F

from unicore import metrics, utils
from unicore.losses import UnicoreLoss, register_loss


@register_loss("unimol")
class UniMolLoss(UnicoreLoss):
    def __init__(self, task):
        super().__init__(task)
        self.padding_idx = task.dictionary.pad()
        self.seed = task.seed
        self.dist_mean = 6.312581655060595
        self.dist_std = 3.3899264663911888

    def forward(self, model, sample, reduce=True):
        tgt_tokens = sample["target"]["tokens_target"]
        masked_tokens = tgt_tokens.ne(self.padding_idx)
        sample_size = masked_tokens.long().sum()
        utils.request_mask_index(masked_tokens)  # count read queued ahead of the encoder (consumed by the head below)
        logits, pred_dist, pred_coord, x_norm, delta_norm = model(
            **sample["net_input"], encoder_masked_tokens=masked_tokens
        )
        # one host read for the number of masked atoms, shared by the model head and every gather below
        midx = utils.mask_to_index(masked_tokens)
        target = tgt_tokens.reshape(-1).index_select(0, midx)
        token_loss = F.nll_loss(F.log_softmax(logits, dim=-1, dtype=torch.float32), target,
                                ignore_index=self.padding_idx, reduction="mean")
        loss = token_loss * self.args.masked_token_loss
        log = {"masked_token_loss": token_loss.data, "sample_size": 1, "bsz": tgt_tokens.size(0),
               "seq_len": tgt_tokens.size(1) * tgt_tokens.size(0)}
        if pred_coord is not None:
            coord_target = sample["target"]["coord_target"]
            coord_loss = F.smooth_l1_loss(pred_coord.reshape(-1, 3).index_select(0, midx).float(),
                                          coord_target.reshape(-1, 3).index_select(0, midx), reduction="mean", beta=1.0)
            loss = loss + coord_loss * self.args.masked_coord_loss
            log["masked_coord_loss"] = coord_loss.data
        if pred_dist is not None:
            L = tgt_tokens.size(1)
            dist_target = sample["target"]["distance_target"].reshape(-1, L).index_select(0, midx)
            non_pad = sample["net_input"]["src_tokens"].ne(self.padding_idx)
            # rows of masked atoms x non-padding columns, as a weight instead of a second boolean gather (which would
            # cost another host synchronisation for its element count)
            w = non_pad.unsqueeze(1).expand(-1, L, -1).reshape(-1, L).index_select(0, midx)
            pd = pred_dist.reshape(-1, L).index_select(0, midx).float()
            td = (dist_target.float() - self.dist_mean) / self.dist_std
            per_elem = F.smooth_l1_loss(pd, td, reduction="none", beta=1.0)
            dist_loss = (per_elem * w).sum() / w.sum().clamp(min=1)
            loss = loss + dist_loss * self.args.masked_dist_loss
            log["masked_dist_loss"] = dist_loss.data
        if self.args.x_norm_loss > 0 and x_norm is not None:
            loss = loss + self.args.x_norm_loss * x_norm
            log["x_norm_loss"] = x_norm.data
        if self.args.delta_pair_repr_norm_loss > 0 and delta_norm is not None:
            loss = loss + self.args.delta_pair_repr_norm_loss * delta_norm
            log["delta_pair_repr_norm_loss"] = delta_norm.data
        log["loss"] = loss.data
        return loss, 1, log

    @staticmethod
    def reduce_metrics(logging_outputs, split="valid") -> None:
        n = sum(log.get("sample_size", 0) for log in logging_outputs)
        bsz = sum(log.get("bsz", 0) for log in logging_outputs)
        metrics.log_scalar("loss", sum(log.get("loss", 0) for log in logging_outputs) / n, n, round=3)
        metrics.log_scalar("seq_len", sum(log.get("seq_len", 0) for log in logging_outputs) / bsz, 1, round=3)
        for key in ("masked_token_loss", "masked_coord_loss", "masked_dist_loss", "x_norm_loss", "delta_pair_repr_norm_loss"):
            if any(key in log for log in logging_outputs):
                metrics.log_scalar(key, sum(log.get(key, 0) for log in logging_outputs) / n, n, round=3)

    @staticmethod
    def logging_outputs_can_be_summed(is_train) -> bool:
        return True
