"""Portable Uni-Mol model + loss (public Uni-Core API only, see the package docstring)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from unicore import metrics, utils
from unicore.losses import UnicoreLoss, register_loss
from unicore.models import BaseUnicoreModel, register_model, register_model_architecture
from unicore.modules import LayerNorm, TransformerEncoderLayer, init_bert_params

DEFAULTS = dict(
    encoder_layers=15, encoder_embed_dim=512, encoder_ffn_embed_dim=2048, encoder_attention_heads=64,
    dropout=0.1, emb_dropout=0.1, attention_dropout=0.1, activation_dropout=0.0, pooler_dropout=0.0,
    max_seq_len=512, activation_fn="gelu", pooler_activation_fn="tanh", post_ln=False,
    masked_token_loss=1.0, masked_coord_loss=5.0, masked_dist_loss=10.0, x_norm_loss=0.01, delta_pair_repr_norm_loss=0.01,
    gaussian_kernels=128,
)


def fill_defaults(args):
    for key, value in DEFAULTS.items():
        if getattr(args, key, None) is None:
            setattr(args, key, value)


class RadialBasis(nn.Module):
    """K Gaussians of an edge-type dependent affine map of the distance."""

    def __init__(self, kernels, edge_types):
        super().__init__()
        self.kernels = kernels
        self.means = nn.Embedding(1, kernels)
        self.stds = nn.Embedding(1, kernels)
        self.mul = nn.Embedding(edge_types, 1)
        self.bias = nn.Embedding(edge_types, 1)
        nn.init.uniform_(self.means.weight, 0, 3)
        nn.init.uniform_(self.stds.weight, 0, 3)
        nn.init.constant_(self.bias.weight, 0)
        nn.init.constant_(self.mul.weight, 1)

    def forward(self, dist, edge_type):
        scale, shift = self.mul(edge_type).type_as(dist), self.bias(edge_type).type_as(dist)
        z = (scale * dist.unsqueeze(-1) + shift).expand(-1, -1, -1, self.kernels).float()
        mean = self.means.weight.float().view(-1)
        std = self.stds.weight.float().view(-1).abs() + 1e-5
        norm = (2 * 3.14159) ** 0.5
        return (torch.exp(-0.5 * ((z - mean) / std) ** 2) / (norm * std)).type_as(self.means.weight)


class TwoLayerHead(nn.Module):
    def __init__(self, d_in, d_out, activation, d_hidden=None):
        super().__init__()
        self.linear1 = nn.Linear(d_in, d_hidden or d_in)
        self.linear2 = nn.Linear(d_hidden or d_in, d_out)
        self.act = utils.get_activation_fn(activation)

    def forward(self, x):
        return self.linear2(self.act(self.linear1(x)))


class AtomTypeHead(nn.Module):
    def __init__(self, dim, vocab, activation):
        super().__init__()
        self.dense = nn.Linear(dim, dim)
        self.act = utils.get_activation_fn(activation)
        self.layer_norm = LayerNorm(dim)
        self.weight = nn.Linear(dim, vocab, bias=False).weight
        self.bias = nn.Parameter(torch.zeros(vocab))

    def forward(self, features, masked=None):
        if masked is not None:
            features = features[masked, :]
        return F.linear(self.layer_norm(self.act(self.dense(features))), self.weight) + self.bias


class PairDistanceHead(nn.Module):
    def __init__(self, heads, activation):
        super().__init__()
        self.dense = nn.Linear(heads, heads)
        self.layer_norm = LayerNorm(heads)
        self.out_proj = nn.Linear(heads, 1)
        self.act = utils.get_activation_fn(activation)

    def forward(self, pair):
        b, n = pair.shape[0], pair.shape[1]
        d = self.out_proj(self.layer_norm(self.act(self.dense(pair)))).view(b, n, n)
        return 0.5 * (d + d.transpose(-1, -2))


class PairBiasEncoder(nn.Module):
    """Transformer encoder whose attention logits of layer i are the attention bias of layer i + 1."""

    def __init__(self, layers, dim, ffn_dim, heads, emb_dropout, dropout, attention_dropout, activation_dropout,
                 activation_fn, post_ln, head_norm):
        super().__init__()
        self.emb_dropout, self.heads = emb_dropout, heads
        self.emb_layer_norm = LayerNorm(dim)
        self.final_layer_norm = None if post_ln else LayerNorm(dim)
        self.final_head_layer_norm = LayerNorm(heads) if head_norm else None
        self.layers = nn.ModuleList([
            TransformerEncoderLayer(embed_dim=dim, ffn_embed_dim=ffn_dim, attention_heads=heads, dropout=dropout,
                                    attention_dropout=attention_dropout, activation_dropout=activation_dropout,
                                    activation_fn=activation_fn, post_ln=post_ln)
            for _ in range(layers)])

    @staticmethod
    def _excess_norm(t, tolerance=1.0):
        return F.relu((t.float().norm(dim=-1) - math.sqrt(t.shape[-1])).abs() - tolerance)

    def forward(self, emb, pair_bias, padding_mask):
        b, n = emb.shape[0], emb.shape[1]
        x = F.dropout(self.emb_layer_norm(emb), p=self.emb_dropout, training=self.training)
        keep = None
        if padding_mask is not None:
            keep = 1.0 - padding_mask.float()
            x = x * keep.unsqueeze(-1).type_as(x)
            # padded keys never receive attention: -inf in their columns of the bias, which then travels with the logits
            pair_bias = pair_bias.view(b, -1, n, n).masked_fill(padding_mask[:, None, None, :], float("-inf")).view(-1, n, n)
        first_bias = pair_bias
        for layer in self.layers:
            x, pair_bias, _ = layer(x, padding_mask=None, attn_bias=pair_bias, return_attn=True)
        x_norm = self._excess_norm(x)
        if keep is None:
            keep = torch.ones_like(x_norm)
        x_norm = ((keep * x_norm).sum(-1) / (1e-10 + keep.sum(-1))).mean()
        if self.final_layer_norm is not None:
            x = self.final_layer_norm(x)

        def pair_major(t):  # [B*H, N, N] -> [B, N, N, H], -inf -> 0
            t = t.masked_fill(t == float("-inf"), 0)
            return t.view(b, -1, n, n).permute(0, 2, 3, 1).contiguous()

        pair = pair_major(pair_bias)
        delta = pair - pair_major(first_bias)
        pair_keep = keep[..., None] * keep[..., None, :]
        d_norm = self._excess_norm(delta)
        delta_norm = ((pair_keep * d_norm).sum((-1, -2)) / (1e-10 + pair_keep.sum((-1, -2)))).mean()
        if self.final_head_layer_norm is not None:
            delta = self.final_head_layer_norm(delta)
        return x, pair, delta, x_norm, delta_norm


@register_model("unimol_portable")
class PortableUniMol(BaseUnicoreModel):
    @staticmethod
    def add_args(parser):
        for flag, kind in (("--encoder-layers", int), ("--encoder-embed-dim", int), ("--encoder-ffn-embed-dim", int),
                           ("--encoder-attention-heads", int), ("--emb-dropout", float), ("--dropout", float),
                           ("--attention-dropout", float), ("--activation-dropout", float), ("--pooler-dropout", float),
                           ("--max-seq-len", int), ("--masked-token-loss", float), ("--masked-dist-loss", float),
                           ("--masked-coord-loss", float), ("--x-norm-loss", float), ("--delta-pair-repr-norm-loss", float),
                           ("--gaussian-kernels", int)):
            parser.add_argument(flag, type=kind)
        parser.add_argument("--activation-fn", choices=utils.get_available_activation_fns())
        parser.add_argument("--pooler-activation-fn", choices=utils.get_available_activation_fns())
        parser.add_argument("--post-ln", type=utils.eval_bool)

    def __init__(self, args, dictionary):
        super().__init__()
        fill_defaults(args)
        self.args = args
        self.padding_idx = dictionary.pad()
        vocab, dim, heads = len(dictionary), args.encoder_embed_dim, args.encoder_attention_heads
        self.embed_tokens = nn.Embedding(vocab, dim, self.padding_idx)
        self.encoder = PairBiasEncoder(
            args.encoder_layers, dim, args.encoder_ffn_embed_dim, heads, args.emb_dropout, args.dropout,
            args.attention_dropout, args.activation_dropout, args.activation_fn, args.post_ln,
            head_norm=args.delta_pair_repr_norm_loss >= 0)
        self.lm_head = AtomTypeHead(dim, vocab, args.activation_fn) if args.masked_token_loss > 0 else None
        self.gbf = RadialBasis(args.gaussian_kernels, vocab * vocab)
        self.gbf_proj = TwoLayerHead(args.gaussian_kernels, heads, args.activation_fn)
        self.pair2coord_proj = TwoLayerHead(heads, 1, args.activation_fn) if args.masked_coord_loss > 0 else None
        self.dist_head = PairDistanceHead(heads, args.activation_fn) if args.masked_dist_loss > 0 else None
        self.apply(init_bert_params)

    @classmethod
    def build_model(cls, args, task):
        return cls(args, task.dictionary)

    def forward(self, src_tokens, src_distance, src_coord, src_edge_type, encoder_masked_tokens=None, **kwargs):
        padding_mask = src_tokens.eq(self.padding_idx)
        n = src_distance.size(-1)
        bias = self.gbf_proj(self.gbf(src_distance, src_edge_type)).permute(0, 3, 1, 2).contiguous().view(-1, n, n)
        x, pair, delta, x_norm, delta_norm = self.encoder(self.embed_tokens(src_tokens), bias, padding_mask)
        logits = coord = dist = None
        if self.lm_head is not None:
            logits = self.lm_head(x, encoder_masked_tokens)
        if self.pair2coord_proj is not None:
            atoms = (torch.sum(1 - padding_mask.type_as(x), dim=1) - 1).view(-1, 1, 1, 1)
            offsets = src_coord.unsqueeze(1) - src_coord.unsqueeze(2)
            coord = src_coord + torch.sum(offsets / atoms * self.pair2coord_proj(delta), dim=2)
        if self.dist_head is not None:
            dist = self.dist_head(pair)
        return logits, dist, coord, x_norm, delta_norm


@register_model_architecture("unimol_portable", "unimol_portable_base")
def portable_base(args):
    fill_defaults(args)


@register_loss("unimol_portable")
class PortableUniMolLoss(UnicoreLoss):
    DIST_MEAN, DIST_STD = 6.312581655060595, 3.3899264663911888

    def __init__(self, task):
        super().__init__(task)
        self.padding_idx = task.dictionary.pad()

    def forward(self, model, sample, reduce=True):
        a = self.args
        target = sample["target"]["tokens_target"]
        masked = target.ne(self.padding_idx)
        logits, dist, coord, x_norm, delta_norm = model(**sample["net_input"], encoder_masked_tokens=masked)
        token_loss = F.nll_loss(F.log_softmax(logits, dim=-1, dtype=torch.float32), target[masked],
                                ignore_index=self.padding_idx, reduction="mean")
        loss = token_loss * a.masked_token_loss
        log = {"masked_token_loss": token_loss.data, "sample_size": 1, "bsz": target.size(0),
               "seq_len": target.size(1) * target.size(0)}
        if coord is not None:
            c = F.smooth_l1_loss(coord[masked].view(-1, 3).float(), sample["target"]["coord_target"][masked].view(-1, 3),
                                 reduction="mean", beta=1.0)
            loss = loss + c * a.masked_coord_loss
            log["masked_coord_loss"] = c.data
        if dist is not None:
            rows = dist[masked, :]
            want = sample["target"]["distance_target"][masked]
            cols = sample["net_input"]["src_tokens"].ne(self.padding_idx).unsqueeze(1).expand(-1, target.size(1), -1)[masked]
            d = F.smooth_l1_loss(rows[cols].view(-1).float(), ((want[cols].view(-1).float() - self.DIST_MEAN) / self.DIST_STD),
                                 reduction="mean", beta=1.0)
            loss = loss + d * a.masked_dist_loss
            log["masked_dist_loss"] = d.data
        if a.x_norm_loss > 0 and x_norm is not None:
            loss = loss + a.x_norm_loss * x_norm
            log["x_norm_loss"] = x_norm.data
        if a.delta_pair_repr_norm_loss > 0 and delta_norm is not None:
            loss = loss + a.delta_pair_repr_norm_loss * delta_norm
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
