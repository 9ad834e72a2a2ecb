"""Transformer encoder / decoder stacks (BERT/XLM style, pre- or post-LN, T5 relative positions).

Parity: reference ``unicore/modules/transformer_encoder_layer.py:15``,
``transformer_encoder.py:16-162`` (``init_bert_params``, ``relative_position_bucket``,
``TransformerEncoder``), ``transformer_decoder_layer.py:15`` and ``transformer_decoder.py:19-169``.
Same constructor arguments, parameter names (=> identical ``state_dict`` keys) and forward
signatures, including ``return_attn`` on the encoder layer.

B200 execution plan for one layer (post-LN shown; pre-LN analogous):
  in_proj GEMM -> fused tcgen05 attention (bias tile + key padding + dropout inside) ->
  out_proj GEMM -> ONE kernel: bias + dropout + residual + LayerNorm ->
  fc1 GEMM -> ONE kernel: bias + GELU -> fc2 GEMM -> ONE kernel: bias + dropout + residual + LN.
The relative-position bias stays ``[1, H, L, L]`` (never repeated over the batch) and the key
padding mask stays ``[B, L]`` (never merged into a ``[B*H, L, L]`` tensor of ``-inf``) whenever the
fused attention kernel is usable; otherwise the reference's merged-mask formulation is used.
"""
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from unicore import ops, utils

from .attention import CrossMultiheadAttention, SelfMultiheadAttention
from .norm import LayerNorm


# ------------------------------------------------------------------------------------------------
# initialisation + relative positions
# ------------------------------------------------------------------------------------------------
def init_bert_params(module):
    """N(0, 0.02) for Linear/Embedding weights, zero biases and padding rows.

    Values are drawn from the CPU generator (like the reference, so a given seed yields the same
    initial weights) but in one shot per tensor in the tensor's own dtype/device afterwards.
    """
    if not getattr(module, "can_global_init", True):
        return

    def draw(weight):
        fresh = torch.empty(weight.shape, dtype=torch.float32, device="cpu").normal_(mean=0.0, std=0.02)
        weight.copy_(fresh.to(device=weight.device, dtype=weight.dtype))

    with torch.no_grad():
        if isinstance(module, nn.Linear):
            draw(module.weight.data)
            if module.bias is not None:
                module.bias.data.zero_()
        if isinstance(module, nn.Embedding):
            draw(module.weight.data)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()


def relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
    """Signed T5-style bucketing: exact buckets for small |d|, log-spaced up to ``max_distance``."""
    sign = torch.sign(relative_position)
    half = num_buckets // 2
    dist = torch.abs(relative_position)
    exact = half // 2
    top = half - 1 - exact
    log_bucket = exact + torch.ceil(
        torch.log(dist.float() / exact) / math.log((max_distance - 1) / exact) * top
    ).long()
    log_bucket = torch.clamp(log_bucket, max=half - 1)
    return torch.where(dist < exact, dist, log_bucket) * sign


def _build_rp_bucket(max_seq_len, bins, max_rel_pos):
    pos = torch.arange(max_seq_len, dtype=torch.long)
    bucket = relative_position_bucket(pos[None, :] - pos[:, None], num_buckets=bins, max_distance=max_rel_pos)
    return bucket - bucket.min()


def fill_with_neg_inf(t):
    return t.float().fill_(float("-inf")).type_as(t)


def bulid_future_mask(seq_len):  # (sic) name kept for API compatibility
    return torch.triu(fill_with_neg_inf(torch.zeros([seq_len, seq_len])), 1)


# ------------------------------------------------------------------------------------------------
# shared block pieces
# ------------------------------------------------------------------------------------------------
def _linear_no_bias(layer: nn.Linear, x):
    """GEMM only; the bias is applied by the fused epilogue kernel that follows."""
    return ops.linear(x, layer.weight, None)


class _BlockMixin:
    """Residual/FFN plumbing shared by encoder and decoder layers."""

    def _fused_ok(self, x):
        return ops.use_native(x) and x.dtype in (torch.float16, torch.bfloat16) and self._gelu

    def _attn_residual(self, attn_module, norm, residual, attn_in, **kw):
        """post-LN: LN(residual + drop(attn(x)));  pre-LN: residual + drop(attn(LN(x)))."""
        return_attn = kw.pop("return_attn", False)
        h = attn_in if self.post_ln else norm(attn_in)
        extra = None
        if isinstance(attn_module, SelfMultiheadAttention):
            o, logits, probs = attn_module.attend(h, kw.get("key_padding_mask"), kw.get("attn_bias"), return_attn)
            if return_attn:
                extra = (logits, probs)
        else:
            o = attn_module.attend(h, kw["key"], kw["value"], kw.get("key_padding_mask"), kw.get("attn_bias"))
        proj = attn_module.out_proj
        if self.post_ln and self._fused_ok(o):
            # out_proj bias + dropout + residual + LayerNorm in one pass over the GEMM output
            out = ops.bias_dropout_add_layer_norm(
                ops.linear(o, proj.weight, None), proj.bias, residual, norm.weight, norm.bias,
                self.dropout, norm.eps, self.training,
            )
            return out, extra
        out = residual + F.dropout(ops.linear(o, proj.weight, proj.bias), p=self.dropout, training=self.training)
        if self.post_ln:
            out = norm(out)
        return out, extra

    def _ffn_residual(self, x):
        residual = x
        if self._fused_ok(x):
            h = x if self.post_ln else self.final_layer_norm(x)
            h = ops.bias_gelu(_linear_no_bias(self.fc1, h), self.fc1.bias)
            h = F.dropout(h, p=self.activation_dropout, training=self.training)
            if self.post_ln:
                return ops.bias_dropout_add_layer_norm(
                    _linear_no_bias(self.fc2, h), self.fc2.bias, residual,
                    self.final_layer_norm.weight, self.final_layer_norm.bias,
                    self.dropout, self.final_layer_norm.eps, self.training,
                )
            h = ops.linear(h, self.fc2.weight, self.fc2.bias)
            return residual + F.dropout(h, p=self.dropout, training=self.training)
        h = x if self.post_ln else self.final_layer_norm(x)
        h = self.activation_fn(self.fc1(h))
        h = F.dropout(h, p=self.activation_dropout, training=self.training)
        h = self.fc2(h)
        h = residual + F.dropout(h, p=self.dropout, training=self.training)
        return self.final_layer_norm(h) if self.post_ln else h


class TransformerEncoderLayer(nn.Module, _BlockMixin):
    def __init__(
        self,
        embed_dim: int = 768,
        ffn_embed_dim: int = 3072,
        attention_heads: int = 8,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.0,
        activation_fn: str = "gelu",
        post_ln=False,
    ) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.activation_dropout = activation_dropout
        self.activation_fn = utils.get_activation_fn(activation_fn)
        self._gelu = activation_fn == "gelu"
        self.self_attn = SelfMultiheadAttention(self.embed_dim, attention_heads, dropout=attention_dropout)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.fc1 = nn.Linear(self.embed_dim, ffn_embed_dim)
        self.fc2 = nn.Linear(ffn_embed_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.post_ln = post_ln

    def forward(
        self,
        x: torch.Tensor,
        attn_bias: Optional[torch.Tensor] = None,
        padding_mask: Optional[torch.Tensor] = None,
        return_attn: bool = False,
    ) -> torch.Tensor:
        x, extra = self._attn_residual(
            self.self_attn, self.self_attn_layer_norm, x, x,
            key_padding_mask=padding_mask, attn_bias=attn_bias, return_attn=return_attn,
        )
        x = self._ffn_residual(x)
        if not return_attn:
            return x
        attn_weights, attn_probs = extra
        return x, attn_weights, attn_probs


class TransformerDecoderLayer(nn.Module, _BlockMixin):
    def __init__(
        self,
        embed_dim: int = 768,
        ffn_embed_dim: int = 3072,
        attention_heads: int = 8,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.0,
        activation_fn: str = "gelu",
        post_ln=False,
    ) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        self.attention_dropout = attention_dropout
        self.dropout = dropout
        self.activation_dropout = activation_dropout
        self.activation_fn = utils.get_activation_fn(activation_fn)
        self._gelu = activation_fn == "gelu"
        self.self_attn = SelfMultiheadAttention(self.embed_dim, attention_heads, dropout=attention_dropout)
        self.self_attn_layer_norm = LayerNorm(self.embed_dim)
        self.encoder_attn = CrossMultiheadAttention(self.embed_dim, attention_heads, dropout=attention_dropout)
        self.encoder_attn_layer_norm = LayerNorm(self.embed_dim)
        self.fc1 = nn.Linear(self.embed_dim, ffn_embed_dim)
        self.fc2 = nn.Linear(ffn_embed_dim, self.embed_dim)
        self.final_layer_norm = LayerNorm(self.embed_dim)
        self.post_ln = post_ln

    def forward(
        self,
        x: torch.Tensor,
        encoder_out: torch.Tensor = None,
        attn_bias: Optional[torch.Tensor] = None,
        padding_mask: Optional[torch.Tensor] = None,
        encoder_attn_bias: Optional[torch.Tensor] = None,
        encoder_padding_mask: Optional[torch.Tensor] = None,
    ) -> torch.Tensor:
        x, _ = self._attn_residual(
            self.self_attn, self.self_attn_layer_norm, x, x, key_padding_mask=padding_mask, attn_bias=attn_bias
        )
        if encoder_out is not None:
            x, _ = self._attn_residual(
                self.encoder_attn, self.encoder_attn_layer_norm, x, x,
                key=encoder_out, value=encoder_out,
                key_padding_mask=encoder_padding_mask, attn_bias=encoder_attn_bias,
            )
        return self._ffn_residual(x)


# ------------------------------------------------------------------------------------------------
# stacks
# ------------------------------------------------------------------------------------------------
class _StackMixin:
    def _init_rel_pos(self, rel_pos, rel_pos_bins, max_rel_pos):
        self.rel_pos = rel_pos
        if not rel_pos:
            return
        if rel_pos_bins % 2 != 0:
            raise ValueError("rel_pos_bins must be even")
        self.rel_pos_bins = rel_pos_bins
        self.max_rel_pos = max_rel_pos
        self.relative_attention_bias = nn.Embedding(rel_pos_bins, self.attention_heads)
        self.rp_bucket = _build_rp_bucket(self.max_seq_len, rel_pos_bins, max_rel_pos)

    def get_rel_pos_bias(self, x):
        """``[H, L, L]`` bias from the bucket table (positions assumed ordered)."""
        if self.rp_bucket.device != x.device:
            self.rp_bucket = self.rp_bucket.to(x.device)
        seq_len = x.size(1)
        weight = self.relative_attention_bias.weight
        one_hot_bytes = self.rel_pos_bins * seq_len * seq_len * 2
        if weight.is_cuda and weight.dtype in (torch.float16, torch.bfloat16) and one_hot_bytes <= (64 << 20):
            # The bucket table is a constant: with its one-hot matrix E [bins, L*L] (cached) the lookup is the GEMM
            # W^T E - it lands directly in [H, L, L] order (no permute copy), and autograd's backward is the GEMM
            # dBias E^T with fp32 accumulation instead of ATen's embedding backward (a radix sort of L*L indices plus
            # a segmented reduction: ~0.2 ms and 12 launches per BERT-base step).  Same values: each output element is
            # one table entry times 1.
            return (weight.t() @ self._bucket_one_hot(seq_len, weight)).view(-1, seq_len, seq_len)
        values = F.embedding(self.rp_bucket[:seq_len, :seq_len], weight)
        return values.permute(2, 0, 1).contiguous()

    def _bucket_one_hot(self, seq_len, like):
        key = (seq_len, like.device, like.dtype)
        cache = getattr(self, "_rp_one_hot", None)
        if cache is None or cache[0] != key:
            idx = self.rp_bucket[:seq_len, :seq_len].reshape(1, -1)
            one_hot = torch.zeros(self.rel_pos_bins, idx.numel(), dtype=like.dtype, device=like.device)
            one_hot.scatter_(0, idx, 1.0)
            self._rp_one_hot = cache = (key, one_hot)
        return cache[1]

    def _embed_prologue(self, emb, padding_mask):
        x = self.emb_layer_norm(emb)
        x = F.dropout(x, p=self.emb_dropout, training=self.training)
        if padding_mask is not None:
            x = x * (1 - padding_mask.unsqueeze(-1).type_as(x))
        return x

    def _compose_bias(self, x, attn_mask, padding_mask, future_mask=None):
        """Combine user mask, rel-pos bias, causal mask and key padding into what layers receive.

        Returns ``(attn_bias, key_padding_mask)``.  Fast form: bias ``[1, H, L, L]`` + separate
        padding mask (consumed by the fused attention kernel).  Reference form: everything merged
        into ``[B*H, L, L]`` with ``-inf`` at padded keys.
        """
        bsz, seq_len = x.size(0), x.size(1)
        heads = self.attention_heads
        rel = self.get_rel_pos_bias(x) if self.rel_pos else None  # [H, L, L]
        if attn_mask is None:
            bias = rel.unsqueeze(0) if rel is not None else None  # [1, H, L, L]
            if future_mask is not None:
                fm = future_mask[:seq_len, :seq_len]
                bias = fm.view(1, 1, seq_len, seq_len).expand(1, heads, seq_len, seq_len) if bias is None else bias + fm
            fused_capable = x.is_cuda and ops.USE_NATIVE and x.dtype in (torch.float16, torch.bfloat16) \
                and (self.embed_dim // heads) == 64
            if bias is None or fused_capable or padding_mask is None:
                return bias, padding_mask
            merged = bias.expand(bsz, heads, seq_len, seq_len).clone()
        else:
            merged = attn_mask.view(bsz, -1, seq_len, seq_len)
            if rel is not None:
                merged = merged + rel.unsqueeze(0)
            if future_mask is not None:
                merged = merged + future_mask[:seq_len, :seq_len]
            if padding_mask is None:
                return merged.reshape(-1, seq_len, seq_len), None
            if merged.shape[1] != heads:
                merged = merged.expand(bsz, heads, seq_len, seq_len)
            merged = merged.clone() if merged.data_ptr() == attn_mask.data_ptr() else merged
        merged = merged.masked_fill(padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf"))
        return merged.reshape(-1, seq_len, seq_len), None


class TransformerEncoder(nn.Module, _StackMixin):
    def __init__(
        self,
        encoder_layers: int = 6,
        embed_dim: int = 768,
        ffn_embed_dim: int = 3072,
        attention_heads: int = 8,
        emb_dropout: float = 0.1,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.0,
        max_seq_len: int = 256,
        activation_fn: str = "gelu",
        rel_pos: bool = True,
        rel_pos_bins: int = 32,
        max_rel_pos: int = 128,
        post_ln: bool = False,
    ) -> None:
        super().__init__()
        self.emb_dropout = emb_dropout
        self.max_seq_len = max_seq_len
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        self.emb_layer_norm = LayerNorm(self.embed_dim)
        self.final_layer_norm = None if post_ln else LayerNorm(self.embed_dim)
        self.layers = nn.ModuleList(
            [
                TransformerEncoderLayer(
                    embed_dim=self.embed_dim,
                    ffn_embed_dim=ffn_embed_dim,
                    attention_heads=attention_heads,
                    dropout=dropout,
                    attention_dropout=attention_dropout,
                    activation_dropout=activation_dropout,
                    activation_fn=activation_fn,
                    post_ln=post_ln,
                )
                for _ in range(encoder_layers)
            ]
        )
        self._init_rel_pos(rel_pos, rel_pos_bins, max_rel_pos)

    def forward(
        self,
        emb: torch.Tensor,
        attn_mask: Optional[torch.Tensor] = None,
        padding_mask: Optional[torch.Tensor] = None,
    ) -> torch.Tensor:
        x = self._embed_prologue(emb, padding_mask)
        attn_bias, key_padding = self._compose_bias(x, attn_mask, padding_mask)
        for layer in self.layers:
            x = layer(x, padding_mask=key_padding, attn_bias=attn_bias)
        if self.final_layer_norm is not None:
            x = self.final_layer_norm(x)
        return x


class TransformerDecoder(nn.Module, _StackMixin):
    def __init__(
        self,
        decoder_layers: int = 6,
        embed_dim: int = 768,
        ffn_embed_dim: int = 3072,
        attention_heads: int = 8,
        emb_dropout: float = 0.1,
        dropout: float = 0.1,
        attention_dropout: float = 0.1,
        activation_dropout: float = 0.0,
        max_seq_len: int = 256,
        activation_fn: str = "gelu",
        rel_pos: bool = True,
        rel_pos_bins: int = 32,
        max_rel_pos: int = 128,
        post_ln: bool = False,
        auto_regressive: bool = True,
    ) -> None:
        super().__init__()
        self.emb_dropout = emb_dropout
        self.max_seq_len = max_seq_len
        self.embed_dim = embed_dim
        self.attention_heads = attention_heads
        self.emb_layer_norm = LayerNorm(self.embed_dim)
        self.auto_regressive = auto_regressive
        self._future_mask = bulid_future_mask(self.max_seq_len) if auto_regressive else None
        self.final_layer_norm = None if post_ln else LayerNorm(self.embed_dim)
        self.layers = nn.ModuleList(
            [
                TransformerDecoderLayer(
                    embed_dim=self.embed_dim,
                    ffn_embed_dim=ffn_embed_dim,
                    attention_heads=attention_heads,
                    dropout=dropout,
                    attention_dropout=attention_dropout,
                    activation_dropout=activation_dropout,
                    activation_fn=activation_fn,
                    post_ln=post_ln,
                )
                for _ in range(decoder_layers)
            ]
        )
        self._init_rel_pos(rel_pos, rel_pos_bins, max_rel_pos)

    def _future(self, x):
        if not self.auto_regressive:
            return None
        if self._future_mask.device != x.device or self._future_mask.dtype != x.dtype:
            self._future_mask = self._future_mask.to(device=x.device, dtype=x.dtype)
        return self._future_mask

    def get_future_mask(self, x, attn_mask):
        """Reference-format helper: returns ``[B*H, L, L]`` mask including the causal part."""
        if not self.auto_regressive:
            return attn_mask
        fm = self._future(x)[: x.size(1), : x.size(1)]
        if attn_mask is None:
            return fm.contiguous().unsqueeze(0).repeat(x.size(0) * self.attention_heads, 1, 1)
        return attn_mask + fm

    def forward(
        self,
        emb,
        encoder_out: Optional[torch.Tensor] = None,
        padding_mask: Optional[torch.Tensor] = None,
        encoder_padding_mask: Optional[torch.Tensor] = None,
        attn_mask: Optional[torch.Tensor] = None,
        encoder_attn_mask: Optional[torch.Tensor] = None,
    ) -> torch.Tensor:
        x = self._embed_prologue(emb, padding_mask)
        attn_bias, key_padding = self._compose_bias(x, attn_mask, padding_mask, future_mask=self._future(x))
        for layer in self.layers:
            x = layer(
                x,
                encoder_out=encoder_out,
                padding_mask=key_padding,
                attn_bias=attn_bias,
                encoder_padding_mask=encoder_padding_mask,
                encoder_attn_bias=encoder_attn_mask,
            )
        if self.final_layer_norm is not None:
            x = self.final_layer_norm(x)
        return x
