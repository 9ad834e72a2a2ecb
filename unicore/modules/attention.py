"""Multi-head attention modules.

Parity: reference ``unicore/modules/multihead_attention.py`` - ``SelfMultiheadAttention:12``
(packed ``in_proj``, ``scaling = (head_dim * scaling_factor) ** -0.5``, additive ``attn_bias``,
boolean ``key_padding_mask``, ``return_attn`` yielding ``(out, pre-softmax logits incl. bias,
probabilities)``) and ``CrossMultiheadAttention:121``.

B200 path: unless ``return_attn`` is requested, the whole ``QK^T -> +bias/mask -> softmax ->
dropout -> @V`` chain runs in the tcgen05 kernel behind ``ops.fused_attention``; q/k/v are strided
views of the projection output (no transpose/contiguous copies) and the bias is *not* expanded over
the batch.  With ``return_attn`` the materialised path (bmm + ``softmax_dropout`` kernel) is used.
"""
from typing import Optional

import torch
from torch import Tensor, nn

from unicore import ops


def _bias_as_4d(attn_bias: Optional[Tensor], bsz: int, heads: int, tgt: int, src: int) -> Optional[Tensor]:
    """Normalise the accepted bias layouts to ``[1|B, H, tgt, src]``."""
    if attn_bias is None:
        return None
    if attn_bias.dim() == 4:
        return attn_bias
    if attn_bias.dim() == 3:
        lead = attn_bias.shape[0]
        if lead == bsz * heads:
            return attn_bias.view(bsz, heads, tgt, src)
        if lead == heads:
            return attn_bias.view(1, heads, tgt, src)
    if attn_bias.dim() == 2:
        return attn_bias.view(1, 1, tgt, src).expand(1, heads, tgt, src)
    raise ValueError("unsupported attn_bias shape {}".format(tuple(attn_bias.shape)))


def _materialised_attention(qh, kh, vh, key_padding_mask, attn_bias4, dropout, training, return_attn):
    """qh, kh, vh: head-major ``[B, H, L, D]`` (contiguous, query already scaled).

    Returns ``(out [B, Lq, H*D], logits?, probs?)``; scores are materialised (``return_attn`` / head dims the
    tcgen05 kernel does not cover).  Parity: reference ``multihead_attention.py:78-118``.
    """
    bsz, heads, tgt_len, dim = qh.shape
    src_len = kh.shape[2]
    w4 = torch.bmm(qh.view(bsz * heads, tgt_len, dim), kh.view(bsz * heads, src_len, dim).transpose(1, 2))
    w4 = w4.view(bsz, heads, tgt_len, src_len)
    pad = None
    if key_padding_mask is not None:
        # additive [B, 1, 1, K] mask: consumed by the softmax kernel's broadcast instead of a masked_fill_ pass
        pad = torch.zeros(bsz, 1, 1, src_len, dtype=w4.dtype, device=w4.device)
        pad.masked_fill_(key_padding_mask[:, None, None, :].to(torch.bool), float("-inf"))
    bias = attn_bias4.to(w4.dtype) if attn_bias4 is not None else None
    if not return_attn:
        attn = ops.softmax_dropout(w4, dropout, training, mask=pad, bias=bias)
        logits = probs = None
    else:
        # logits (scores + mask + bias) are an output: the pair representation of the next layer
        attn, logits = ops.softmax_dropout_with_logits(w4, dropout, training, mask=pad, bias=bias)
        logits = logits.view(bsz * heads, tgt_len, src_len)
        probs = attn.view(bsz * heads, tgt_len, src_len)
    out = torch.bmm(attn.view(bsz * heads, tgt_len, src_len), vh.view(bsz * heads, src_len, dim))
    return ops.merge_heads(out.view(bsz, heads, tgt_len, dim)), logits, probs


class SelfMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.1, bias=True, scaling_factor=1):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise ValueError("embed_dim must be divisible by num_heads")
        self.scaling = (self.head_dim * scaling_factor) ** -0.5
        self.in_proj = nn.Linear(embed_dim, embed_dim * 3, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

    def attend(self, query, key_padding_mask=None, attn_bias=None, return_attn=False):
        """Everything up to (not including) ``out_proj``: returns ``(o [B, L, E], logits, probs)``."""
        bsz, tgt_len, embed_dim = query.size()
        if embed_dim != self.embed_dim:
            raise ValueError("query dim {} != embed_dim {}".format(embed_dim, self.embed_dim))
        if key_padding_mask is not None and key_padding_mask.dim() == 0:
            key_padding_mask = None
        qkv = ops.linear(query, self.in_proj.weight, self.in_proj.bias)
        qkv = qkv.view(bsz, tgt_len, 3, self.num_heads, self.head_dim)
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]  # strided views [B, L, H, D]
        bias4 = _bias_as_4d(attn_bias, bsz, self.num_heads, tgt_len, tgt_len)
        if not return_attn and ops.fused_attention_supported(q, k, v, bias4, key_padding_mask):
            o = ops.fused_attention_qkvpacked(
                qkv, bias=bias4, key_padding_mask=key_padding_mask,
                dropout_p=self.dropout, training=self.training, scale=self.scaling,
            ).reshape(bsz, tgt_len, embed_dim)
            return o, None, None
        qh, kh, vh = ops.split_heads(qkv.view(bsz, tgt_len, 3 * embed_dim), 3, self.num_heads, self.scaling)
        return _materialised_attention(qh, kh, vh, key_padding_mask, bias4, self.dropout, self.training, return_attn)

    def forward(
        self,
        query,
        key_padding_mask: Optional[Tensor] = None,
        attn_bias: Optional[Tensor] = None,
        return_attn: bool = False,
    ) -> Tensor:
        o, logits, probs = self.attend(query, key_padding_mask, attn_bias, return_attn)
        o = self.out_proj(o)
        return (o, logits, probs) if return_attn else o


class CrossMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0.1, bias=True, scaling_factor=1):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise ValueError("embed_dim must be divisible by num_heads")
        self.scaling = (self.head_dim * scaling_factor) ** -0.5
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)

    def attend(self, query, key, value, key_padding_mask=None, attn_bias=None):
        """Everything up to (not including) ``out_proj``."""
        bsz, tgt_len, embed_dim = query.size()
        if embed_dim != self.embed_dim:
            raise ValueError("query dim {} != embed_dim {}".format(embed_dim, self.embed_dim))
        src_len = key.size(1)
        if key_padding_mask is not None and key_padding_mask.dim() == 0:
            key_padding_mask = None
        q = self.q_proj(query).view(bsz, tgt_len, self.num_heads, self.head_dim)
        k = self.k_proj(key).view(bsz, src_len, self.num_heads, self.head_dim)
        v = self.v_proj(value).view(bsz, src_len, self.num_heads, self.head_dim)
        bias4 = _bias_as_4d(attn_bias, bsz, self.num_heads, tgt_len, src_len)
        if ops.fused_attention_supported(q, k, v, bias4, key_padding_mask):
            return ops.fused_attention(
                q, k, v, bias=bias4, key_padding_mask=key_padding_mask,
                dropout_p=self.dropout, training=self.training, scale=self.scaling,
            ).reshape(bsz, tgt_len, embed_dim)
        (qh,) = ops.split_heads(q.view(bsz, tgt_len, embed_dim), 1, self.num_heads, self.scaling)
        (kh,) = ops.split_heads(k.view(bsz, src_len, embed_dim), 1, self.num_heads)
        (vh,) = ops.split_heads(v.view(bsz, src_len, embed_dim), 1, self.num_heads)
        o, _, _ = _materialised_attention(qh, kh, vh, key_padding_mask, bias4, self.dropout, self.training, False)
        return o

    def forward(
        self,
        query,
        key,
        value,
        key_padding_mask: Optional[Tensor] = None,
        attn_bias: Optional[Tensor] = None,
    ) -> Tensor:
        return self.out_proj(self.attend(query, key, value, key_padding_mask, attn_bias))
