"""Fused multi-head attention: ``softmax(scale * q k^T + bias + key_padding) -> dropout -> @ v``.

The reference materialises the ``[B*H, L, L]`` score tensor three times per layer
(``unicore/modules/multihead_attention.py:47-114``: bmm -> softmax_dropout -> bmm, plus four
transpose+contiguous copies).  The sm_100a kernel (``csrc/attn/fmha_sm100.cu``) keeps scores in
tensor memory: ``tcgen05.mma`` computes S = QK^T into TMEM, softmax warps read it with
``tcgen05.ld``, add the bias tile / padding mask, apply Philox dropout, write P (16-bit) to shared
memory and a second ``tcgen05.mma`` accumulates O += PV in TMEM.  q/k/v are consumed directly from
the packed ``in_proj`` output through strides (no transposes) and O is written as ``[B, L, H*D]``.
Backward recomputes P from the saved log-sum-exp (flash-attention style) and regenerates the
dropout mask from the same Philox counters.
"""
import math

import torch
import torch.nn.functional as F

from ._native import native, use_native


def attention_reference(q, k, v, bias=None, key_padding_mask=None, dropout_p=0.0, training=True, scale=None):
    """Plain PyTorch implementation with the same signature (fallback + test oracle).

    q: [B, Lq, H, D], k/v: [B, Lk, H, D]; bias broadcastable to [B, H, Lq, Lk];
    key_padding_mask: [B, Lk] bool (True = masked). Returns [B, Lq, H, D].
    """
    B, Lq, H, D = q.shape
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    scores = torch.einsum("bqhd,bkhd->bhqk", q, k) * scale
    if bias is not None:
        scores = scores + bias.to(scores.dtype)
    if key_padding_mask is not None:
        scores = scores.masked_fill(key_padding_mask[:, None, None, :].to(torch.bool), float("-inf"))
    probs = F.softmax(scores.float(), dim=-1).to(q.dtype)
    probs = F.dropout(probs, p=dropout_p, training=training)
    return torch.einsum("bhqk,bkhd->bqhd", probs, v)


class _FusedAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, bias, key_padding_mask, dropout_p, training, scale):
        p = float(dropout_p) if training else 0.0
        out, lse, bits = native().fmha_fwd(q, k, v, bias, key_padding_mask, p, float(scale))
        ctx.save_for_backward(q, k, v, out, lse, bias, key_padding_mask, bits)
        ctx.p = p
        ctx.scale = float(scale)
        ctx.need_dbias = bias is not None and bias.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, bias, kpm, bits = ctx.saved_tensors
        dq, dk, dv, dbias = native().fmha_bwd(
            dout.contiguous(), q, k, v, out, lse, bias, kpm, ctx.p, ctx.scale, bits, ctx.need_dbias, False
        )
        if dbias is not None and dbias.dtype != bias.dtype:
            dbias = dbias.to(bias.dtype)
        return dq, dk, dv, dbias, None, None, None, None


class _FusedAttentionPackedFn(torch.autograd.Function):
    """Self-attention on a packed ``[B, L, 3, H, 64]`` projection: q/k/v are strided slices on the way
    in, and the backward kernels write dq/dk/dv directly into ONE packed gradient tensor."""

    @staticmethod
    def forward(ctx, qkv, bias, key_padding_mask, dropout_p, training, scale):
        p = float(dropout_p) if training else 0.0
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        out, lse, bits = native().fmha_fwd(q, k, v, bias, key_padding_mask, p, float(scale))
        ctx.save_for_backward(qkv, out, lse, bias, key_padding_mask, bits)
        ctx.p = p
        ctx.scale = float(scale)
        ctx.need_dbias = bias is not None and bias.requires_grad
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, bias, kpm, bits = ctx.saved_tensors
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        dq, _dk, _dv, dbias = native().fmha_bwd(
            dout.contiguous(), q, k, v, out, lse, bias, kpm, ctx.p, ctx.scale, bits, ctx.need_dbias, True
        )
        dqkv = dq._base  # the packed [B, L, 3, H, 64] tensor the three gradients are views of
        if dbias is not None and dbias.dtype != bias.dtype:
            dbias = dbias.to(bias.dtype)
        return dqkv, dbias, None, None, None, None


def fused_attention_supported(q, k, v, bias=None, key_padding_mask=None) -> bool:
    if not use_native(q, k, v, bias, key_padding_mask) or not hasattr(native(), "fmha_fwd"):
        return False
    if key_padding_mask is not None and key_padding_mask.dtype != torch.bool:
        return False
    for t in (q, k, v):
        if t.data_ptr() % 16 != 0 or any(s % 8 != 0 for s in t.stride()[:3]) or t.stride(2) != 64:
            return False
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        return False
    if q.dim() != 4 or q.shape[-1] != 64 or k.shape[-1] != 64:
        return False
    if q.stride(-1) != 1 or k.stride(-1) != 1 or v.stride(-1) != 1:
        return False
    if q.shape[1] % 8 != 0 or k.shape[1] % 8 != 0:
        return False
    if bias is not None:
        if bias.dim() != 4 or bias.shape[0] not in (1, q.shape[0]) or bias.shape[1] != q.shape[2]:
            return False
        if bias.shape[2] != q.shape[1] or bias.shape[3] != k.shape[1]:
            return False
        if bias.dtype not in (q.dtype, torch.float32):
            return False
        if bias.dtype == q.dtype and bias.is_contiguous() and bias.data_ptr() % 16 != 0:
            return False
    return True


def fused_attention(q, k, v, bias=None, key_padding_mask=None, dropout_p=0.0, training=True, scale=None):
    """Dispatch to the tcgen05 kernel when shapes/dtypes allow, else to the PyTorch reference."""
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    if fused_attention_supported(q, k, v, bias, key_padding_mask):
        if bias is not None:
            # the kernel stages the bias tile next to P in shared memory in the compute dtype
            bias = bias.to(q.dtype).contiguous()
        if key_padding_mask is not None:
            key_padding_mask = key_padding_mask.to(torch.bool).contiguous()
        return _FusedAttentionFn.apply(q, k, v, bias, key_padding_mask, dropout_p, training, scale)
    return attention_reference(q, k, v, bias, key_padding_mask, dropout_p, training, scale)


def fused_attention_qkvpacked(qkv, bias=None, key_padding_mask=None, dropout_p=0.0, training=True, scale=None):
    """Self-attention on the packed in_proj output ``qkv [B, L, 3, H, D]`` (returns ``[B, L, H, D]``)."""
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    if scale is None:
        scale = 1.0 / math.sqrt(q.shape[-1])
    if qkv.is_contiguous() and fused_attention_supported(q, k, v, bias, key_padding_mask):
        if bias is not None:
            bias = bias.to(q.dtype).contiguous()
        if key_padding_mask is not None:
            key_padding_mask = key_padding_mask.to(torch.bool).contiguous()
        return _FusedAttentionPackedFn.apply(qkv, bias, key_padding_mask, dropout_p, training, scale)
    return fused_attention(q, k, v, bias, key_padding_mask, dropout_p, training, scale)
