"""Fused element-wise / reduction ops around the GEMMs of a transformer layer.

These fusions do not exist in the reference (it runs GELU, three dropouts, residual adds,
LayerNorm and the fp32 log-softmax as separate ATen kernels,
``unicore/modules/transformer_encoder_layer.py:79-94``, ``losses/masked_lm.py:31-36``); on B200
every one of them is HBM-bound, so they are merged into the minimum number of passes:

* ``bias_gelu(x, bias)``                      - GEMM-output bias add + exact GELU (erf), fwd/bwd;
* ``bias_dropout_add_layer_norm(...)``        - ``LN(residual + dropout(x + bias))`` (post-LN) in
  one pass, also returning the pre-LN sum when requested (pre-LN);
* ``softmax_cross_entropy(logits, target)``   - summed NLL of an fp32 log-softmax without
  materialising the ``[N, V]`` fp32 log-probabilities; backward writes the gradient in the
  logits' dtype in place of the logits buffer.
Kernels: ``csrc/fused/*.cu``.  PyTorch fallbacks are bit-for-bit the reference formulation.
"""
import math
from typing import Optional

import torch
import torch.nn.functional as F

from ._native import aligned_param, native, use_native


# ------------------------------------------------------------------------------------------------
# bias + GELU
# ------------------------------------------------------------------------------------------------
class _BiasGeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias):
        ctx.save_for_backward(x, bias)
        return native().bias_gelu_fwd(x, bias)

    @staticmethod
    def backward(ctx, dy):
        x, bias = ctx.saved_tensors
        dx = native().bias_gelu_bwd(dy.contiguous(), x, bias)
        dbias = dx.view(-1, dx.shape[-1]).sum(dim=0) if bias is not None else None
        return dx, dbias


def bias_gelu(x: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    if use_native(x, bias) and x.dtype in (torch.float16, torch.bfloat16) and x.is_contiguous() and x.shape[-1] % 8 == 0:
        return _BiasGeluFn.apply(x, aligned_param(bias, x.dtype))
    return F.gelu(x + bias if bias is not None else x)


# ------------------------------------------------------------------------------------------------
# bias + dropout + residual add (+ LayerNorm)
# ------------------------------------------------------------------------------------------------
class _BiasDropoutAddLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, residual, ln_weight, ln_bias, p, eps, training):
        p = float(p) if training else 0.0
        shape = x.shape
        x2 = x.contiguous().view(-1, shape[-1])
        r2 = residual.contiguous().view(-1, shape[-1])
        y, mean, rstd, summed, seed, offset = native().bias_dropout_add_ln_fwd(x2, bias, r2, ln_weight, ln_bias, p, eps)
        ctx.save_for_backward(summed, ln_weight, mean, rstd)
        ctx.p = p
        ctx.rng = (seed, offset)
        ctx.has_bias = bias is not None
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        summed, ln_weight, mean, rstd = ctx.saved_tensors
        dy2 = dy.contiguous().view(summed.shape)
        # dsum: gradient w.r.t. (residual + dropout(x+bias)); dx = dropout-masked dsum
        dsum, dx, dgamma, dbeta = native().bias_dropout_add_ln_bwd(
            dy2, summed, mean, rstd, ln_weight, ctx.p, ctx.rng[0], ctx.rng[1]
        )
        dbias = dx.sum(dim=0) if ctx.has_bias else None
        return dx.view(dy.shape), dbias, dsum.view(dy.shape), dgamma, dbeta, None, None, None


def bias_dropout_add_layer_norm(x, bias, residual, ln_weight, ln_bias, p, eps, training):
    """``LayerNorm(residual + dropout(x + bias))`` - the post-LN block epilogue in one pass."""
    if (
        use_native(x, bias, residual, ln_weight, ln_bias)
        and x.dtype in (torch.float16, torch.bfloat16)
        and x.shape[-1] % 8 == 0
        and x.shape[-1] <= 8192
        and ln_weight is not None and ln_bias is not None
        and hasattr(native(), "bias_dropout_add_ln_fwd")
    ):
        w = aligned_param(ln_weight, x.dtype)
        b = aligned_param(ln_bias, x.dtype)
        bb = aligned_param(bias, x.dtype)
        return _BiasDropoutAddLNFn.apply(x, bb, residual, w, b, p, eps, training)
    from .norm_ops import layer_norm

    h = x + bias if bias is not None else x
    h = residual + F.dropout(h, p=p, training=training)
    return layer_norm(h, (h.shape[-1],), ln_weight, ln_bias, eps)


# ------------------------------------------------------------------------------------------------
# softmax cross entropy (sum reduction, fp32 math)
# ------------------------------------------------------------------------------------------------
class _SoftmaxXentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        loss_rows, lse = native().softmax_xent_fwd(logits, target, int(ignore_index))
        ctx.save_for_backward(logits, target, lse)
        ctx.ignore_index = int(ignore_index)
        return loss_rows.sum()

    @staticmethod
    def backward(ctx, dloss):
        logits, target, lse = ctx.saved_tensors
        dlogits = native().softmax_xent_bwd(logits, target, lse, dloss.float().reshape(1), ctx.ignore_index)
        return dlogits, None, None


def softmax_cross_entropy(logits: torch.Tensor, target: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    """``nll_loss(log_softmax(logits, fp32), target, reduction='sum', ignore_index=...)``."""
    if (
        use_native(logits, target)
        and logits.dim() == 2
        and logits.dtype in (torch.float16, torch.bfloat16, torch.float32)
        and logits.is_contiguous()
        and logits.numel() > 0
        and hasattr(native(), "softmax_xent_fwd")
    ):
        return _SoftmaxXentFn.apply(logits, target.contiguous(), ignore_index)
    return F.nll_loss(
        F.log_softmax(logits, dim=-1, dtype=torch.float32), target, ignore_index=ignore_index, reduction="sum"
    )
