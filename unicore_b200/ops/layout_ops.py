"""Token-major <-> head-major relayout (``csrc/fused/head_permute.cu``).

``split_heads`` turns a packed projection ``[B, L, T*H*D]`` (``T`` = 3 for q|k|v) into ``T`` contiguous
``[B, H, L, D]`` tensors with ONE kernel (the query scale folded in), ``merge_heads`` is the inverse; each is
the other's backward, so the packed gradient of the in-projection is gathered by one kernel from the three
gradients autograd hands back (reference formulation: ``unicore/modules/multihead_attention.py:62-76,105-110``
- view / transpose / contiguous chains whose backward zero-fills and adds three full-size tensors).
"""
from typing import Optional, Sequence, Tuple

import torch

from ._native import native, use_native


def _eligible(x: torch.Tensor, head_dim: int) -> bool:
    return (
        use_native(x)
        and x.dtype in (torch.float16, torch.bfloat16, torch.float32)
        and (head_dim * x.element_size()) % 16 == 0
        and x.numel() > 0
    )


def _split_reference(x, num_slices, num_heads, scale0):
    bsz, seq_len, width = x.shape
    parts = x.view(bsz, seq_len, num_slices, num_heads, width // (num_slices * num_heads)).permute(2, 0, 3, 1, 4)
    out = [parts[t].contiguous() for t in range(num_slices)]
    if scale0 != 1.0:
        out[0] = out[0] * scale0
    return tuple(out)


def _merge_reference(heads: Sequence[Optional[torch.Tensor]], scale0):
    ref = next(h for h in heads if h is not None)
    full = [torch.zeros_like(ref) if h is None else h for h in heads]
    if scale0 != 1.0:
        full[0] = full[0] * scale0
    bsz, num_heads, seq_len, dim = ref.shape
    return torch.stack(full, dim=0).permute(1, 3, 0, 2, 4).reshape(bsz, seq_len, len(full) * num_heads * dim)


class _SplitHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, num_slices, num_heads, scale0):
        ctx.scale0 = scale0
        return tuple(native().split_heads(x, num_slices, num_heads, scale0))

    @staticmethod
    def backward(ctx, *grads):
        grads = [None if g is None else g.contiguous() for g in grads]
        return native().merge_heads(grads, ctx.scale0), None, None, None


class _MergeHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scale0, *heads):
        ctx.scale0 = scale0
        ctx.count = len(heads)
        ctx.num_heads = heads[0].shape[1]
        return native().merge_heads(list(heads), scale0)

    @staticmethod
    def backward(ctx, grad):
        parts = native().split_heads(grad.contiguous(), ctx.count, ctx.num_heads, ctx.scale0)
        return (None,) + tuple(parts)


def split_heads(x: torch.Tensor, num_slices: int, num_heads: int, scale0: float = 1.0) -> Tuple[torch.Tensor, ...]:
    """``[B, L, T*H*D] -> T x [B, H, L, D]`` (contiguous); slice 0 is multiplied by ``scale0``."""
    if x.dim() != 3 or x.shape[-1] % (num_slices * num_heads) != 0:
        raise ValueError("expected [B, L, T*H*D], got {}".format(tuple(x.shape)))
    head_dim = x.shape[-1] // (num_slices * num_heads)
    if 1 <= num_slices <= 4 and _eligible(x, head_dim):
        return _SplitHeadsFn.apply(x.contiguous(), num_slices, num_heads, float(scale0))
    return _split_reference(x.contiguous(), num_slices, num_heads, float(scale0))


def merge_heads(*heads: torch.Tensor, scale0: float = 1.0) -> torch.Tensor:
    """``T x [B, H, L, D] -> [B, L, T*H*D]``; slice 0 is multiplied by ``scale0``."""
    ref = heads[0]
    if ref.dim() != 4:
        raise ValueError("expected [B, H, L, D], got {}".format(tuple(ref.shape)))
    if 1 <= len(heads) <= 4 and _eligible(ref, ref.shape[-1]):
        return _MergeHeadsFn.apply(float(scale0), *[h.contiguous() for h in heads])
    return _merge_reference(list(heads), float(scale0))
