"""Python surface of the hand-written sm_100a kernels (+ PyTorch fallbacks for CPU runs)."""
from ._native import HAS_CUDA_EXT, USE_NATIVE, native, use_native  # noqa: F401
from .attention_ops import (  # noqa: F401
    attention_reference,
    fused_attention,
    fused_attention_qkvpacked,
    fused_attention_supported,
)
from .fused_ops import (  # noqa: F401
    bias_dropout_add_layer_norm,
    bias_gelu,
    embedding,
    gaussian_basis,
    gaussian_basis_reference,
    linear,
    softmax_cross_entropy,
    vocab_projection,
)
from .norm_ops import layer_norm, rms_norm  # noqa: F401
from .optim_ops import (  # noqa: F401
    ema_update_,
    fp32_to_bf16_sr,
    fused_adam,
    multi_tensor_l2norm,
    multi_tensor_scale_,
)
from .layout_ops import heads_to_pair, merge_heads, pair_tail, pair_to_heads, split_heads  # noqa: F401
from .softmax_ops import softmax_dropout, softmax_dropout_with_logits  # noqa: F401
