"""A Uni-Mol style plug-in written ONLY against the public Uni-Core API (``unicore.modules``, ``unicore.models``,
``unicore.losses``, ``unicore.tasks``, ``unicore.data``) - it loads unchanged into the reference framework and into this
one ("plug-in written once, run under both frameworks").

Same architecture and loss as ``examples/unimol`` (15-layer pair-bias encoder, 512 dim, 64 heads, Gaussian distance
basis, masked atom / coordinate / distance heads) in the plain PyTorch formulation: the pair representation is threaded
through the layers as the attention bias and comes back as the attention logits (``return_attn=True``).
``examples/unimol`` is this framework's optimised build of the same model (fused Gaussian basis, logits-mode softmax,
register-tile layout kernels); this package is the common denominator used for cross-framework measurements.

    --user-dir examples/unimol_portable --task synthetic_unimol_portable --loss unimol_portable --arch unimol_portable_base
"""
from . import model, task  # noqa: F401
