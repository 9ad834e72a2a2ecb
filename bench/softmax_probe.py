#!/usr/bin/env python3
"""One forward + backward of the logits-mode softmax at Uni-Mol's shape (for ``ncu -k regex:softmax_dropout``)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore import ops  # noqa: E402

B, H, L = 32, 64, 248
torch.manual_seed(0)
x = torch.randn(B, H, L, L, device="cuda", dtype=torch.bfloat16)
bias = torch.randn(B, H, L, L, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
pad = torch.zeros(B, 1, 1, L, device="cuda", dtype=torch.bfloat16)
pad[:, :, :, -7:] = float("-inf")
for _ in range(2):
    out, z = ops.softmax_dropout_with_logits(x, 0.1, True, mask=pad, bias=bias)
    torch.autograd.backward([out, z], [torch.ones_like(out), torch.ones_like(z)])
torch.cuda.synchronize()
