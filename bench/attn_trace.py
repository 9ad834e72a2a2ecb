#!/usr/bin/env python3
"""Print the per-phase clock64() deltas of one CTA of the fused attention forward kernel
(UNICORE_FMHA_TRACE=1; profiling aid, see csrc/attn/fmha_fwd_sm100.cu UB_TRACE)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from unicore import ops  # noqa: E402

B, H, L = 32, 12, 512
torch.manual_seed(0)
qkv = (torch.randn(B, L, 3, H, 64, device="cuda") * 0.5).half()
bias = torch.randn(1, H, L, L, device="cuda").half()
kpm = torch.zeros(B, L, dtype=torch.bool, device="cuda")
for name, kw in (("plain", {}), ("bias_mask_dropout", dict(bias=bias, key_padding_mask=kpm, dropout_p=0.1))):
    for it in range(3):
        if it == 2:
            os.environ["UNICORE_FMHA_TRACE"] = "1"
            print("==", name, flush=True)
        q = qkv.detach().clone().requires_grad_(True)
        b = kw.get("bias")
        kw2 = dict(kw)
        if b is not None:
            kw2["bias"] = b.detach().clone().requires_grad_(True)
        out = ops.fused_attention_qkvpacked(q, training=True, **kw2)
        out.backward(torch.ones_like(out))
        torch.cuda.synchronize()
    os.environ.pop("UNICORE_FMHA_TRACE", None)
print("bwd phases: 1 issue prefetch(i+1) | 2 lse/delta/bits loads | 3 wait S,dP | 4 softmax-grad math + dBias red + P/dS stores |"
      " 5 wait prefetch | 6 fence+sync | 7 issue dV,dK,dQ MMAs (+S,dP of i+1) | 8 wait dQ | 9 dQ read-out + red")
print("phases: 1 wait PV(j-1) | 2 issue copies+sync (K landed) | 3 QK issue | 4 wait bias | 5 wait S | 6 sync |"
      " 7 logits+max | 8 xchg sync | 9 exp/dropout/P store | 10 O rescale | 11 wait V + sync")
