#!/usr/bin/env python3
"""Times the raw fmha_bwd binding (no autograd glue) under the profiling switches of
UNICORE_FMHA_DEBUG (1: no dBias reds, 2: no dQ reds) to attribute backward time."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unicore_b200.ops._native import native

B, H, L, D = 32, 12, 512, 64
dt = torch.float16
torch.manual_seed(0)
qkv = (torch.randn(B, L, 3, H, D, device="cuda") * 0.5).to(dt)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
bias = torch.randn(1, H, L, L, device="cuda").to(dt)
kpm = torch.zeros(B, L, dtype=torch.bool, device="cuda")
dout = torch.randn(B, L, H, D, device="cuda").to(dt)
n = native()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for use_bias, p in ((True, 0.1), (False, 0.0)):
    out, lse, bits = n.fmha_fwd(q, k, v, bias if use_bias else None, kpm, p, 0.125)
    for flags in (0, 1, 2, 3):
        os.environ["UNICORE_FMHA_DEBUG"] = str(flags)
        ts = []
        for it in range(6):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            n.fmha_bwd(dout, q, k, v, out, lse, bias if use_bias else None, kpm, p, 0.125, bits, use_bias)
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        print(json.dumps({"bias": use_bias, "p": p, "flags": flags, "bwd_total_ms": round(min(ts[1:]), 4)}), flush=True)
os.environ["UNICORE_FMHA_DEBUG"] = "0"
