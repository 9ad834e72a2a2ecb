#!/usr/bin/env python3
"""Embedding-gradient paths, ours vs ATen (the reference uses ``nn.Embedding`` / ``F.embedding`` for both):

* token table  [30522, 768] fp16, 32 x 512 indices: ``ops.embedding`` backward (fp32 red scatter + finalize,
  ``csrc/fused/embedding.cu``) vs ``F.embedding`` backward (radix sort + segmented reduction);
* relative-position bias [32 buckets, 12 heads] over a 512 x 512 bucket table: the cached one-hot GEMM of
  ``unicore/modules/transformer.py`` vs ``F.embedding`` + permute.

Forward + backward device time per call (CUDA events, L2 flushed between iterations), one JSON line each with clocks.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from op_compare import Clocks  # noqa: E402  (bench/op_compare.py)
from unicore import ops  # noqa: E402
from unicore.modules import TransformerEncoder  # noqa: E402


def timed(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        total += a.elapsed_time(b)
    return total / iters * 1e3   # us


def main():
    clocks = Clocks()
    torch.manual_seed(0)
    V, D, B, L = 30522, 768, 32, 512
    w = (torch.randn(V, D, device="cuda") * 0.02).half().requires_grad_(True)
    tok = torch.randint(5, V, (B, L), device="cuda")
    dy = torch.randn(B, L, D, device="cuda").half()

    def run(lookup):
        w.grad = None
        lookup(tok, w, 0).backward(dy)

    for name, fn in (("ours", ops.embedding), ("aten", F.embedding)):
        mark = clocks.mark()
        us = timed(lambda: run(fn))
        print(json.dumps({"op": "embedding fwd+bwd [32x512] of [30522,768] fp16", "impl": name, "us": round(us, 1),
                          "clocks": clocks.since(mark)}), flush=True)

    enc = TransformerEncoder(encoder_layers=1, embed_dim=768, ffn_embed_dim=768, attention_heads=12, max_seq_len=512,
                             rel_pos=True).cuda().half()
    x = torch.zeros(1, L, 768, device="cuda", dtype=torch.half)
    g = torch.randn(12, L, L, device="cuda").half()
    table = enc.relative_attention_bias.weight

    def gemm_path():
        table.grad = None
        enc.get_rel_pos_bias(x).backward(g)

    def aten_path():
        table.grad = None
        F.embedding(enc.rp_bucket[:L, :L], table).permute(2, 0, 1).contiguous().backward(g)

    for name, fn in (("ours (one-hot GEMM)", gemm_path), ("aten", aten_path)):
        mark = clocks.mark()
        us = timed(fn)
        print(json.dumps({"op": "relative-position bias fwd+bwd [12,512,512] from [32,12] fp16", "impl": name,
                          "us": round(us, 1), "clocks": clocks.since(mark)}), flush=True)
    clocks.stop()


if __name__ == "__main__":
    main()
