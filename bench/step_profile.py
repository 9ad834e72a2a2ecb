#!/usr/bin/env python3
"""Per-kernel time breakdown of the flagship training step (torch.profiler / CUPTI, warm caches).
Prints the top kernels by total device time over N steps; used to decide what to fuse next."""
import os, sys, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B  # repo-root bench.py

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--impl", default="ours")
ap.add_argument("--top", type=int, default=45)
a, _ = ap.parse_known_args()
sys.argv = [sys.argv[0]]
args = B.parse()
args.impl = a.impl
B.setup_paths(a.impl)
torch.cuda.set_device(0)
targs, task, trainer = B.build_trainer(args, a.impl, 1, 0, 0)
d = task.dictionary
batches = B.make_batches(4, args.batch_size, args.seq_len, len(d), d.pad(), task.mask_idx,
                         special=[d.pad(), d.unk(), d.bos(), d.eos(), task.mask_idx], seed=1)
dev = [{"net_input": {"src_tokens": b["net_input"]["src_tokens"].cuda()}, "target": b["target"].cuda()} for b in batches]
for i in range(4):
    trainer.train_step([dev[i % 4]])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(a.steps):
        trainer.train_step([dev[i % 4]])
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0)
    if t > 0:
        rows.append((t / a.steps, e.count / a.steps, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total device time per step: %.2f ms over %d kernels/step" % (tot / 1e3, sum(r[1] for r in rows)))
for t, c, k in rows[: a.top]:
    print("%9.1f us %6.1f  %5.1f%%  %s" % (t, c, 100 * t / tot, k[:110]))
