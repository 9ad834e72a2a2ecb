#!/usr/bin/env python3
"""Per-kernel time breakdown of the flagship training step (torch.profiler / CUPTI, warm caches).

Prints the top kernels by total device time over N steps (rank 0), the device-busy fraction of the
step and the largest idle gaps with the kernels around them; used to decide what to fuse next.
Single GPU: ``python bench/step_profile.py``; data parallel: launch under ``torch.distributed.run``.
Extra flags are handed to ``bench.py``'s parser (e.g. ``--ddp-backend b200``).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench as B  # noqa: E402  repo-root bench.py

ap = argparse.ArgumentParser()
ap.add_argument("--prof-steps", type=int, default=3)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--gaps", type=int, default=12)
ap.add_argument("--cprofile", type=int, default=0, help="also run this many steps under cProfile (host-side cost)")
ap.add_argument("--trace", default="", help="write a chrome trace (CPU + CUDA activities) of the profiled steps here")
a, rest = ap.parse_known_args()
sys.argv = [sys.argv[0]] + rest
args = B.parse()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
B.setup_paths(args.impl)
torch.cuda.set_device(local_rank)
if world > 1:
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="nccl", init_method="env://", world_size=world, rank=rank)
    dist.all_reduce(torch.zeros(1, device="cuda"))
targs, task, trainer = B.build_trainer(args, args.impl, world, rank, local_rank)
d = task.dictionary
batches = B.make_batches(4, args.batch_size, args.seq_len, len(d), d.pad(), task.mask_idx,
                         special=[d.pad(), d.unk(), d.bos(), d.eos(), task.mask_idx], seed=1 + rank)
dev = [{"net_input": {"src_tokens": b["net_input"]["src_tokens"].cuda()}, "target": b["target"].cuda()} for b in batches]
for i in range(4):
    trainer.train_step([dev[i % 4]])
torch.cuda.synchronize()
if a.cprofile > 0:
    import cProfile
    import pstats
    import time

    pr = cProfile.Profile()
    t0 = time.time()
    pr.enable()
    for i in range(a.cprofile):
        trainer.train_step([dev[i % 4]])
    torch.cuda.synchronize()
    pr.disable()
    if rank == 0:
        print("cProfile: %.2f ms/step wall over %d steps" % ((time.time() - t0) * 1e3 / a.cprofile, a.cprofile))
        st = pstats.Stats(pr)
        st.sort_stats("tottime").print_stats(35)
        st.sort_stats("cumulative").print_stats(45)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

acts = [ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if a.trace else [])
with profile(activities=acts, with_stack=bool(a.trace)) as prof:
    for i in range(a.prof_steps):
        trainer.train_step([dev[i % 4]])
    torch.cuda.synchronize()
if rank == 0 and a.trace:
    prof.export_chrome_trace(a.trace)
if rank == 0:
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = getattr(e, "cuda_time_total", 0)
        if t > 0:
            rows.append((t / a.prof_steps, e.count / a.prof_steps, e.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print("total device time per step: %.2f ms over %d kernels/step" % (tot / 1e3, sum(r[1] for r in rows)))
    for t, c, k in rows[: a.top]:
        print("%9.1f us %6.1f  %5.1f%%  %s" % (t, c, 100 * t / tot, k[:110]))
    # timeline: busy fraction + the largest idle gaps (all streams merged)
    evs = []
    for e in prof.events():
        tr = getattr(e, "time_range", None)
        if tr is None or getattr(e, "device_type", None) is None:
            continue
        if "cuda" not in str(e.device_type).lower():
            continue
        evs.append((tr.start, tr.end, e.name))
    evs.sort()
    if evs:
        span = evs[-1][1] - evs[0][0]
        busy, cur_end, gaps = 0.0, evs[0][0], []
        prev_name = ""
        for s, en, name in evs:
            if s > cur_end:
                gaps.append((s - cur_end, prev_name, name))
                busy += en - s
                cur_end = en
                prev_name = name
            elif en > cur_end:
                busy += en - cur_end
                cur_end = en
                prev_name = name
        print("timeline: span %.2f ms/step, busy %.2f ms/step (%.1f%%), %d gaps/step" % (
            span / 1e3 / a.prof_steps, busy / 1e3 / a.prof_steps, 100 * busy / span, len(gaps) / a.prof_steps))
        gaps.sort(reverse=True)
        small = sum(g[0] for g in gaps if g[0] < 20)
        print("sum of gaps < 20 us: %.2f ms/step" % (small / 1e3 / a.prof_steps))
        for g, before, after in gaps[: a.gaps]:
            print("  gap %8.1f us  after %-60s before %s" % (g, before[:60], after[:60]))
