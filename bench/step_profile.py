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
ap.add_argument("--host", action="store_true",
                help="also record host activities and list the CUDA runtime calls the host spent time in (blocking "
                     "synchronisations show up here) and the launch lead of the host over the device per step")
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

acts = [ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if (a.trace or a.host) else [])
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

if rank == 0 and a.host:
    # host side: which runtime calls block, and how far ahead of the device the launches are
    runtime = {}
    launches = []   # (host time of the launch call)
    for e in prof.events():
        tr = getattr(e, "time_range", None)
        if tr is None or "cuda" in str(getattr(e, "device_type", "")).lower():
            continue
        name = e.name
        if name.startswith("cuda") or name.startswith("cu"):
            tot, cnt, mx = runtime.get(name, (0.0, 0, 0.0))
            dur = tr.end - tr.start
            runtime[name] = (tot + dur, cnt + 1, max(mx, dur))
            if "Launch" in name or name in ("cudaMemcpyAsync", "cudaMemsetAsync"):
                launches.append(tr.start)
    print("host runtime calls (per step): total us, calls, longest us")
    for name, (tot, cnt, mx) in sorted(runtime.items(), key=lambda kv: -kv[1][0])[:12]:
        print("  %-34s %9.1f %7.1f %9.1f" % (name[:34], tot / a.prof_steps, cnt / a.prof_steps, mx))
    kern = sorted((s, en) for s, en, _n in evs)
    launches.sort()
    if launches and kern:
        # lead = device start of the k-th kernel minus host time of the k-th launch call (same order on one stream)
        n = min(len(launches), len(kern))
        lead = [kern[k][0] - launches[k] for k in range(n)]
        per = max(1, n // (a.prof_steps * 10))
        print("host lead over the device (us) at every %d-th launch of the profiled steps:" % per)
        print("  " + " ".join("%.0f" % lead[k] for k in range(0, n, per)))
