#!/usr/bin/env python3
"""Host (cProfile) and device (CUPTI kernel table) breakdown of one Uni-Mol training step on one GPU.

    python bench/unimol_profile.py            # prints the 22 hottest host functions and the 28 hottest kernels
"""
import cProfile
import importlib
import os
import pstats
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "examples"))


def main():
    importlib.import_module("unimol")
    from torch.profiler import ProfilerActivity, profile

    from unicore import options, tasks, utils
    from unicore.trainer import Trainer

    flags = [
        "--task", "synthetic_unimol", "--loss", "unimol", "--arch", "unimol_base", "--synthetic-num-samples", "256",
        "--optimizer", "adam", "--adam-betas", "(0.9, 0.99)", "--clip-norm", "1.0", "--lr", "1e-4",
        "--lr-scheduler", "polynomial_decay", "--warmup-updates", "100", "--total-num-update", "100000",
        "--max-update", "100000", "--batch-size", "32", "--seed", "1", "--no-save", "--disable-validation",
        "--log-format", "none", "--distributed-world-size", "1", "--bf16",
    ]
    args = options.parse_args_and_arch(options.get_training_parser(), input_args=flags)
    task = tasks.setup_task(args)
    trainer = Trainer(args, task, task.build_model(args), task.build_loss(args))
    trainer._total_train_steps = args.max_update
    task.load_dataset("train")
    ds = task.dataset("train")
    batches = [utils.move_to_cuda(ds.collater([ds[k * 32 + i] for i in range(32)])) for k in range(2)]
    for i in range(3):
        trainer.train_step([batches[i % 2]])
    torch.cuda.synchronize()

    prof = cProfile.Profile()
    t0 = time.time()
    prof.enable()
    for i in range(4):
        trainer.train_step([batches[i % 2]])
    torch.cuda.synchronize()
    prof.disable()
    print("cProfile: %.1f ms/step wall" % ((time.time() - t0) * 250))
    pstats.Stats(prof).sort_stats("tottime").print_stats(22)

    with profile(activities=[ProfilerActivity.CUDA]) as kprof:
        for i in range(2):
            trainer.train_step([batches[i % 2]])
        torch.cuda.synchronize()
    rows = []
    for e in kprof.key_averages():
        t = getattr(e, "device_time_total", 0)
        if t > 0:
            rows.append((t / 2, e.count / 2, e.key))
    rows.sort(reverse=True)
    total = sum(r[0] for r in rows)
    print("total device %.1f ms over %d kernels" % (total / 1e3, sum(r[1] for r in rows)))
    for t, c, k in rows[:28]:
        print("%9.1f us %6.1f %5.1f%% %s" % (t, c, 100 * t / total, k[:120]))


if __name__ == "__main__":
    main()
