#!/usr/bin/env python3
"""Replicated vs sharded optimizer step under ``--ddp-backend b200`` (run with torchrun on >= 2 GPUs).

Trains the same tiny BERT twice from the same seed - once with the stock fused Adam on every rank, once with
``UNICORE_B200_SHARD_OPTIMIZER=1|2`` (Adam on a 1/N shard + parameter all-gather in one kernel; mode 2 also stops
the gradient buckets after their reduce-scatter half) - and compares the
16-bit parameters, the fp32 master weights and the Adam moments after ``consolidate_state``.  The per-element
arithmetic is the same formula in two kernels (FMA contraction may differ), so fp32 state must agree to ~1e-6 and
the 16-bit parameters to one rounding step; a shard that was not updated or not gathered shows up as ~lr * steps.
Prints one JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "examples"))


def run(mode, steps, precision, rank, local_rank, world):
    import torch

    from unicore import options, tasks, utils
    from unicore.trainer import Trainer

    os.environ["UNICORE_B200_SHARD_OPTIMIZER"] = str(int(mode))
    flags = [
        "--task", "synthetic_mlm", "--loss", "masked_lm", "--arch", "bert_base", "--encoder-layers", "2",
        "--encoder-embed-dim", "128", "--encoder-ffn-embed-dim", "256", "--encoder-attention-heads", "2",
        "--synthetic-vocab-size", "509", "--synthetic-seq-len", "64", "--max-seq-len", "64", "--optimizer", "adam",
        "--lr", "1e-3", "--lr-scheduler", "fixed", "--weight-decay", "0.01", "--clip-norm", "1.0", "--max-update", "100",
        "--batch-size", "8", "--seed", "11", "--no-save", "--disable-validation", "--log-format", "none",
        "--distributed-world-size", str(world), "--distributed-rank", str(rank), "--device-id", str(local_rank),
        "--ddp-backend", "b200", "--" + precision,
    ]
    if precision == "fp16":
        flags += ["--fp16-init-scale", "4"]
    args = options.parse_args_and_arch(options.get_training_parser(), input_args=flags)
    args.distributed_rank, args.device_id = rank, local_rank
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    task = tasks.setup_task(args)
    model = task.build_model(args)
    trainer = Trainer(args, task, model, task.build_loss(args))
    trainer._total_train_steps = args.max_update
    task.load_dataset("train")
    ds = task.dataset("train")
    for i in range(steps):
        rows = [ds[(i * world + rank) * 8 + k] for k in range(8)]
        trainer.train_step([utils.move_to_cuda(ds.collater(rows))])
    if hasattr(trainer.optimizer, "resolve_pending_overflow"):
        trainer.optimizer.resolve_pending_overflow()
    trainer.consolidate_optimizer_state()
    torch.cuda.synchronize()
    sharded = getattr(trainer.optimizer, "_sharded", None) is not None
    params = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
    opt = trainer.optimizer.state_dict()
    state = {k: {n: t.detach().float().cpu() for n, t in v.items() if torch.is_tensor(t)} for k, v in opt["state"].items()}
    masters = [g["params"][0].detach().float().cpu() for g in trainer.optimizer.fp32_params]
    return sharded, params, state, masters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--precision", default="bf16", choices=["fp16", "bf16"])
    ap.add_argument("--mode", type=int, default=1, choices=[1, 2],
                    help="1: contiguous shard after the full all-reduce; 2: reduce-scatter-only buckets, slice-wise shard")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="nccl", init_method="env://", world_size=world, rank=rank)
    importlib.import_module("bert")
    sharded0, p0, s0, m0 = run(0, a.steps, a.precision, rank, local_rank, world)
    sharded1, p1, s1, m1 = run(a.mode, a.steps, a.precision, rank, local_rank, world)
    diffs = {"params": float((p0 - p1).abs().max())}
    diffs["master"] = max(float((x - y).abs().max()) for x, y in zip(m0, m1))
    for k in s0:
        for n in s0[k]:
            diffs["state_{}_{}".format(k, n)] = float((s0[k][n] - s1[k][n]).abs().max())
    fp32_worst = max(v for k, v in diffs.items() if k != "params")
    worst = torch.tensor([diffs["params"], fp32_worst], device="cuda")
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"summary": "sharded_optimizer_check", "world": world, "mode": a.mode, "replicated_was_sharded": sharded0,
                          "sharded_active": sharded1, "max_param_diff": float(worst[0].item()),
                          "max_fp32_state_diff": float(worst[1].item()), "diffs": diffs}))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
