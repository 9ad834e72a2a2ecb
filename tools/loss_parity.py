#!/usr/bin/env python3
"""Loss trajectory of N optimizer updates of a tiny BERT, fp32 on the CPU, for either implementation.

    python tools/loss_parity.py --impl ours      --init /tmp/init.pt --steps 6
    python tools/loss_parity.py --impl reference --init /tmp/init.pt --steps 6

The first arm that runs creates ``--init`` (the model's initial ``state_dict``); the other one loads it, so both
start from identical weights (the parameter names are the same by design).  Dropout is off, the batches are the
deterministic synthetic ones of ``bench.py``, the optimizer is Adam with clipping and a polynomial schedule.  One
process per arm: the reference lives in ``baseline/_ref`` under the same package name and cannot share an
interpreter with this repo.  Prints ``{"impl": ..., "losses": [...]}`` (the logged loss of every update).
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (path / task helpers shared with the benchmark)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--init", required=True)
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    why = bench.setup_paths(a.impl)
    if why is not None:
        print(json.dumps({"impl": a.impl, "unavailable": why}))
        return 0
    import torch

    if a.impl == "reference":
        import bert  # noqa: F401  reference examples/bert
    else:
        import importlib

        sys.path.insert(0, os.path.join(REPO, "examples"))
        importlib.import_module("bert")
    from unicore import options, tasks
    from unicore.trainer import Trainer

    bench.register_bench_task(a.impl)
    flags = [
        "--task", "bench_mlm", "--bench-vocab", "512", "--loss", "masked_lm", "--arch", "bert_base",
        "--encoder-layers", "2", "--encoder-embed-dim", "64", "--encoder-ffn-embed-dim", "128",
        "--encoder-attention-heads", "4", "--max-seq-len", "32", "--dropout", "0.0", "--emb-dropout", "0.0",
        "--attention-dropout", "0.0", "--activation-dropout", "0.0", "--pooler-dropout", "0.0",
        "--optimizer", "adam", "--adam-betas", "(0.9, 0.98)", "--adam-eps", "1e-6", "--weight-decay", "0.01",
        "--clip-norm", "1.0", "--lr-scheduler", "polynomial_decay", "--lr", "1e-3", "--warmup-updates", "2",
        "--total-num-update", "20", "--max-update", "20", "--batch-size", "4", "--seed", "1", "--num-workers", "0",
        "--log-format", "none", "--disable-validation", "--no-save", "--distributed-world-size", "1", "--cpu",
    ]
    args = options.parse_args_and_arch(options.get_training_parser(), input_args=flags)
    torch.manual_seed(args.seed)
    task = tasks.setup_task(args)
    model = task.build_model(args)
    if os.path.exists(a.init):
        model.load_state_dict(torch.load(a.init, map_location="cpu"))
    else:
        torch.save(model.state_dict(), a.init)
    trainer = Trainer(args, task, model, task.build_loss(args))
    trainer._total_train_steps = args.max_update
    d = task.dictionary
    batches = bench.make_batches(4, 4, 24, len(d), d.pad(), task.mask_idx,
                                 special=[d.pad(), d.unk(), d.bos(), d.eos(), task.mask_idx], seed=99)
    losses = []
    for i in range(a.steps):
        out = trainer.train_step([batches[i % len(batches)]])
        losses.append(float(out["loss"]))  # what the trainer logs: bits per masked token, 3 decimals
    print(json.dumps({"impl": a.impl, "losses": losses}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
