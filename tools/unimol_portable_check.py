#!/usr/bin/env python3
"""Run the portable Uni-Mol plug-in (examples/unimol_portable) for a few CPU updates under either framework.

    python tools/unimol_portable_check.py --impl ours      --init /tmp/u.pt --steps 3
    python tools/unimol_portable_check.py --impl reference --init /tmp/u.pt --steps 3

Same initial weights (first arm writes ``--init``), same synthetic molecules, dropout off: the logged losses of the two
frameworks must agree - the plug-in is one piece of code, the framework underneath is what differs.
"""
import argparse
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["ours", "reference"], required=True)
    ap.add_argument("--init", required=True)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    why = bench.setup_paths(a.impl)
    if why is not None:
        print(json.dumps({"impl": a.impl, "unavailable": why}))
        return 0
    import torch

    sys.path.insert(0, os.path.join(REPO, "examples"))
    importlib.import_module("unimol_portable")
    from unicore import options, tasks
    from unicore.trainer import Trainer

    flags = [
        "--task", "synthetic_unimol_portable", "--loss", "unimol_portable", "--arch", "unimol_portable_base",
        "--encoder-layers", "2", "--encoder-embed-dim", "32", "--encoder-ffn-embed-dim", "64",
        "--encoder-attention-heads", "4", "--gaussian-kernels", "16", "--synthetic-num-samples", "16",
        "--synthetic-min-atoms", "5", "--synthetic-max-atoms", "14", "--synthetic-atom-types", "9",
        "--dropout", "0.0", "--emb-dropout", "0.0", "--attention-dropout", "0.0", "--activation-dropout", "0.0",
        "--optimizer", "adam", "--adam-betas", "(0.9, 0.99)", "--adam-eps", "1e-6", "--clip-norm", "1.0", "--lr", "1e-3",
        "--lr-scheduler", "fixed", "--batch-size", "4", "--seed", "1", "--num-workers", "0", "--log-format", "none",
        "--disable-validation", "--no-save", "--distributed-world-size", "1", "--cpu", "--max-update", "10",
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
    task.load_dataset("train")
    ds = task.dataset("train")
    losses = []
    for i in range(a.steps):
        batch = ds.collater([ds[4 * i + k] for k in range(4)])
        out = trainer.train_step([batch])
        losses.append(float(out["loss"]))
    print(json.dumps({"impl": a.impl, "losses": losses}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
