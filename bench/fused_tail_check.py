#!/usr/bin/env python3
"""The fused optimizer tail kernel against its PyTorch specification, and b200 training against c10d (torchrun, >= 2 GPUs).

Part 1 (kernel parity): identical synthetic state on two ``FusedTail`` plans over the same geometry - one updated by
``csrc/comm/fused_step.cu`` (with real reduce-scatter bucket kernels in front of it and a random subset of buckets
left pending for the tail), the other by ``reference_tail`` (NCCL all-reduces + the formulas written out).  Compared:
16-bit parameters on every rank, compact fp32 master / moments, EMA slices, the state vector (norm, multiplier,
overflow flag), the statistics sums, zeroed gradients.  Cases: fp16 / bf16, clip on / off, overflow injection, EMA,
several updates in a row (parity bits, slot reuse).

Part 2 (training parity): the same tiny BERT trained for a few updates with ``--ddp-backend b200`` (fused tail) and
with ``--ddp-backend c10d`` (NCCL DDP + replicated fused Adam): losses, final parameters.

Rank 0 prints one JSON line per case and a summary line.
"""
import argparse
import importlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "examples"))


def kernel_case(comm, dtype, n_groups, numel, bucket_bytes, steps, clip, ema, inject_overflow, seed):
    import torch
    import torch.distributed as dist

    from unicore_b200.parallel.comm import TAG_BUCKET
    from unicore_b200.parallel.fused_tail import FusedTail, adam_hyper
    from unicore_b200.parallel.reference_tail import PlainComm, _PlainBuffer, reference_tail

    rank, world, dev = comm.rank, comm.world, comm.device
    numels = [numel, max(8, (numel // 5) // 8 * 8)][:n_groups]
    gk = [comm.allocate(n, dtype) for n in numels]
    pk = [comm.allocate(n, dtype) for n in numels]
    tail_k = FusedTail(comm, gk, pk, bucket_bytes, seed=seed)
    plain = PlainComm(comm.group, dev)
    gr = [_PlainBuffer(torch.zeros(n, dtype=dtype, device=dev), rank, world) for n in numels]
    pr = [_PlainBuffer(torch.zeros(n, dtype=dtype, device=dev), rank, world) for n in numels]
    tail_r = FusedTail(plain, gr, pr, bucket_bytes, seed=seed)
    gen = torch.Generator(device=dev).manual_seed(1000 + seed)       # same on every rank: the initial state
    lgen = torch.Generator(device=dev).manual_seed(77 + 13 * rank + seed)  # rank specific: the local gradients

    def compact_state(tail):
        ms, avs, sqs = [], [], []
        for g, n in enumerate(numels):
            full = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(5 + g + seed)) * 0.05
            ms.append(tail.to_compact(full, g))
            avs.append(torch.zeros_like(ms[-1]))
            sqs.append(torch.zeros_like(ms[-1]))
        return ms, avs, sqs

    mk, ak, sk = compact_state(tail_k)
    mr, ar, sr = compact_state(tail_r)
    ema_k = [torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(9 + g)) * 0.05 for g, n in enumerate(numels)] if ema else None
    ema_r = [e.clone() for e in ema_k] if ema else None
    worst = {"param": 0.0, "master": 0.0, "moment": 0.0, "ema": 0.0, "state": 0.0, "stats": 0.0, "grad_left": 0.0}
    flags = []
    for step in range(1, steps + 1):
        hyp = [adam_hyper(1e-3, 0.9, 0.98, 1e-6, step, True, 0.01 if g == 0 else 0.0) for g in range(len(numels))]
        for g, n in enumerate(numels):
            local = (torch.randn(n, device=dev, generator=lgen) * 0.02).to(dtype)
            if inject_overflow and step == 2 and rank == world - 1 and g == 0:
                local[n // 3] = float("inf")
            gk[g].tensor.copy_(local)
            gr[g].tensor.copy_(local)
        stats = torch.tensor([3.0 + rank, 100.0 * (rank + 1), float(step)], dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        # a pseudo-random prefix of the buckets is reduce-scattered by the bucket kernels (in order), the rest is pending
        n_b = len(tail_k.buckets)
        launched = (step * 7 + seed) % (n_b + 1)
        launched = min(launched, n_b - 1)
        for b in tail_k.buckets[:launched]:
            comm.reduce_scatter(gk[b.group], b.lo, b.hi - b.lo, scale=1.0 / world, sq_out=tail_k.sq_slots(b.index),
                                tag=TAG_BUCKET + b.index)
        kw = dict(hypers=hyp, factor=float(world) / 4.0, max_norm=clip, clip_eps=0.0 if dtype == torch.float16 else 1e-6,
                  ema_decay=0.99, stats_src=stats, denom_index=1)
        st_k = tail_k.launch(masters=mk, exp_avgs=ak, exp_avg_sqs=sk, emas=ema_k,
                             pending=[b.index for b in tail_k.buckets[launched:]], **kw).clone()
        stats_k = tail_k.stats_dst[:3].clone()
        st_r = reference_tail(tail_r, masters=mr, exp_avgs=ar, exp_avg_sqs=sr, emas=ema_r, **kw).clone()
        stats_r = tail_r.stats_dst[:3].clone()
        torch.cuda.synchronize()
        comm.check_health()
        flags.append((float(st_k[2]), float(st_r[2])))
        rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-12)) if a.numel() else 0.0  # noqa: E731
        if float(st_r[2]) == 0.0:
            worst["state"] = max(worst["state"], rel(st_k[:2], st_r[:2]))
        worst["stats"] = max(worst["stats"], rel(stats_k, stats_r))
        for g in range(len(numels)):
            worst["param"] = max(worst["param"], float((pk[g].tensor.float() - pr[g].tensor.float()).abs().max()))
            worst["master"] = max(worst["master"], float((mk[g] - mr[g]).abs().max()) if mk[g].numel() else 0.0)
            worst["moment"] = max(worst["moment"], rel(ak[g], ar[g]), rel(sk[g], sr[g]))
            worst["grad_left"] = max(worst["grad_left"], float(gk[g].tensor.float().abs().max()))
            if ema:
                for lo, hi in tail_k.owned_ranges(g):
                    worst["ema"] = max(worst["ema"], float((ema_k[g][lo:hi] - ema_r[g][lo:hi]).abs().max()))
    t = torch.tensor([worst[k] for k in sorted(worst)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = dict(zip(sorted(worst), [float(x) for x in t]))
    out["overflow_flags_match"] = all(a == b for a, b in flags)
    out["overflow_seen"] = any(a != 0.0 for a, _ in flags)
    del gk, pk, tail_k
    torch.cuda.empty_cache()
    return out


def train(backend, steps, precision, rank, local_rank, world, ema):
    import torch

    from unicore import options, tasks, utils
    from unicore.trainer import Trainer

    flags = [
        "--task", "synthetic_mlm", "--loss", "masked_lm", "--arch", "bert_base", "--encoder-layers", "2",
        "--encoder-embed-dim", "128", "--encoder-ffn-embed-dim", "256", "--encoder-attention-heads", "2",
        "--synthetic-vocab-size", "509", "--synthetic-seq-len", "64", "--max-seq-len", "64", "--optimizer", "adam",
        "--lr", "1e-3", "--lr-scheduler", "fixed", "--weight-decay", "0.01", "--clip-norm", "1.0", "--max-update", "100",
        "--batch-size", "8", "--seed", "11", "--no-save", "--disable-validation", "--log-format", "none",
        "--distributed-world-size", str(world), "--distributed-rank", str(rank), "--device-id", str(local_rank),
        "--ddp-backend", backend, "--bucket-cap-mb", "1", "--dropout", "0.0", "--attention-dropout", "0.0",
        "--emb-dropout", "0.0", "--" + precision,
    ]
    if precision == "fp16":
        flags += ["--fp16-init-scale", "4", "--deferred-overflow-check"]
    if ema:
        flags += ["--ema-decay", "0.99"]
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
    losses = []
    for i in range(steps):
        rows = [ds[(i * world + rank) * 8 + k] for k in range(8)]
        out = trainer.train_step([utils.move_to_cuda(ds.collater(rows))])
        losses.append(float(out["loss"]) if out is not None else float("nan"))
    if hasattr(trainer.optimizer, "resolve_pending_overflow"):
        trainer.optimizer.resolve_pending_overflow()
    trainer.consolidate_optimizer_state()
    torch.cuda.synchronize()
    tail = bool(getattr(trainer.optimizer, "uses_fused_tail", False))
    params = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
    opt = trainer.optimizer.state_dict() if (rank == 0 or not tail) else None
    ema_flat = None
    if trainer.ema is not None:
        ema_flat = torch.cat([v.detach().float().reshape(-1) for v in trainer.ema.model_ema.parameters()]).cpu()
    return tail, losses, params, opt, ema_flat, trainer.get_num_updates()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--skip-kernel", action="store_true")
    ap.add_argument("--skip-train", action="store_true")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend="nccl", init_method="env://", world_size=world, rank=rank)
    dist.all_reduce(torch.zeros(1, device="cuda"))
    failures = 0

    def emit(rec):
        if rank == 0:
            print(json.dumps(rec), flush=True)

    if not a.skip_kernel:
        from unicore_b200.parallel.comm import SymmComm

        comm = SymmComm()
        emit({"communicator": {"provider": comm.provider, "world": world, "nvls": bool(comm.flags.multicast_ptr)}})
        cases = [
            dict(dtype="bfloat16", n_groups=1, numel=1 << 20, bucket_bytes=256 << 10, steps=3, clip=1.0, ema=False, inject_overflow=False),
            dict(dtype="float16", n_groups=2, numel=(1 << 20) + 4096 + 8, bucket_bytes=200 << 10, steps=4, clip=0.05, ema=True, inject_overflow=False),
            dict(dtype="float16", n_groups=2, numel=300008, bucket_bytes=64 << 10, steps=3, clip=1.0, ema=True, inject_overflow=True),
            dict(dtype="bfloat16", n_groups=2, numel=8 * 1000 + 8, bucket_bytes=4 << 10, steps=3, clip=0.0, ema=True, inject_overflow=False),
        ]
        for i, c in enumerate(cases):
            kw = dict(c)
            kw["dtype"] = getattr(torch, c["dtype"])
            res = kernel_case(comm, seed=i, **kw)
            # 2 GPUs reduce by peer loads in rank order with fp32 accumulation: bit-compatible with the specification.
            # With NVLS (world > 2) the switch adds in its own order before the single rounding to 16 bits, so a few
            # reduced gradients differ from the specification's by one unit in the last place of the 16-bit type
            # (2^-8 relative for bf16, 2^-11 for fp16), and so do the quantities derived from them.
            bf16 = c["dtype"] == "bfloat16"
            in_switch = world > 2 and bool(comm.flags.multicast_ptr)
            one_ulp = (2.0 ** -7 if bf16 else 2.0 ** -10) * 0.3  # |param| <= ~0.3
            tol = dict(master=2e-6, moment=1e-4, state=1e-4)
            if in_switch:
                tol = dict(master=5e-5, moment=4e-2, state=2e-3) if bf16 else dict(master=8e-6, moment=6e-3, state=3e-4)
            ok = (res["param"] <= 2 * one_ulp and res["master"] < tol["master"] and res["moment"] < tol["moment"]
                  and res["ema"] < 2e-6 and res["state"] < tol["state"] and res["stats"] < 1e-12
                  and res["grad_left"] == 0.0 and res["overflow_flags_match"]
                  and res["overflow_seen"] == c["inject_overflow"])
            failures += 0 if ok else 1
            emit({"case": "kernel", "config": c, "ok": ok, **res})

    if not a.skip_train:
        importlib.import_module("bert")
        for precision, ema in (("bf16", False), ("fp16", True)):
            t0, l0, p0, o0, e0, u0 = train("c10d", a.steps, precision, rank, local_rank, world, ema)
            t1, l1, p1, o1, e1, u1 = train("b200", a.steps, precision, rank, local_rank, world, ema)
            rec = {"case": "train", "precision": precision, "ema": ema, "tail_active": t1, "c10d_used_tail": t0,
                   "loss_c10d": l0, "loss_b200": l1, "updates": [u0, u1],
                   "max_param_diff": float((p0 - p1).abs().max())}
            if rank == 0:
                diffs = []
                for k in o0["state"]:
                    for name in ("exp_avg", "exp_avg_sq"):
                        x, y = o0["state"][k][name].float().cpu(), o1["state"][k][name].float().cpu()
                        rec.setdefault("state_numel", []).append([x.numel(), y.numel()])
                        diffs.append(float((x - y).abs().max() / (x.abs().max() + 1e-12)))
                rec["max_rel_moment_diff"] = max(diffs)
            if e0 is not None and e1 is not None:
                rec["max_ema_diff"] = float((e0 - e1).abs().max())
            worst = torch.tensor([rec["max_param_diff"]], device="cuda")
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            rec["max_param_diff"] = float(worst)
            loss_gap = max(abs(x - y) for x, y in zip(l0, l1))
            # c10d averages 16-bit gradients with NCCL (16-bit ring adds), the tail accumulates in fp32: one rounding step
            ok = t1 and not t0 and loss_gap < 3e-2 and rec["max_param_diff"] < 4e-3 and u0 == u1
            if rank == 0:
                ok = ok and rec["max_rel_moment_diff"] < 5e-2 and all(x == y for x, y in rec["state_numel"])
            failures += 0 if ok else 1
            rec["ok"] = bool(ok)
            emit(rec)

    f = torch.tensor([failures], device="cuda")
    dist.all_reduce(f, op=dist.ReduceOp.MAX)
    emit({"summary": "fused_tail_check", "world": world, "failures": int(f)})
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
