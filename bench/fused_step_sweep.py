#!/usr/bin/env python3
"""Gradient reduction + optimizer update per bucket size: three ways of doing the same work (BASELINE config 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29532 bench/fused_step_sweep.py [--max-mb 256]

For every size (1 MB .. --max-mb of 16-bit gradients, x4 steps) it times, with CUDA events, max over ranks:

  nccl      NCCL all-reduce + the replicated fused Adam kernel          (what --ddp-backend c10d does)
  symm      our all-reduce kernel (auto: one-shot / two-shot / NVLS) + the replicated fused Adam kernel
  sharded   our reduce-scatter half + Adam on the 1/N shard with the parameter all-gather in the kernel's own
            multimem.st / peer stores                                    (UNICORE_B200_SHARD_OPTIMIZER=2 path)

and checks that the three leave the same 16-bit parameters behind.  Rank 0 prints one JSON object per size.
EXPERIMENTAL: the sharded kernels have not run on hardware yet.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=256)
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float16"])
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", init_method="env://")
    dist.all_reduce(torch.zeros(1, device="cuda"))

    from unicore import ops
    from unicore_b200.parallel.symm_dp import ShardedAdamStepper, SymmAllReduce, symm_available

    assert symm_available(), "symmetric memory / native kernels unavailable"
    dtype = getattr(torch, args.dtype)
    red = SymmAllReduce()
    hyper = dict(lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, step=1, bias_correction=True, weight_decay=0.01)
    mb = 1
    while mb <= args.max_mb:
        n = mb * 1024 * 1024 // 2
        grads = red.allocate(n, dtype)      # symmetric gradient arena
        params = red.allocate(n, dtype)     # symmetric parameter arena
        torch.manual_seed(1)
        init = (torch.randn(n, device="cuda") * 0.02)
        local_grad = (torch.randn(n, device="cuda", generator=torch.Generator("cuda").manual_seed(100 + rank)) * 1e-2).to(dtype)
        state = {}

        def reset(which):
            state[which] = dict(master=init.clone(), m=torch.zeros(n, device="cuda"), v=torch.zeros(n, device="cuda"))
            params.tensor.copy_(init.to(dtype))
            grads.tensor.copy_(local_grad)

        def work_item(st, p_half, g):
            return [dict(p=st["master"], g=g, m=st["m"], v=st["v"], p_half=p_half, **hyper)]

        def run_nccl():
            grads.tensor.copy_(local_grad)
            dist.all_reduce(grads.tensor)
            ops.fused_adam(work_item(state["nccl"], params.tensor, grads.tensor), grad_scale=float(world))

        def run_symm():
            grads.tensor.copy_(local_grad)
            red(grads, 0, n, scale=1.0 / world)
            ops.fused_adam(work_item(state["symm"], params.tensor, grads.tensor), grad_scale=1.0)

        class _Flat:  # the stepper only needs .data_ptr() and .grad
            def __init__(self):
                self.grad = grads.tensor

            def data_ptr(self):
                return params.tensor.data_ptr()

        stepper = ShardedAdamStepper(red, [params])
        epv = 8
        nvec = n // epv
        per = -(-nvec // world)
        stepper.bucket_slices = {grads.tensor.data_ptr(): [(min(nvec, per * rank) * epv, min(nvec, per * (rank + 1)) * epv)]}
        flat = _Flat()

        def run_sharded():
            grads.tensor.copy_(local_grad)
            red(grads, 0, n, scale=1.0 / world, scatter_only=True)
            st = state["sharded"]
            stepper.step(flat, st["master"], st["m"], st["v"], grad_scale=1.0, **hyper)

        results, finals = {}, {}
        for name, fn in (("nccl", run_nccl), ("symm", run_symm), ("sharded", run_sharded)):
            reset(name)
            fn()
            torch.cuda.synchronize()
            dist.barrier()
            finals[name] = params.tensor.float().clone()
            reset(name)
            results[name] = timed(fn, iters=10 if mb <= 64 else 5)
        diff = max(float((finals["nccl"] - finals[k]).abs().max()) for k in ("symm", "sharded"))
        d = torch.tensor([diff], device="cuda")
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"grad_mb": mb, "world": world, "dtype": args.dtype, "ms": results,
                              "max_param_diff_vs_nccl": float(d.item())}))
        del grads, params
        torch.cuda.empty_cache()
        mb *= 4
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
