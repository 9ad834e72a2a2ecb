"""CPU tests of the fused optimizer tail's host side (``--ddp-backend b200``): shard geometry, ordered bucket launches,
and - through the PyTorch specification of the kernel (``UNICORE_B200_REFERENCE_TAIL=1``, gloo, 2 ranks) - the optimizer
and trainer plumbing around it: compact fp32 state, checkpoint gather / reshard, EMA on the shard, statistics summed
inside the tail, deferred overflow handling."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from test_cpu_training import losses_of, run_cli  # noqa: E402

from unicore.optim.fp16_optimizer import flatten_parameters  # noqa: E402
from unicore_b200.parallel.fused_tail import FusedTail, plan_buckets  # noqa: E402
from unicore_b200.parallel.reference_tail import PlainComm, _PlainBuffer  # noqa: E402


def fake_comm(rank, world):
    comm = PlainComm.__new__(PlainComm)
    comm.group, comm.device, comm.rank, comm.world = None, torch.device("cpu"), rank, world
    return comm


def test_tail_plan_gives_every_element_exactly_one_owner():
    """Buckets tile every arena; the ranks' slices tile every bucket on 16-byte boundaries; a rank's compact shard is
    its slices back to back; the plan never exceeds the kernel's range table and orders buckets back to front."""
    for world in (2, 3, 8):
        for numels in ([8], [1000000, 4096], [85056, 24, 1024 * 1024 + 8]):
            plans = []
            for rank in range(world):
                comm = fake_comm(rank, world)
                grads = [_PlainBuffer(torch.zeros(n, dtype=torch.bfloat16), rank, world) for n in numels]
                params = [_PlainBuffer(torch.zeros(n, dtype=torch.bfloat16), rank, world) for n in numels]
                plans.append(FusedTail(comm, grads, params, bucket_bytes=64 * 1024))
            t0 = plans[0]
            assert len(t0.buckets) <= 64
            for g, n in enumerate(numels):
                tiles = sorted((b.lo, b.hi) for b in t0.buckets if b.group == g)
                assert tiles[0][0] == 0 and tiles[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(tiles[:-1], tiles[1:]))
                owned = sorted(r for t in plans for r in t.owned_ranges(g))
                assert all(lo % 8 == 0 and hi % 8 == 0 for lo, hi in owned)
                assert owned[0][0] == 0 and owned[-1][1] == n
                assert all(a[1] == b[0] for a, b in zip(owned[:-1], owned[1:]))
                assert sum(t.compact_numels[g] for t in plans) == n
                # back to front: within a group the bucket order is descending in offset
                order = [b.lo for b in t0.buckets if b.group == g]
                assert order == sorted(order, reverse=True)
            for t in plans:  # compact offsets are the running sum of the rank's slices, per group
                run = [0] * len(numels)
                for b in t.buckets:
                    assert b.compact_off == run[b.group]
                    run[b.group] += b.own_hi - b.own_lo
            # every rank derives the same bucket list (it drives the collective order)
            assert all([(b.group, b.lo, b.hi) for b in t.buckets] == [(b.group, b.lo, b.hi) for b in t0.buckets]
                       for t in plans)
    # the plan grows its buckets instead of overflowing the range table
    assert len(plan_buckets([64 * 1024 * 1024], 2, 1024, 64)) <= 64


def test_compact_and_full_views_agree():
    world, n = 4, 4096 + 24
    full = torch.arange(n, dtype=torch.float32)
    pieces = []
    for rank in range(world):
        comm = fake_comm(rank, world)
        buf = lambda: [_PlainBuffer(torch.zeros(n, dtype=torch.bfloat16), rank, world)]  # noqa: E731
        tail = FusedTail(comm, buf(), buf(), bucket_bytes=2048)
        compact = tail.to_compact(full[: n - 3], 0)  # a master that is shorter than the padded arena
        assert compact.numel() == tail.compact_numels[0]
        pieces.append((tail, compact))
    rebuilt = torch.zeros(n)
    for tail, compact in pieces:
        off = 0
        for lo, hi in tail.owned_ranges(0):
            rebuilt[lo:hi] = compact[off:off + hi - lo]
            off += hi - lo
    assert torch.equal(rebuilt[: n - 3], full[: n - 3]) and float(rebuilt[n - 3:].abs().sum()) == 0


def test_engine_launches_buckets_strictly_in_index_order():
    """Overlap logic of ``--ddp-backend b200`` without a GPU: gradient-ready hooks count buckets down, a bucket is
    launched exactly once, only when it AND all earlier buckets are complete (every rank issues the same sequence of
    collectives), the last bucket is left to the tail kernel, ``no_sync`` micro-batches launch nothing, and a hook that
    fires twice does not release a bucket early."""
    from unicore_b200.parallel.symm_dp import SymmDataParallel

    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                torch.nn.Linear(16, 4))
    eng = SymmDataParallel.__new__(SymmDataParallel)
    torch.nn.Module.__init__(eng)
    comm = fake_comm(0, 2)
    eng.module, eng.comm, eng.world_size = model, comm, 2
    eng.bucket_bytes, eng.accumulate_grads, eng.want_fused_tail = 96 * 2, False, True
    eng._grad_buffers, eng._param_buffers, eng._hooks, eng._buckets, eng._param_buckets = [], [], [], [], {}
    eng._next, eng._started, eng._covers_all_params, eng.tail, eng._sq_slots, eng._sq_total = 0, False, False, None, None, None
    eng._seen = set()

    def alloc(pool):
        def fn(numel, dtype, device):
            buf = _PlainBuffer(torch.zeros(-(-numel // 8) * 8, dtype=dtype), 0, 2)
            pool.append(buf)
            return buf.tensor[:numel]
        return fn

    model.bfloat16()
    flats = flatten_parameters(list(model.parameters()), grad_alloc=alloc(eng._grad_buffers),
                               param_alloc=alloc(eng._param_buffers))
    enabled = []
    optimizer = types.SimpleNamespace(fp16_params=[{"params": flats}], args=types.SimpleNamespace(seed=1),
                                      enable_fused_tail=lambda engine, tail: enabled.append(tail) or True)
    ready, launched = set(), []

    def fake_launch(bucket):
        base = flats[0].grad.data_ptr()
        for p in model.parameters():
            off = (p.grad.data_ptr() - base) // p.grad.element_size()
            if off < bucket.hi and off + p.grad.numel() > bucket.lo:
                assert id(p) in ready, "bucket launched before one of its gradients was ready"
        launched.append(bucket.index)
        bucket.launched = True

    eng._launch = fake_launch
    eng.attach_optimizer(optimizer)
    assert enabled and eng.tail is enabled[0] and eng._covers_all_params
    n_buckets = len(eng._buckets)
    assert n_buckets == -(-eng._grad_buffers[0].tensor.numel() // 96) and n_buckets >= 4
    inner = eng._on_grad_ready

    def on_ready(p):
        ready.add(id(p))
        inner(p)
        inner(p)  # a re-entrant / duplicated hook must be harmless

    for h in eng._hooks:
        h.remove()
    eng._hooks = [p.register_post_accumulate_grad_hook(on_ready) for p in model.parameters()]
    x = torch.randn(5, 8).bfloat16()
    with eng.no_sync():
        model(x).sum().backward()
    assert launched == []  # accumulation micro-batch: no communication
    ready.clear()
    eng._reset_counters()
    model(x).sum().backward()
    assert launched == list(range(n_buckets - 1))  # in order, each once, the last one is the tail's
    assert eng.pending_buckets() == [n_buckets - 1]
    assert eng._buckets[0].lo > eng._buckets[-1].lo  # the end of the arena (last layers) goes first


TAIL_ENV = {"UNICORE_B200_REFERENCE_TAIL": "1"}


def run_tail_cli(extra, env_extra=None):
    old = {k: os.environ.get(k) for k in TAIL_ENV}
    os.environ.update(TAIL_ENV if env_extra is None else env_extra)
    try:
        return run_cli(extra, nproc=2)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("precision", [["--bf16"], ["--fp16", "--fp16-init-scale", "4", "--deferred-overflow-check"]])
def test_fused_tail_trains_like_the_replicated_fused_optimizer(precision):
    """2 ranks, gloo: --ddp-backend b200 (PyTorch specification of the tail kernel: reduce-scatter, norm + statistics
    exchange, clip, Adam + EMA on the shard, parameter all-gather) follows the same loss curve as c10d DDP with the
    replicated fused-Adam step, logs the same global statistics, and counts the same updates."""
    common = ["--disable-validation", "--no-save", "--max-update", "6", "--ema-decay", "0.99", "--bucket-cap-mb", "1"] + precision
    tail = run_tail_cli(["--ddp-backend", "b200"] + common)
    repl = run_tail_cli(["--ddp-backend", "c10d"] + common)
    assert "training on 2 devices" in tail
    a, b = losses_of(tail), losses_of(repl)
    assert len(a) == len(b) == 6 and a == pytest.approx(b, abs=3e-2)
    assert all("bsz=16" in l for l in tail.splitlines() if "train_inner" in l)
    gn = [float(l.split("gnorm=")[1].split(",")[0]) for l in tail.splitlines() if "train_inner" in l and "gnorm=" in l]
    gn_ref = [float(l.split("gnorm=")[1].split(",")[0]) for l in repl.splitlines() if "train_inner" in l and "gnorm=" in l]
    assert len(gn) == 6 and gn == pytest.approx(gn_ref, rel=5e-2)


def test_fused_tail_checkpoint_has_the_reference_schema_and_resumes(tmp_path):
    """The sharded state is re-assembled into the reference checkpoint layout (full-length moments, EMA), and a run
    resumed from it continues exactly like an uninterrupted one (state is re-sharded on load)."""
    save = str(tmp_path / "ck")
    base = ["--ddp-backend", "b200", "--bf16", "--ema-decay", "0.99", "--save-dir", save, "--tmp-save-dir", save,
            "--disable-validation", "--save-interval-updates", "3", "--synthetic-num-samples", "256", "--bucket-cap-mb", "1"]
    run_tail_cli(base + ["--max-update", "3"])
    ck = torch.load(os.path.join(save, "checkpoint_last.pt"), map_location="cpu", weights_only=False)
    opt = ck["last_optimizer_state"]
    assert ck["optimizer_history"][-1]["optimizer_name"] == "FP16Optimizer" and ck["optimizer_history"][-1]["num_updates"] == 3
    n_model = sum(v.numel() for k, v in ck["model"].items() if k != "lm_head.weight")
    flat = sum(opt["state"][i]["exp_avg"].numel() for i in (0, 1))
    assert flat >= n_model and flat - n_model < 200           # full length, not one rank's shard
    assert all(float(opt["state"][i]["exp_avg_sq"].abs().sum()) > 0 for i in (0, 1))
    # every region of the moments was written by its owner: no large zero holes (a missing shard would be one)
    for i in (0, 1):
        v = opt["state"][i]["exp_avg_sq"]
        if v.numel() >= 64:
            halves = v[: v.numel() // 2], v[v.numel() // 2:]
            assert all(float((h != 0).float().mean()) > 0.5 for h in halves)
    assert set(ck["ema"]) == {"params", "decay"}
    ema_w, model_w = ck["ema"]["params"], ck["model"]
    moved = [k for k in model_w if model_w[k].dtype.is_floating_point and not torch.equal(ema_w[k].float(), model_w[k].float())]
    assert len(moved) > 0  # the EMA trails the weights everywhere it should
    resumed = losses_of(run_tail_cli(base + ["--max-update", "6"]))
    straight = losses_of(run_tail_cli(["--ddp-backend", "b200", "--bf16", "--ema-decay", "0.99", "--no-save",
                                       "--disable-validation", "--max-update", "6", "--synthetic-num-samples", "256",
                                       "--bucket-cap-mb", "1"]))
    assert len(resumed) == 3 and resumed == pytest.approx(straight[3:6], abs=2e-2)
