"""CPU unit tests of helper layers that the end-to-end tests only touch indirectly: tensor-tree utilities,
argument evaluators, timing meters, metric logging helpers, checkpoint file helpers, the buffered iterator and
rendezvous inference from the environment."""
import argparse
import os
import time

import pytest
import torch

from unicore import checkpoint_utils, metrics, utils
from unicore.data import iterators
from unicore.distributed import utils as dist_utils
from unicore.logging import meters


def test_tensor_tree_helpers():
    x = torch.arange(2 * 3 * 4 * 5).view(2, 3, 4, 5).float()
    assert utils.permute_final_dims(x, [1, 0]).shape == (2, 3, 5, 4)
    assert torch.equal(utils.permute_final_dims(x, [2, 0, 1]), x.permute(0, 3, 1, 2))
    assert utils.flatten_final_dims(x, 2).shape == (2, 3, 20)
    mask = torch.tensor([[1.0, 1.0, 0.0]])
    val = torch.tensor([[2.0, 4.0, 100.0]])
    assert utils.masked_mean(mask, val, dim=-1).item() == pytest.approx(3.0)
    assert torch.equal(utils.one_hot(torch.tensor([0, 2]), 3), torch.tensor([[1.0, 0, 0], [0, 0, 1.0]]))
    data = torch.arange(24).view(2, 3, 4)
    inds = torch.tensor([[0, 2], [1, 1]])
    got = utils.batched_gather(data, inds, dim=1, num_batch_dims=1)
    assert got.shape == (2, 2, 4) and torch.equal(got[1, 0], data[1, 1]) and torch.equal(got[0, 1], data[0, 2])
    with pytest.raises(ValueError):
        utils.batched_gather(data, inds, dim=0, num_batch_dims=1)
    tree = {"a": torch.ones(2), "b": [torch.zeros(1), (torch.ones(1),)], "c": {"d": torch.full((1,), 3.0)}}
    doubled = utils.tree_map(lambda t: t * 2, tree, torch.Tensor)
    assert doubled["a"].sum() == 4 and doubled["b"][1][0].item() == 2 and doubled["c"]["d"].item() == 6
    with pytest.raises(ValueError):
        utils.tree_map(lambda t: t, {"a": 3}, torch.Tensor)
    stacked = utils.dict_multimap(torch.stack, [{"x": torch.ones(2), "n": {"y": torch.zeros(1)}}] * 3)
    assert stacked["x"].shape == (3, 2) and stacked["n"]["y"].shape == (3, 1)
    moved = utils.apply_to_sample(lambda t: t + 1, {"k": [torch.zeros(1), (torch.zeros(1), 5)], "s": "text"})
    assert moved["k"][0].item() == 1 and moved["k"][1][0].item() == 1 and moved["k"][1][1] == 5 and moved["s"] == "text"
    assert utils.apply_to_sample(lambda t: t, {}) == {}


def test_argument_evaluators_and_activations():
    assert utils.csv_str_list("a,b") == ["a", "b"]
    assert utils.eval_str_list("[1, 2.5]") == [1.0, 2.5] and utils.eval_str_list("3", int) == [3]
    assert utils.eval_str_list("2**7", int) == [128] and utils.eval_str_list(None) is None
    assert utils.eval_str_dict("{'a': 1}") == {"a": 1} and utils.eval_str_dict(None) is None
    assert utils.eval_bool("True") is True and utils.eval_bool("0") is False and utils.eval_bool(None, True) is True
    assert utils.eval_bool("not_a_name", default=False) is False
    names = utils.get_available_activation_fns()
    assert {"relu", "gelu", "tanh", "linear"} <= set(names)
    t = torch.linspace(-2, 2, 9)
    assert torch.allclose(utils.get_activation_fn("gelu")(t), torch.nn.functional.gelu(t))
    assert torch.equal(utils.get_activation_fn("linear")(t), t)
    with pytest.raises(RuntimeError):
        utils.get_activation_fn("swishish")
    assert utils.has_parameters(torch.nn.Linear(2, 2)) and not utils.has_parameters(torch.nn.ReLU())
    state = utils.get_rng_state()
    a = torch.rand(3)
    utils.set_rng_state(state)
    assert torch.equal(a, torch.rand(3))


def test_clip_grad_norm_returns_preclip_norm_and_scales():
    p = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))]
    p[0].grad = torch.full((4,), 3.0)
    p[1].grad = torch.full((3,), 4.0)
    expect = (4 * 9 + 3 * 16) ** 0.5
    norm = utils.clip_grad_norm_(p, 1.0)
    assert float(norm) == pytest.approx(expect, rel=1e-6)
    after = torch.cat([q.grad for q in p]).norm().item()
    assert after == pytest.approx(1.0, rel=1e-4)
    p[0].grad = torch.full((4,), 0.01)
    p[1].grad = None
    assert float(utils.clip_grad_norm_(p, 1.0)) == pytest.approx(0.02, rel=1e-5)
    assert p[0].grad[0].item() == pytest.approx(0.01)  # below the threshold: untouched
    assert float(utils.clip_grad_norm_([torch.nn.Parameter(torch.zeros(1))], 1.0)) == 0.0


def test_time_meters():
    tm = meters.TimeMeter()
    tm.update(5)
    tm.update(5)
    time.sleep(0.02)
    assert tm.n == 10 and tm.elapsed_time >= 0.02 and 0 < tm.avg < 10 / 0.02
    saved = tm.state_dict()
    restored = meters.TimeMeter()
    restored.load_state_dict(saved)
    assert restored.n == 10 and restored.elapsed_time >= saved["init"] >= 0.02
    sw = meters.StopwatchMeter(round=3)
    sw.stop()  # never started: ignored
    assert sw.sum == 0 and sw.n == 0
    hooked = []
    sw.start()
    time.sleep(0.01)
    sw.stop(n=2, prehook=lambda: hooked.append(1))
    assert sw.n == 2 and sw.sum >= 0.01 and hooked == [1] and sw.avg == pytest.approx(sw.sum / 2)
    assert meters.safe_round(3.14159, 2) == 3.14 and meters.safe_round(torch.tensor(2.71828), 1) == 2.7
    assert meters.safe_round("n/a", 2) == "n/a"


def test_metric_logging_helpers():
    metrics.reset()
    with metrics.aggregate("probe") as agg:
        assert agg in metrics.get_active_aggregators()
        metrics.log_scalar("loss", 2.0, weight=2)
        metrics.log_scalar("loss", 4.0, weight=2)
        metrics.log_derived("double_loss", lambda m: (m["loss"].avg or 0) * 2)
        metrics.log_speed("wps", 100)
        metrics.log_start_time("wall")
        metrics.log_stop_time("wall", weight=1)
        metrics.log_custom(meters.AverageMeter, "custom", 7.0)
    vals = metrics.get_smoothed_values("probe")
    assert vals["loss"] == pytest.approx(3.0) and vals["double_loss"] == pytest.approx(6.0) and vals["custom"] == 7.0
    assert metrics.get_meter("probe", "wps") is not None and "wall" in metrics.get_meters("probe")
    metrics.reset_meter("probe", "loss")
    assert metrics.get_meter("probe", "loss").count == 0
    metrics.reset_meters("probe")
    assert metrics.get_smoothed_values("probe")["custom"] is None  # an empty average has no value
    metrics.reset()


def test_checkpoint_file_helpers(tmp_path):
    d = tmp_path / "ck"
    checkpoint_utils.verify_checkpoint_directory(str(d))
    assert d.is_dir() and not (d / "dummy").exists()
    for name in ("checkpoint1.pt", "checkpoint10.pt", "checkpoint2.pt", "checkpoint_3_40.pt", "checkpoint_last.pt"):
        checkpoint_utils.torch_persistent_save({"name": name, "args": argparse.Namespace(lr=1.0)}, str(d / name))
    assert not any(p.suffix == ".tmp" for p in d.iterdir())
    epochs = checkpoint_utils.checkpoint_paths(str(d))
    assert [os.path.basename(p) for p in epochs] == ["checkpoint10.pt", "checkpoint2.pt", "checkpoint1.pt"]
    inter = checkpoint_utils.checkpoint_paths(str(d), pattern=r"checkpoint_\d+_(\d+)\.pt")
    assert [os.path.basename(p) for p in inter] == ["checkpoint_3_40.pt"]
    assert checkpoint_utils.checkpoint_paths(str(tmp_path / "missing")) == []
    state = checkpoint_utils.load_checkpoint_to_cpu(str(d / "checkpoint2.pt"), arg_overrides={"lr": 0.5})
    assert state["name"] == "checkpoint2.pt" and state["args"].lr == 0.5


def test_buffered_iterator_preserves_order_and_length():
    class Source:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

        def __iter__(self):
            for i in range(self.n):
                yield {"x": torch.full((2,), float(i))}

    buf = iterators.BufferedIterator(3, Source(7))
    assert len(buf) == 7
    got = [int(b["x"][0]) for b in buf]
    assert got == list(range(7))
    short = iterators.BufferedIterator(2, iterators.CountingIterator(list(range(10)))).take(4)
    assert len(short) == 4 and list(short) == [0, 1, 2, 3]


def test_rendezvous_is_inferred_from_the_environment(monkeypatch):
    def fresh(world=1):
        return argparse.Namespace(distributed_init_method=None, distributed_world_size=world, distributed_rank=0,
                                  distributed_port=-1, device_id=0, distributed_no_spawn=False, cpu=True)

    for var in ("MASTER_ADDR", "MASTER_PORT", "WORLD_SIZE", "RANK", "LOCAL_RANK", "SLURM_STEP_NODELIST", "SLURM_JOB_NODELIST"):
        monkeypatch.delenv(var, raising=False)
    args = fresh()
    dist_utils.infer_init_method(args)
    assert args.distributed_init_method is None  # single process: nothing to do
    monkeypatch.setenv("MASTER_ADDR", "127.0.0.1")
    monkeypatch.setenv("MASTER_PORT", "29533")
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("LOCAL_RANK", "1")
    args = fresh()
    dist_utils.infer_init_method(args)
    assert args.distributed_init_method == "env://" and args.distributed_world_size == 4
    assert args.distributed_rank == 3 and args.device_id == 1 and args.distributed_no_spawn
    for var in ("MASTER_ADDR", "MASTER_PORT", "WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(var)
    args = fresh(world=2)
    dist_utils.infer_init_method(args)
    assert args.distributed_init_method.startswith("tcp://") and "127.0.0.1" in args.distributed_init_method
    explicit = fresh()
    explicit.distributed_init_method = "tcp://10.0.0.1:1234"
    dist_utils.infer_init_method(explicit)
    assert explicit.distributed_init_method == "tcp://10.0.0.1:1234"
