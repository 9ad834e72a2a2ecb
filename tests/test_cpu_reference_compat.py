"""Compatibility proofs against the reference implementation: its example plug-in
runs unmodified through this framework's ``--user-dir``, and checkpoints travel in both directions.  The reference is
used as installed in ``baseline/_ref`` (or ``/root/reference``); the tests skip when neither is present."""
import json
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY = sys.executable


def reference_example_dir():
    for base in ("/root/reference/examples", os.path.join(ROOT, "baseline", "_ref", "examples")):
        if os.path.isfile(os.path.join(base, "bert", "task.py")):
            return os.path.join(base, "bert")
    return None


def test_reference_bert_example_runs_unmodified_through_user_dir(tmp_path):
    """``--user-dir <reference>/examples/bert``: the reference's own task.py / model.py (its imports, its registry
    decorators, its dataset pipeline, its BERT built from ``unicore.modules``) train, validate and checkpoint here."""
    pytest.importorskip("tokenizers")
    example = reference_example_dir()
    if example is None:
        pytest.skip("reference examples not available")
    rng = random.Random(0)
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa", "lambda", "mu"]
    for split, n in (("train", 96), ("valid", 24)):
        with open(tmp_path / (split + ".txt"), "w") as f:
            for _ in range(n):
                f.write(" ".join(rng.choice(words) for _ in range(rng.randint(6, 14))) + " .\n")
    data = tmp_path / "data"
    prep = os.path.join(ROOT, "examples", "bert", "example_data", "prepare_data.py")
    out = subprocess.run([PY, prep, "--train", str(tmp_path / "train.txt"), "--valid", str(tmp_path / "valid.txt"),
                          "--out", str(data), "--build-dict", "--min-count", "1", "--format", "records"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    save = str(tmp_path / "ck")
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    cmd = [PY, os.path.join(ROOT, "unicore_cli", "train.py"), str(data), "--user-dir", example,
           "--task", "bert", "--loss", "masked_lm", "--arch", "bert_base", "--encoder-layers", "2",
           "--encoder-embed-dim", "32", "--encoder-ffn-embed-dim", "64", "--encoder-attention-heads", "4",
           "--max-seq-len", "32", "--optimizer", "adam", "--lr", "1e-3", "--lr-scheduler", "fixed", "--batch-size", "8",
           "--max-update", "6", "--log-format", "simple", "--log-interval", "1", "--num-workers", "0", "--cpu",
           "--valid-subset", "valid", "--validate-interval-updates", "3", "--save-interval-updates", "3",
           "--save-dir", save, "--distributed-world-size", "1", "--seed", "1"]
    run = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-4000:]
    losses = [float(l.split("loss=")[1].split(",")[0]) for l in run.stdout.splitlines() if "train_inner" in l and "loss=" in l]
    assert len(losses) >= 5 and all(v == v for v in losses) and "valid" in run.stdout
    assert os.path.isfile(os.path.join(save, "checkpoint_last.pt"))
    # it really was the reference's plug-in: its modules come from the reference tree, the framework from this repo
    assert example.startswith(("/root/reference", os.path.join(ROOT, "baseline", "_ref")))


def _interop(impl, init, *extra):
    tool = os.path.join(ROOT, "tools", "checkpoint_interop.py")
    out = subprocess.run([PY, tool, "--impl", impl, "--init", init] + list(extra), env=dict(os.environ, OMP_NUM_THREADS="1"),
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_checkpoints_travel_between_the_reference_and_this_framework(tmp_path):
    """bf16 + weight decay (two flat optimizer groups) + clipping + EMA on the CPU: each side trains 3 updates, saves,
    and trains 2 more; the OTHER side loads that file through its own ``Trainer.load_checkpoint`` and must reproduce
    the 2 losses, the update count, the learning rate and the EMA (flat layout, reference
    ``unicore/optim/fp16_optimizer.py:54-121``; schema ``unicore/trainer.py:224-284``)."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "unicore")):
        pytest.skip("reference not installed under baseline/_ref")
    init = str(tmp_path / "init.pt")
    sides = {}
    for impl in ("reference", "ours"):
        rec = _interop(impl, init, "--steps", "3", "--save", str(tmp_path / (impl + ".pt")), "--more", "2")
        if "unavailable" in rec:
            pytest.skip(rec["unavailable"])
        sides[impl] = rec
    a, b = sides["reference"], sides["ours"]
    assert a["saved_keys"] == b["saved_keys"] and a["saved_opt_state_numel"] == b["saved_opt_state_numel"]
    assert a["losses_before"] == pytest.approx(b["losses_before"], abs=2e-3)
    for loader, writer in (("ours", "reference"), ("reference", "ours")):
        got = _interop(loader, init, "--load", str(tmp_path / (writer + ".pt")), "--more", "2")
        assert got["loaded_updates"] == 3 and got["loaded_epoch"] == 1
        assert got["losses_after"] == pytest.approx(sides[writer]["losses_after"], abs=3e-3), (loader, writer)
        assert got["ema_checksum"] == pytest.approx(sides[writer]["ema_checksum"], rel=1e-5)


def test_portable_unimol_plugin_gives_the_same_losses_under_both_frameworks(tmp_path):
    """``examples/unimol_portable`` only uses the public Uni-Core API: the identical plug-in code trains under the
    reference and under this framework, from the same weights, to the same losses."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "unicore")):
        pytest.skip("reference not installed under baseline/_ref")
    tool = os.path.join(ROOT, "tools", "unimol_portable_check.py")
    init = str(tmp_path / "u.pt")
    runs = {}
    for impl in ("ours", "reference"):
        out = subprocess.run([PY, tool, "--impl", impl, "--init", init, "--steps", "3"], env=dict(os.environ, OMP_NUM_THREADS="1"),
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-3000:]
        runs[impl] = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        if "unavailable" in runs[impl]:
            pytest.skip(runs[impl]["unavailable"])
    assert len(runs["ours"]["losses"]) == 3 and runs["ours"]["losses"][-1] < runs["ours"]["losses"][0]
    assert runs["ours"]["losses"] == pytest.approx(runs["reference"]["losses"], abs=2e-3)
