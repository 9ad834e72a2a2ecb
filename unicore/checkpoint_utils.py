"""Checkpoint naming, atomic writes, asynchronous fan-out copies, pruning and resume policy.

Behaviour follows reference ``unicore/checkpoint_utils.py`` (``save_checkpoint:83``,
``ckp_copy_fun:23``, ``load_checkpoint:165``, ``load_checkpoint_to_cpu:244``,
``checkpoint_paths:261``, ``torch_persistent_save:280``, ``verify_checkpoint_directory:300``):
ONE file is written (rank 0, to ``--tmp-save-dir``), a 1-thread pool then copies it to every
applicable name (``checkpoint{E}.pt``, ``checkpoint_{E}_{U}.pt``, ``checkpoint_best.pt``,
``checkpoint.best_{metric}_{val:.2f}.pt``, ``checkpoint_last.pt``) and prunes old files.
The file schema is SURVEY.md Appendix B (cross-loadable with the reference).
"""
import ast
import collections
import logging
import os
import re
import shutil
import traceback

import torch

logger = logging.getLogger(__name__)


class _BestTracker:
    """Best validation score seen by this process (the reference keeps it in function attributes)."""

    value = None

    @classmethod
    def update(cls, val, maximize):
        if val is None:
            return
        if cls.value is None:
            cls.value = val
        else:
            cls.value = max(val, cls.value) if maximize else min(val, cls.value)

    @classmethod
    def reset(cls):
        cls.value = None


def _prune(args, end_of_epoch):
    def drop(paths):
        for old in paths:
            if os.path.lexists(old):
                os.remove(old)
                logger.info("removed {}".format(old))

    root = args.save_dir
    if not end_of_epoch and args.keep_interval_updates > 0:
        drop(checkpoint_paths(root, pattern=r"checkpoint_\d+_(\d+)\.pt")[args.keep_interval_updates:])
    if args.keep_last_epochs >= 0:
        drop(checkpoint_paths(root, pattern=r"checkpoint(\d+)\.pt")[args.keep_last_epochs:])
    if args.keep_best_checkpoints > 0:
        ranked = checkpoint_paths(
            root, pattern=r"checkpoint\.best_{}_(\d+\.?\d*)\.pt".format(args.best_checkpoint_metric)
        )
        if not args.maximize_best_checkpoint_metric:
            ranked = ranked[::-1]
        drop(ranked[args.keep_best_checkpoints:])


def ckp_copy_fun(src, checkpoints, end_of_epoch, args):
    """Runs in the copy thread: replicate ``src`` under every name, drop the temp file, prune."""
    copied = False
    for dst in checkpoints:
        if os.path.abspath(src) == os.path.abspath(dst):
            continue
        try:
            logger.info("copy {} to {}".format(src, dst))
            shutil.copyfile(src, dst)
            copied = True
        except Exception:  # noqa: BLE001
            logger.info("copy failed, please copy it manaully")
    try:
        same_dir = os.path.abspath(args.tmp_save_dir) == os.path.abspath(args.save_dir)
        if not same_dir and copied and os.path.lexists(src):
            logger.info("removing temp file {} ...".format(src))
            os.remove(src)
        _prune(args, end_of_epoch)
    except Exception:  # noqa: BLE001
        logger.info("remove old ckps error")
    logger.info("finished async ckp saving.")


def save_checkpoint(args, trainer, epoch_itr, val_loss, ckp_copy_thread, do_save=True):
    from unicore import meters

    if trainer.data_parallel_rank == 0:
        os.makedirs(args.save_dir, exist_ok=True)

    previous_best = _BestTracker.value
    _BestTracker.update(val_loss, args.maximize_best_checkpoint_metric)
    save_checkpoint.best = _BestTracker.value  # attribute kept for API compatibility

    if args.no_save or not do_save:
        return
    if hasattr(trainer, "consolidate_optimizer_state"):
        trainer.consolidate_optimizer_state()  # every rank takes part; only the master writes below
    if not trainer.should_save_checkpoint_on_current_rank:
        return

    timer = meters.StopwatchMeter()
    timer.start()
    epoch = epoch_itr.epoch
    end_of_epoch = epoch_itr.end_of_epoch()
    updates = trainer.get_num_updates()
    logger.info("Preparing to save checkpoint for epoch {} @ {} updates".format(epoch, updates))

    def at_least_as_good(a, b):
        return a >= b if args.maximize_best_checkpoint_metric else a <= b

    is_new_best = val_loss is not None and (previous_best is None or at_least_as_good(val_loss, _BestTracker.value))
    suffix = trainer.checkpoint_suffix
    conds = collections.OrderedDict()
    conds["checkpoint{}{}.pt".format(epoch, suffix)] = (
        end_of_epoch and not args.no_epoch_checkpoints and epoch % args.save_interval == 0
    )
    conds["checkpoint_{}_{}{}.pt".format(epoch, updates, suffix)] = (
        not end_of_epoch and args.save_interval_updates > 0 and updates % args.save_interval_updates == 0
    )
    conds["checkpoint_best{}.pt".format(suffix)] = is_new_best
    if val_loss is not None and args.keep_best_checkpoints > 0:
        conds["checkpoint.best_{}_{:.2f}.pt".format(args.best_checkpoint_metric, val_loss)] = is_new_best
    conds["checkpoint_last{}.pt".format(suffix)] = not args.no_last_checkpoints

    extra_state = {"train_iterator": epoch_itr.state_dict(), "val_loss": val_loss}
    if _BestTracker.value is not None:
        extra_state["best"] = _BestTracker.value

    names = [fn for fn, cond in conds.items() if cond]
    if not names:
        return
    targets = [os.path.join(args.save_dir, fn) for fn in names]
    staging = os.path.join(args.tmp_save_dir, names[0])
    trainer.save_checkpoint(staging, extra_state)
    if ckp_copy_thread is not None:
        ckp_copy_thread.apply_async(ckp_copy_fun, (staging, targets, end_of_epoch, args))
    else:
        ckp_copy_fun(staging, targets, end_of_epoch, args)
    timer.stop()
    logger.info(
        "Saved checkpoint {} (epoch {} @ {} updates, score {}) (writing took {} seconds)".format(
            staging, epoch, updates, val_loss, timer.sum
        )
    )


def load_checkpoint(args, trainer, **passthrough_args):
    """Resume (or fine-tune) according to the checkpoint flags; returns ``(extra_state, epoch_itr)``."""
    reset_optimizer = args.reset_optimizer
    reset_lr_scheduler = args.reset_lr_scheduler
    optimizer_overrides = ast.literal_eval(args.optimizer_overrides)
    reset_meters = args.reset_meters
    reset_dataloader = args.reset_dataloader

    if args.finetune_from_model is not None and (
        reset_optimizer or reset_lr_scheduler or reset_meters or reset_dataloader
    ):
        raise ValueError(
            "--finetune-from-model can not be set together with either --reset-optimizer or "
            "reset_lr_scheduler or reset_meters or reset_dataloader"
        )

    suffix = trainer.checkpoint_suffix
    default_name = "checkpoint_last.pt"
    if args.restore_file == default_name:
        path = os.path.join(args.save_dir, "checkpoint_last{}.pt".format(suffix))
        first_launch = not os.path.exists(path)
        if args.finetune_from_model is not None and first_launch:
            if not os.path.exists(args.finetune_from_model):
                raise ValueError("--finetune-from-model {} does not exist".format(args.finetune_from_model))
            path = args.finetune_from_model
            reset_optimizer = reset_lr_scheduler = reset_meters = reset_dataloader = True
            logger.info(
                "loading pretrained model from {}: optimizer, lr scheduler, meters, dataloader will be reset".format(path)
            )
    elif suffix is not None:
        path = args.restore_file.replace(".pt", suffix + ".pt")
    else:
        path = args.restore_file

    if args.restore_file != default_name and args.finetune_from_model:
        raise ValueError(
            "--finetune-from-model and --restore-file (non-default value) can not be specified together: " + str(args)
        )

    extra_state, epoch_itr = trainer.load_checkpoint(
        path,
        reset_optimizer,
        reset_lr_scheduler,
        reset_dataloader,
        optimizer_overrides,
        reset_meters=reset_meters,
        **passthrough_args,
    )
    if extra_state is not None and "best" in extra_state and not reset_optimizer and not reset_meters:
        _BestTracker.value = extra_state["best"]
        save_checkpoint.best = extra_state["best"]
    return extra_state, epoch_itr


def load_checkpoint_to_cpu(path, arg_overrides=None, load_on_all_ranks=False):
    """Read a checkpoint onto the CPU (pickled ``args`` Namespace => ``weights_only=False``)."""
    with open(path, "rb") as f:
        state = torch.load(f, map_location=torch.device("cpu"), weights_only=False)
    if "args" in state and state["args"] is not None and arg_overrides is not None:
        for name, value in arg_overrides.items():
            setattr(state["args"], name, value)
    return state


def checkpoint_paths(path, pattern=r"checkpoint(\d+)\.pt"):
    """Files in ``path`` matching ``pattern``, sorted by the first group, descending."""
    matcher = re.compile(pattern)
    entries = []
    if not os.path.isdir(path):
        return []
    for i, name in enumerate(os.listdir(path)):
        m = matcher.fullmatch(name)
        if m is None:
            continue
        key = float(m.group(1)) if len(m.groups()) > 0 else i
        entries.append((key, name))
    return [os.path.join(path, name) for _, name in sorted(entries, reverse=True)]


def torch_persistent_save(obj, filename):
    """Atomic save: write ``<file>.tmp`` then rename; up to 3 attempts."""
    tmp = filename + ".tmp"
    for attempt in range(3):
        try:
            with open(tmp, "wb") as f:
                torch.save(obj, f)
            os.replace(tmp, filename)
            return
        except Exception:  # noqa: BLE001
            if attempt == 2:
                logger.error(traceback.format_exc())
                raise


def verify_checkpoint_directory(save_dir: str) -> None:
    os.makedirs(save_dir, exist_ok=True)
    probe = os.path.join(save_dir, "dummy")
    try:
        with open(probe, "w"):
            pass
    except OSError as e:
        logger.warning("Unable to access checkpoint save directory: {}".format(save_dir))
        raise e
    else:
        os.remove(probe)
