"""Checkpoint state: what goes into the file and how a file is turned back into a running trainer.

Schema = SURVEY.md App. B (reference ``unicore/trainer.py:224-284`` writes it, ``:299-483`` reads it):
``args, model, loss, optimizer_history[{loss_name, optimizer_name, lr_scheduler_state, num_updates}], task_state,
extra_state{metrics, previous_training_time, train_iterator, ...}, last_optimizer_state, ema{params, decay}``.
``checkpoint_utils.save_checkpoint`` first runs ``trainer.consolidate_optimizer_state()`` on EVERY rank (sharded
optimizer state and EMA slices are merged by collectives there), then the master alone assembles and writes the state;
on load the master reads and broadcasts.
"""
import logging
import os

from unicore import checkpoint_utils, utils
from unicore.distributed import utils as distributed_utils
from unicore.logging import meters, metrics

logger = logging.getLogger(__name__)


def assemble(trainer) -> dict:
    optimizer = trainer.optimizer
    # a deferred overflow decision may still take back the last update: settle it BEFORE the counters are recorded
    settle = getattr(optimizer, "resolve_pending_overflow", None)
    if settle is not None:
        settle()
    history_entry = {
        "loss_name": type(trainer.get_loss()).__name__,
        "optimizer_name": type(optimizer).__name__,
        "lr_scheduler_state": trainer.lr_scheduler.state_dict(),
        "num_updates": trainer.get_num_updates(),
    }
    state = {
        "args": trainer.args,
        "model": trainer.model.state_dict(),
        "loss": trainer.loss.state_dict() if utils.has_parameters(trainer.loss) else None,
        "optimizer_history": list(trainer._optim_history or []) + [history_entry],
        "task_state": trainer.task.state_dict() if trainer.task is not None else {},
        "extra_state": {
            "metrics": metrics.state_dict(),
            "previous_training_time": trainer.cumulative_training_time(),
        },
    }
    if not trainer.args.no_save_optimizer_state:
        state["last_optimizer_state"] = optimizer.state_dict()
    if trainer.ema is not None:
        state["ema"] = trainer.ema.state_dict()  # (sharded EMA slices were merged by consolidate_optimizer_state)
    return state


def read_and_share(trainer, filename):
    """Rank 0 reads ``filename`` (if it exists); every rank returns the same state dict or ``None``."""
    shared = trainer.data_parallel_world_size > 1
    lead = trainer.data_parallel_rank == 0
    group = trainer.data_parallel_process_group
    found = os.path.isfile(filename) if lead else None
    if shared:
        found = distributed_utils.broadcast_object(found, src_rank=0, group=group)
    if not found:
        return None
    state = checkpoint_utils.load_checkpoint_to_cpu(filename) if lead else None
    if shared:
        logger.info("Broadcast checkpoint from rank_0")
        state = distributed_utils.broadcast_object(state, src_rank=0, group=group)
    return state


def restore_weights(trainer, state, filename) -> str:
    """Load model (or EMA-as-model) and loss weights.  Returns which of the two was loaded: "model" | "ema"."""
    args = trainer.args
    try:
        if args.load_from_ema:
            logger.info("loading ema state to model")
            report = trainer.model.load_state_dict(state["ema"]["params"], strict=False, model_args=args)
            kind = "ema"
        else:
            report = trainer.model.load_state_dict(state["model"], strict=False, model_args=args)
            state.pop("model")  # free host memory early
            kind = "model"
        for label in ("missing_keys", "unexpected_keys"):
            keys = getattr(report, label, None) if report is not None else None
            if keys:
                logger.warning("Error in loading model state, {} {}".format(label, keys))
        if utils.has_parameters(trainer.get_loss()):
            trainer.get_loss().load_state_dict(state["loss"], strict=True)
            state.pop("loss")
    except Exception:
        raise Exception(
            "Cannot load model parameters from checkpoint {}; please ensure that the architectures match.".format(filename)
        )
    return kind


def restore_meters(extra_state, reset_meters: bool) -> None:
    itr_state = extra_state["train_iterator"]
    if itr_state.get("version", 1) >= 2 and itr_state["iterations_in_epoch"] == 0:
        reset_meters = True  # the checkpoint was taken at an epoch boundary
    if "metrics" in extra_state and not reset_meters:
        metrics.load_state_dict(extra_state["metrics"])
        for meter in metrics.get_meters("default").values():
            if isinstance(meter, meters.TimeMeter):
                meter.reset()  # wall-clock anchors of the old process mean nothing now


def check_compatible(last, trainer) -> None:
    for what, recorded, current in (
        ("Loss", last["loss_name"], type(trainer.get_loss()).__name__),
        ("Optimizer", last["optimizer_name"], type(trainer.optimizer).__name__),
    ):
        if recorded != current:
            raise ValueError(
                "{} does not match; please reset the optimizer (--reset-optimizer). {} vs {}".format(what, recorded, current)
            )
