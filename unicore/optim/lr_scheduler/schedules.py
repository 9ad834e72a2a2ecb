"""The nine built-in learning-rate schedules.

Formulas and flags follow the reference one-file-per-schedule implementations
(``unicore/optim/lr_scheduler/*.py``); each class cites its counterpart.  Two reference flags that
are read but never declared (``--lr-patience``, ``--phase-ratio``) are declared here.
"""
import math
from collections.abc import Collection

import torch.optim.lr_scheduler

from unicore.utils import eval_str_list

from . import UnicoreLRScheduler, register_lr_scheduler


def _single_lr(args, who):
    lr = args.lr
    if isinstance(lr, Collection):
        if len(lr) > 1:
            raise ValueError(
                "Cannot use a fixed learning rate schedule with {}. Consider --lr-scheduler=fixed instead. ({})".format(who, lr)
            )
        lr = lr[0]
    return lr


def _epoch_lr(args, epoch_index):
    lrs = args.lr
    return lrs[min(epoch_index, len(lrs) - 1)]


@register_lr_scheduler("pass_through")
class PassThroughScheduleSchedule(UnicoreLRScheduler):
    """Delegate to a scheduler attached to the optimizer (``pass_through.py:11``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        if not hasattr(optimizer, "lr_scheduler") or optimizer.lr_scheduler is None:
            raise ValueError("Pass-through schedule can only be used with optimizers with their own schedulers")

    def state_dict(self):
        return self.optimizer.lr_scheduler.state_dict()

    def load_state_dict(self, state_dict):
        self.optimizer.lr_scheduler.load_state_dict(state_dict)

    def step_begin_epoch(self, epoch):
        return self.optimizer.lr_scheduler.step_begin_epoch(epoch)

    def step_update(self, num_updates):
        return self.optimizer.lr_scheduler.step_update(num_updates)


@register_lr_scheduler("fixed")
class FixedLRSchedule(UnicoreLRScheduler):
    """Per-epoch LR list, optional forced annealing and linear warm-up (``fixed_schedule.py:13``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        self.lr = args.lr[0]
        self.warmup_factor = 1.0 / args.warmup_updates if args.warmup_updates > 0 else 1

    @staticmethod
    def add_args(parser):
        parser.add_argument("--force-anneal", "--fa", type=int, metavar="N", help="force annealing at specified epoch")
        parser.add_argument("--lr-shrink", default=0.1, type=float, metavar="LS",
                            help="shrink factor for annealing, lr_new = (lr * lr_shrink)")
        parser.add_argument("--warmup-updates", default=0, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")

    def state_dict(self):
        return {"lr": self.lr}

    def load_state_dict(self, state_dict):
        if "lr" in state_dict:
            self.lr = state_dict["lr"]

    def get_next_lr(self, epoch):
        anneal_at = self.args.force_anneal
        if anneal_at is None or epoch < anneal_at:
            return _epoch_lr(self.args, epoch - 1)
        return self.args.lr[-1] * self.args.lr_shrink ** (epoch + 1 - anneal_at)

    def step_begin_epoch(self, epoch):
        self.lr = self.get_next_lr(epoch)
        self.optimizer.set_lr(self.warmup_factor * self.lr)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        warm = self.args.warmup_updates
        if warm > 0 and num_updates < warm:
            self.warmup_factor = (num_updates + 1) / float(warm)
            self.optimizer.set_lr(self.warmup_factor * self.lr)
        else:
            self.optimizer.set_lr(self.lr)
        return self.optimizer.get_lr()


@register_lr_scheduler("inverse_sqrt")
class InverseSquareRootSchedule(UnicoreLRScheduler):
    """Linear warm-up then ``lr * sqrt(warmup/t)`` (``inverse_square_root_schedule.py:14``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        peak = _single_lr(args, "inverse_sqrt")
        if args.warmup_init_lr < 0:
            args.warmup_init_lr = 0 if args.warmup_updates > 0 else peak
        self.lr_step = (peak - args.warmup_init_lr) / args.warmup_updates
        self.decay_factor = peak * args.warmup_updates ** 0.5
        self.lr = args.warmup_init_lr
        self.optimizer.set_lr(self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--warmup-updates", default=4000, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--warmup-init-lr", default=-1, type=float, metavar="LR",
                            help="initial learning rate during warmup phase; default is args.lr")

    def step(self, epoch, val_loss=None):
        super().step(epoch, val_loss)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        if num_updates < self.args.warmup_updates:
            self.lr = self.args.warmup_init_lr + num_updates * self.lr_step
        else:
            self.lr = self.decay_factor * num_updates ** -0.5
        self.optimizer.set_lr(self.lr)
        return self.lr


@register_lr_scheduler("polynomial_decay")
class PolynomialDecayLRSchedule(UnicoreLRScheduler):
    """Linear warm-up then polynomial decay to ``--end-learning-rate``
    (``polynomial_decay_schedule.py:12``; ``--warmup-ratio`` uses the trainer-provided total)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        if self.args.warmup_ratio > 0:
            if total_train_steps is None:
                raise ValueError("--warmup-ratio needs the total number of training steps")
            self.warmup_updates = int(self.args.warmup_ratio * total_train_steps)
            self.total_num_update = total_train_steps
        else:
            if args.total_num_update <= 0:
                raise ValueError("--total-num-update must be positive")
            self.warmup_updates = args.warmup_updates
            self.total_num_update = args.total_num_update
        self.lr = args.lr[0]
        self.warmup_factor = 1.0 / self.warmup_updates if self.warmup_updates > 0 else 1
        self.end_learning_rate = args.end_learning_rate
        self.power = args.power
        self.optimizer.set_lr(self.warmup_factor * self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--force-anneal", "--fa", type=int, metavar="N", help="force annealing at specified epoch")
        parser.add_argument("--warmup-updates", default=0, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--warmup-ratio", default=-1.0, type=float, metavar="N",
                            help="warmup the learning rate linearly for the first N-percent updates")
        parser.add_argument("--end-learning-rate", default=0.0, type=float)
        parser.add_argument("--power", default=1.0, type=float)
        parser.add_argument("--total-num-update", default=1000000, type=int)

    def get_next_lr(self, epoch):
        if self.args.force_anneal is None or epoch < self.args.force_anneal:
            return _epoch_lr(self.args, epoch)
        return self.optimizer.get_lr()

    def step_begin_epoch(self, epoch):
        self.lr = self.get_next_lr(epoch)
        self.optimizer.set_lr(self.warmup_factor * self.lr)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        warm = self.warmup_updates
        if warm > 0 and num_updates <= warm:
            self.warmup_factor = num_updates / float(warm)
            lr = self.warmup_factor * self.lr
        elif num_updates >= self.total_num_update:
            lr = self.end_learning_rate
        else:
            remaining = 1 - (num_updates - warm) / (self.total_num_update - warm)
            lr = (self.lr - self.end_learning_rate) * remaining ** self.power + self.end_learning_rate
        self.optimizer.set_lr(lr)
        return self.optimizer.get_lr()


@register_lr_scheduler("exponential_decay")
class ExponentialDecayLRSchedule(UnicoreLRScheduler):
    """Linear warm-up then ``lr * ratio^(t/decay_steps)`` (smooth or stair)
    (``exponential_decay_schedule.py:12``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        self.warmup_updates = args.warmup_updates
        self.lr = args.lr[0]
        self.warmup_factor = 1.0 / self.warmup_updates if self.warmup_updates > 0 else 1.0
        self.decay_ratio = args.decay_ratio
        self.decay_steps = args.decay_steps
        self.stair_decay = getattr(args, "stair_decay", False)
        self.optimizer.set_lr(self.warmup_factor * self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--warmup-updates", default=1000, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--decay-ratio", default=0.95, type=float)
        parser.add_argument("--decay-steps", default=500, type=int)
        parser.add_argument("--stair-decay", action="store_true")

    def step_update(self, num_updates):
        if self.warmup_updates > 0 and num_updates <= self.warmup_updates:
            self.warmup_factor = num_updates / float(self.warmup_updates)
            lr = self.warmup_factor * self.lr
        elif self.stair_decay:
            lr = self.lr * float(self.decay_ratio ** int(num_updates // self.decay_steps))
        else:
            lr = self.lr * float(self.decay_ratio ** ((num_updates - self.warmup_updates) / self.decay_steps))
        self.optimizer.set_lr(lr)
        return self.optimizer.get_lr()


@register_lr_scheduler("cosine")
class CosineLRSchedule(UnicoreLRScheduler):
    """Warm-up + cyclic cosine annealing with period growth ``--t-mult`` and per-cycle shrink
    (``cosine_lr_scheduler.py:15``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        self.max_lr = _single_lr(args, "cosine")
        if args.min_lr is None:
            args.min_lr = 0.0
        if not self.max_lr > args.min_lr:
            raise ValueError("max_lr (={}) must be more than min_lr (={})".format(args.lr, args.min_lr))
        if total_train_steps is None:
            raise ValueError("cosine schedule needs the total number of training steps")
        self.warmup_updates = (
            int(args.warmup_ratio * total_train_steps) if args.warmup_ratio > 0 else args.warmup_updates
        )
        if args.warmup_init_lr < 0:
            args.warmup_init_lr = args.min_lr
        self.t_mult = args.t_mult
        self.period = args.lr_period_updates
        if self.period <= 0:
            self.period = total_train_steps - self.warmup_updates
        self.lr_step = (self.max_lr - args.warmup_init_lr) / self.warmup_updates if self.warmup_updates > 0 else 1
        self.lr_shrink = args.lr_shrink
        self.lr = args.warmup_init_lr
        self.optimizer.set_lr(self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--warmup-updates", default=0, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--warmup-ratio", default=-1.0, type=float, metavar="N",
                            help="warmup the learning rate linearly for the first N-percent updates")
        parser.add_argument("--warmup-init-lr", default=-1, type=float, metavar="LR",
                            help="initial learning rate during warmup phase; default is args.lr")
        parser.add_argument("--min-lr", type=float, metavar="LR", help="min learning rate")
        parser.add_argument("--max-lr", type=float, metavar="LR", help="max learning rate, must be more than args.lr")
        parser.add_argument("--t-mult", default=1, type=float, metavar="LR", help="factor to grow the length of each period")
        parser.add_argument("--lr-period-updates", default=-1, type=float, metavar="LR",
                            help="initial number of updates per period")
        parser.add_argument("--lr-shrink", default=0.1, type=float, metavar="LS", help="shrink factor for annealing")

    def step(self, epoch, val_loss=None):
        super().step(epoch, val_loss)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        if num_updates < self.warmup_updates:
            self.lr = self.args.warmup_init_lr + num_updates * self.lr_step
        else:
            t = num_updates - self.warmup_updates
            if self.t_mult != 1:
                # geometric series of period lengths: find the cycle index i containing t
                cycle = math.floor(math.log(1 - t / self.period * (1 - self.t_mult), self.t_mult))
                span = self.t_mult ** cycle * self.period
                start = (1 - self.t_mult ** cycle) / (1 - self.t_mult) * self.period
                frac = float(t - start) / span
            else:
                cycle = 0
                frac = min(1.0, float(t) / self.period)
            shrink = self.lr_shrink ** cycle
            lo, hi = self.args.min_lr * shrink, self.max_lr * shrink
            self.lr = lo + 0.5 * (hi - lo) * (1 + math.cos(math.pi * frac))
        self.optimizer.set_lr(self.lr)
        return self.lr


@register_lr_scheduler("reduce_lr_on_plateau")
class ReduceLROnPlateauLRSchedule(UnicoreLRScheduler):
    """Shrink the LR when the validation metric stalls; optional warm-up
    (``reduce_lr_on_plateau.py:16``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        if len(args.lr) > 1:
            raise ValueError(
                "Cannot use a fixed learning rate schedule with reduce_lr_on_plateau. Consider --lr-scheduler=fixed instead."
            )
        self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.optimizer.optimizer,
            patience=getattr(args, "lr_patience", 0),
            factor=args.lr_shrink,
            mode="max" if args.maximize_best_checkpoint_metric else "min",
            threshold=args.lr_threshold,
        )
        peak = args.lr[0]
        if args.warmup_init_lr < 0:
            args.warmup_init_lr = 0 if args.warmup_updates > 0 else peak
        if args.warmup_updates > 0:
            self.lr_step = (peak - args.warmup_init_lr) / args.warmup_updates
        self.warmup_end = args.warmup_updates <= 0
        self.lr = args.warmup_init_lr
        self.optimizer.set_lr(self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--lr-shrink", default=0.1, type=float, metavar="LS",
                            help="shrink factor for annealing, lr_new = (lr * lr_shrink)")
        parser.add_argument("--lr-threshold", default=1e-4, type=float, metavar="LT",
                            help="Threshold for measuring the new optimum, to only focus on significant changes")
        parser.add_argument("--lr-patience", default=0, type=int,
                            help="number of epochs without improvement before shrinking the lr")
        parser.add_argument("--warmup-updates", default=0, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--warmup-init-lr", default=-1, type=float, metavar="LR",
                            help="initial learning rate during warmup phase; default is args.lr")

    def state_dict(self):
        return {"best": self.lr_scheduler.best, "last_epoch": self.lr_scheduler.last_epoch}

    def load_state_dict(self, state_dict):
        self.lr_scheduler.best = state_dict["best"]
        if "last_epoch" in state_dict:
            self.lr_scheduler.last_epoch = state_dict["last_epoch"]

    def step(self, epoch, val_loss=None):
        if val_loss is not None and self.warmup_end:
            self.lr_scheduler.step(val_loss)
        else:
            self.lr_scheduler.last_epoch = epoch
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        if self.args.warmup_updates > 0:
            if num_updates <= self.args.warmup_updates:
                self.lr = self.args.warmup_init_lr + num_updates * self.lr_step
                self.optimizer.set_lr(self.lr)
            else:
                self.warmup_end = True
        return self.optimizer.get_lr()


@register_lr_scheduler("tri_stage")
class TriStageLRSchedule(UnicoreLRScheduler):
    """Warm-up -> hold -> exponential decay -> constant (``tri_stage_lr_scheduler.py:14``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        peak = _single_lr(args, "tri-stage lr")
        self.peak_lr = peak
        self.init_lr = args.init_lr_scale * peak
        self.final_lr = args.final_lr_scale * peak
        ratio = getattr(args, "phase_ratio", None)
        if ratio is not None:
            if args.max_update <= 0:
                raise ValueError("--phase-ratio needs --max-update")
            if abs(sum(ratio) - 1) > 1e-9:
                raise ValueError("phase ratios must add up to 1")
            self.warmup_steps, self.hold_steps, self.decay_steps = (int(args.max_update * r) for r in ratio)
        else:
            self.warmup_steps, self.hold_steps, self.decay_steps = args.warmup_steps, args.hold_steps, args.decay_steps
        if self.warmup_steps + self.hold_steps + self.decay_steps <= 0:
            raise ValueError("please specify steps or phase_ratio")
        self.warmup_rate = (self.peak_lr - self.init_lr) / self.warmup_steps if self.warmup_steps != 0 else 0
        self.decay_factor = -math.log(args.final_lr_scale) / self.decay_steps
        self.lr = self.init_lr
        self.optimizer.set_lr(self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--warmup-steps", default=4000, type=int, metavar="N",
                            help="warmup the learning rate linearly for the first N updates")
        parser.add_argument("--hold-steps", default=20000, type=int, metavar="N", help="steps in hold stage")
        parser.add_argument("--decay-steps", default=60000, type=int, metavar="N", help="steps in decay stages")
        parser.add_argument("--init-lr-scale", default=0.01, type=float,
                            help="initial learning rate scale during warmup phase; default is 0.01")
        parser.add_argument("--final-lr-scale", default=0.01, type=float, help="final learning rate scale; default to 0.01")
        parser.add_argument("--phase-ratio", default=None, type=lambda s: eval_str_list(s, float),
                            help="(warmup, hold, decay) fractions of --max-update; overrides the *-steps flags")

    def _decide_stage(self, t):
        bounds = (self.warmup_steps, self.hold_steps, self.decay_steps)
        if t < bounds[0]:
            return 0, t
        t -= bounds[0]
        if t < bounds[1]:
            return 1, t
        t -= bounds[1]
        if t <= bounds[2]:
            return 2, t
        return 3, t - bounds[2]

    def step(self, epoch, val_loss=None):
        super().step(epoch, val_loss)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        stage, k = self._decide_stage(num_updates)
        if stage == 0:
            self.lr = self.init_lr + self.warmup_rate * k
        elif stage == 1:
            self.lr = self.peak_lr
        elif stage == 2:
            self.lr = self.peak_lr * math.exp(-self.decay_factor * k)
        else:
            self.lr = self.final_lr
        self.optimizer.set_lr(self.lr)
        return self.lr


@register_lr_scheduler("triangular")
class TriangularLRSchedule(UnicoreLRScheduler):
    """Cyclical triangular LR (https://arxiv.org/abs/1506.01186) (``triangular_lr_scheduler.py:14``)."""

    def __init__(self, args, optimizer, total_train_steps):
        super().__init__(args, optimizer, total_train_steps)
        lr = _single_lr(args, "triangular")
        if not args.max_lr > lr:
            raise ValueError("max_lr must be more than lr")
        self.min_lr = lr
        self.max_lr = args.max_lr
        self.stepsize = args.lr_period_updates // 2
        self.lr_shrink = args.lr_shrink
        self.shrink_min = args.shrink_min
        self.lr = self.min_lr
        self.optimizer.set_lr(self.lr)

    @staticmethod
    def add_args(parser):
        parser.add_argument("--max-lr", required=True, type=float, metavar="LR",
                            help="max learning rate, must be more than args.lr")
        parser.add_argument("--lr-period-updates", default=5000, type=float, metavar="LR",
                            help="initial number of updates per period (cycle length)")
        parser.add_argument("--lr-shrink", default=0.1, type=float, metavar="LS", help="shrink factor for annealing")
        parser.add_argument("--shrink-min", action="store_true", help="if set, also shrinks min lr")

    def step(self, epoch, val_loss=None):
        super().step(epoch, val_loss)
        return self.optimizer.get_lr()

    def step_update(self, num_updates):
        cycle = math.floor(num_updates / (2 * self.stepsize))
        shrink = self.lr_shrink ** cycle
        hi = self.max_lr * shrink
        lo = self.min_lr * shrink if self.shrink_min else self.min_lr
        x = abs(num_updates / self.stepsize - 2 * (cycle + 1) + 1)
        self.lr = lo + (hi - lo) * max(0, 1 - x)
        self.optimizer.set_lr(self.lr)
        return self.lr
