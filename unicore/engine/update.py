"""One optimizer update, written as explicit phases over device-resident step state.

    enter       train mode, gradients cleared, wall-clock meter started
    accumulate  forward + backward over the micro-batches (gradient reduction overlaps the last backward)
    exchange    this rank's statistics go into the ledger; depending on the engine they are summed now or handed to the
                fused optimizer tail, which sums them while it updates the parameters
    apply       finish the reduction, normalise (world / sample size, 1 / loss scale), clip, optimizer (+ EMA)
    report      step counter, meters, lazily materialised logging output

Numerics contract (reference ``unicore/trainer.py:571-803``): gradients end up as
``sum over ranks and micro-batches / sum of sample sizes``; a non-finite norm skips the update and shrinks the loss scale
(fp16) or aborts under the NaN detector (bf16 / fp32); dropout streams are seeded per (update, micro-batch, rank), the
optimizer's stream per update only (stochastic rounding must agree on all ranks).
"""
import contextlib
import logging

import torch

from unicore import utils
from unicore.logging import metrics
from unicore.nan_detector import NanDetector

from .ledger import StatLedger, gather_objects, reduce_ledger

logger = logging.getLogger(__name__)


def _is_oom(exc: BaseException) -> bool:
    return "out of memory" in str(exc)


class UpdateStep:
    def __init__(self, trainer, samples, raise_oom: bool = False):
        self.trainer = trainer
        self.samples = samples
        self.raise_oom = raise_oom
        self.logs = []
        self.log_keys = []
        self.sample_size = 0
        self.ooms = 0
        self.last_was_dummy = False
        self.ledger = None
        self.stats_in_tail = False

    # ------------------------------------------------------------------------------------------------------------
    def run(self):
        self.enter()
        if not self.accumulate():
            return None
        self.exchange()
        overflow, grad_norm = self.apply()
        return self.report(overflow, grad_norm)

    # ------------------------------------------------------------------------------------------------------------
    def enter(self):
        t = self.trainer
        # Module.train() walks the whole module tree (~1.3 ms for BERT-base): only flip when needed
        if not t.model.training:
            t.model.train()
        if not t.loss.training:
            t.loss.train()
        t.zero_grad()
        metrics.log_start_time("train_wall", priority=800, round=2)

    # ------------------------------------------------------------------------------------------------------------
    def _sync_scope(self, index):
        """Gradient accumulation: only the last micro-batch's backward may trigger communication."""
        t = self.trainer
        if t.data_parallel_world_size > 1 and index < len(self.samples) - 1 and hasattr(t.model, "no_sync"):
            return t.model.no_sync()
        return contextlib.nullcontext()

    def accumulate(self) -> bool:
        t = self.trainer
        args, update = t.args, t.get_num_updates()
        for index, raw in enumerate(self.samples):
            sample, self.last_was_dummy = t._prepare_sample(raw)
            try:
                with self._sync_scope(index):
                    with utils.torch_seed(args.seed, update, index, t.data_parallel_rank):
                        loss, size, log = t.task.train_step(
                            sample=sample, model=t.model, loss=t.loss, optimizer=t.optimizer, update_num=update,
                            ignore_grad=self.last_was_dummy,
                        )
                        del loss
                    if args.per_sample_clip_norm > 0:
                        t.optimizer.per_sample_clip_grad_norm(args.per_sample_clip_norm)
                self.logs.append(log)
                self.sample_size = self.sample_size + size
                if t.cuda and update == 0:
                    torch.cuda.empty_cache()  # the first step sets the allocator's high-water mark
            except RuntimeError as exc:
                if not _is_oom(exc):
                    raise
                t._log_oom(exc)
                if self.raise_oom:
                    raise
                logger.warning("attempting to recover from OOM in forward/backward pass")
                self.ooms += 1
                t.zero_grad()
                if t.cuda:
                    torch.cuda.empty_cache()
                if t.data_parallel_world_size == 1:
                    return False
                # distributed: the other ranks are inside this update's collectives - stay in step with them (this
                # rank contributes zero gradients and no sample size) and let the summed `ooms` tell the story
                flush = getattr(t.dp_engine, "flush_after_failure", None)
                if flush is not None:
                    flush()
        if self.last_was_dummy:
            self.sample_size = self.sample_size * 0.0
        self.sample_size = self.sample_size.float() if torch.is_tensor(self.sample_size) else float(self.sample_size)
        return True

    # ------------------------------------------------------------------------------------------------------------
    def exchange(self):
        t = self.trainer
        if t.data_parallel_world_size == 1:
            return
        train_time = t.clock.local()
        summable = t.task.__class__.logging_outputs_can_be_summed(t.get_loss(), is_train=True)
        if not summable:
            self.logs, (self.sample_size, self.ooms, total_time) = gather_objects(
                self.logs, [self.sample_size, self.ooms, train_time], t.data_parallel_process_group,
                max_size=getattr(t.args, "all_gather_list_size", 16384), ignore=self.last_was_dummy,
            )
            t.clock.set_fleet_mean(total_time / t.data_parallel_world_size)
            return
        ledger = StatLedger()
        self.log_keys = ledger.add_logging_outputs(self.logs, ignore=self.last_was_dummy)
        ledger.add("sample_size", self.sample_size)
        ledger.add("ooms", self.ooms)
        ledger.add("train_time", train_time)
        self.ledger = ledger
        opt = t.optimizer
        if getattr(opt, "uses_fused_tail", False) and len(ledger) <= 64:
            # summed INSIDE the optimizer tail; the kernel also divides the gradients by the summed sample size
            packed = ledger.pack(t.device)
            opt.set_step_stats(packed, denom_index=ledger.position("sample_size"))
            self.stats_in_tail = True
            return
        reduce_ledger(ledger, t.device, t.data_parallel_process_group, engine=t.dp_engine)
        self._adopt_sums()

    def _adopt_sums(self):
        t, ledger = self.trainer, self.ledger
        self.logs = ledger.logging_output(self.log_keys)
        self.sample_size = ledger.value("sample_size").float()
        self.ooms = ledger.value("ooms")
        # kept as the device scalar the reduction produced: converting it here would drain the launch queue right
        # after backward on every multi-GPU step; the clock converts lazily
        t.clock.set_fleet_mean(ledger.value("train_time") / t.data_parallel_world_size)

    # ------------------------------------------------------------------------------------------------------------
    def apply(self):
        t = self.trainer
        args, opt, update = t.args, t.optimizer, t.get_num_updates()
        overflow, grad_norm = False, None
        try:
            with torch.autograd.profiler.record_function("reduce-grads"):
                opt.all_reduce_grads(t.model)
                if utils.has_parameters(t.loss) and t.loss is not t.get_loss():
                    opt.all_reduce_grads(t.loss)
            with torch.autograd.profiler.record_function("multiply-grads"):
                # data-parallel engines average over ranks; the contract is sum(grads) / sum(sample sizes)
                numer = t.data_parallel_world_size if t.data_parallel_world_size > 1 else 1
                if self.stats_in_tail:
                    opt.multiply_grads(float(numer))
                else:
                    denom = self.sample_size
                    if not torch.is_tensor(denom) and not denom > 0:
                        denom = 1.0  # (tensors are assumed non-zero: no host sync; a zero python count means zero grads)
                    opt.multiply_grads(numer / denom)
            with torch.autograd.profiler.record_function("clip-grads"):
                grad_norm = t.clip_grad_norm(args.clip_norm)
            if not getattr(opt, "uses_fused_tail", False):
                t._check_grad_norms(grad_norm)
            with torch.autograd.profiler.record_function("optimizer"):
                with utils.torch_seed(args.seed, update):  # rank-invariant stream (stochastic rounding)
                    t.task.optimizer_step(opt, model=t.model, update_num=update)
            if getattr(opt, "uses_fused_tail", False):
                grad_norm = opt.step_grad_norm()  # (the kernel's state slot is reused by the next update)
            if self.stats_in_tail:
                self.ledger.adopt(opt.step_stats())
                self._adopt_sums()
            if t.ema is not None and not t.ema_in_optimizer:
                with torch.autograd.profiler.record_function("ema"):
                    if args.fp16 or args.bf16:
                        t.ema.update(opt.fp32_params)
                    else:
                        t.ema.update(t.model.named_parameters())
        except FloatingPointError:
            # non-finite or inconsistent gradient norm: replay under the NaN detector for a useful message
            t.zero_grad()
            with NanDetector(t.get_model()):
                for index, raw in enumerate(self.samples):
                    sample, _ = t._prepare_sample(raw)
                    with utils.torch_seed(args.seed, update, index, t.data_parallel_rank):
                        t.task.train_step(sample, t.model, t.loss, opt, update, ignore_grad=False)
            raise
        except OverflowError as exc:
            overflow = True
            logger.info("NOTE: gradient overflow detected, ignoring gradient, {}".format(exc))
            grad_norm = torch.zeros((), device=t.device)
            if self.stats_in_tail and self.ledger.sums is None:
                self.ledger.adopt(opt.step_stats())
                self._adopt_sums()
            t.zero_grad()
        except RuntimeError as exc:
            if _is_oom(exc):
                t._log_oom(exc)
                logger.error("OOM during optimization, irrecoverable")
            raise
        return overflow, grad_norm

    # ------------------------------------------------------------------------------------------------------------
    def report(self, overflow, grad_norm):
        t = self.trainer
        args = t.args
        out = None
        if not overflow:
            t.set_num_updates(t.get_num_updates() + 1)
            if t.cuda and t.cuda_env is not None:
                used = torch.cuda.max_memory_allocated() / 1024 ** 3
                torch.cuda.reset_peak_memory_stats()
                metrics.log_scalar("gb_free", t.cuda_env.total_memory_in_GB - used, priority=1500, round=1, weight=0)
            out = t._reduce_and_log_stats(self.logs, self.sample_size, grad_norm)
            every = args.empty_cache_freq
            if t.cuda and every > 0 and (t.get_num_updates() + every - 1) % every == 0:
                torch.cuda.empty_cache()
        if args.fp16:
            metrics.log_scalar("loss_scale", t.optimizer.scaler.loss_scale, priority=700, round=4, weight=0)
        metrics.log_stop_time("train_wall")
        health = getattr(t.dp_engine, "check_health", None)
        if health is not None:
            health()  # a host-memory read: did a collective kernel report a dead or out-of-step peer?
        return out


class LazyStats(object):
    """The logging output of one ``train_step``: a read-only mapping that is materialised on first access.

    Producing the smoothed values means bringing device-resident meters to the host, i.e. waiting for the step to
    finish on the GPU.  Callers that only test ``is not None`` (the CLI between log intervals, the device-timed
    benchmark) never pay for that; callers that read a value get exactly what an eager version would have returned.
    """

    _DROPPED = ("ppl", "wps", "wpb", "bsz")

    def __init__(self, agg, sample_size):
        self._agg, self._sample_size, self._values = agg, sample_size, None

    def _materialise(self):
        if self._values is None:
            values = self._agg.get_smoothed_values()
            values["sample_size"] = self._sample_size
            for key in self._DROPPED:
                values.pop(key, None)
            self._values, self._agg = values, None
        return self._values

    def _single(self, key):
        """One value without materialising the rest: a meter that already holds host numbers (the loss statistics of a
        single-process run, staged behind the forward pass) answers at once; device-resident ones (the gradient norm)
        are only waited for when somebody asks for THEM."""
        if self._values is not None:
            return self._values[key]
        if key == "sample_size":
            return self._sample_size
        if key in self._DROPPED or key.startswith("_") or key not in self._agg:
            raise KeyError(key)
        return self._agg.get_smoothed_value(key)

    def __getitem__(self, key):
        return self._single(key)

    def __contains__(self, key):
        if self._values is not None:
            return key in self._values
        return key == "sample_size" or (key in self._agg and key not in self._DROPPED and not key.startswith("_"))

    def __iter__(self):
        return iter(self._materialise())

    def __len__(self):
        return len(self._materialise())

    def __repr__(self):
        return repr(self._materialise())

    def get(self, key, default=None):
        try:
            return self._single(key)
        except KeyError:
            return default

    def keys(self):
        return self._materialise().keys()

    def values(self):
        return self._materialise().values()

    def items(self):
        return self._materialise().items()
