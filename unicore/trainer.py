"""``Trainer``: the public face of the data-parallel training engine (``unicore/engine``).

The names and the checkpoint schema are the reference's (``unicore/trainer.py:30`` there: ``train_step:571``,
``valid_step:805``, ``save_checkpoint:286``, ``load_checkpoint:299``, ``get_train_iterator:484``, ``lr_step*:853-874``),
because ``unicore-train``, user tasks and existing checkpoints rely on them.  What happens behind them is this
framework's own design:

* one update is a sequence of explicit phases (``engine.update.UpdateStep``) over device-resident step state - the loss
  scale outcome, the squared gradient norm, the overflow flag, the summed sample size never visit the host on the
  step path (``--deferred-overflow-check``; the logging output is a lazily materialised mapping);
* with ``--ddp-backend b200`` everything between the end of backward and the next forward is ONE kernel
  (``unicore_b200/parallel/fused_tail.py``): the last gradient bucket's reduce-scatter, the norm and statistics
  exchange over NVLink peer memory, unscale / normalise / clip / overflow decision, Adam and EMA on a 1/N shard of
  compact fp32 state, and the parameter all-gather by the kernel's own multimem stores.  No NCCL call, no separate
  optimizer, norm, statistics or EMA launch on that path;
* other engines (c10d, legacy_ddp, CPU / gloo) run the same phases with their collectives: statistics as one packed
  fp64 vector, the gradient-norm consistency vector on the same channel;
* device-agnostic: CPU / gloo runs work (the reference hard-codes CUDA in several places).
"""
import logging
import sys
from itertools import chain

import torch

from unicore import optim, utils
from unicore.distributed import utils as distributed_utils
from unicore.ema import ExponentialMovingAverageModel
from unicore.engine import LazyStats, Replica, StatLedger, TrainingClock, UpdateStep, snapshot
from unicore.engine.ledger import gather_objects, reduce_ledger
from unicore.logging import metrics
from unicore.optim import lr_scheduler

logger = logging.getLogger(__name__)

_LazyStats = LazyStats  # (historic private name)


class Trainer(object):
    """Main class for data parallel training (gradients are reduced across replicas every update)."""

    def __init__(self, args, task, model, loss):
        self.args, self.task = args, task
        self.cuda = torch.cuda.is_available() and not getattr(args, "cpu", False)
        self.device = torch.device("cuda") if self.cuda else torch.device("cpu")
        if args.validate_with_ema and not args.ema_decay > 0:
            raise ValueError("--validate-with-ema requires --ema-decay > 0")

        self.replica = Replica(args, model, loss, self.device, distributed=self.data_parallel_world_size > 1)
        self.clock = TrainingClock()
        self._progress = 0               # optimizer updates applied so far
        self._optimizer = self._lr_scheduler = None
        self._optim_history = None
        self._total_train_steps = None
        self._dummy_batch = None
        self._warned = set()
        self.ema_in_optimizer = False    # True when the fused tail carries the EMA update

        self.cuda_env, self.cuda_env_arr = None, None
        if self.cuda:
            self.cuda_env = utils.CudaEnvironment()
            self.cuda_env_arr = [self.cuda_env]
            if self.data_parallel_world_size > 1:
                self.cuda_env_arr = distributed_utils.all_gather_list(self.cuda_env, group=distributed_utils.get_global_group())
            if self.data_parallel_rank == 0:
                utils.CudaEnvironment.pretty_print_cuda_env_list(self.cuda_env_arr)

        live = self.model  # builds the data-parallel wrapper
        self.ema = None
        if args.ema_decay > 0 and (self.data_parallel_rank == 0 or args.validate_with_ema or self._tail_capable()):
            self.ema = ExponentialMovingAverageModel(args, live, args.ema_decay, is_flattened=(args.fp16 or args.bf16))
        metrics.log_start_time("wall", priority=790, round=2)

    # -- topology -----------------------------------------------------------------------------------------------
    @property
    def data_parallel_world_size(self):
        return 1 if self.args.distributed_world_size == 1 else distributed_utils.get_data_parallel_world_size()

    @property
    def data_parallel_rank(self):
        return 0 if self.args.distributed_world_size == 1 else distributed_utils.get_data_parallel_rank()

    @property
    def data_parallel_process_group(self):
        return distributed_utils.get_data_parallel_group()

    @property
    def is_data_parallel_master(self):
        return self.data_parallel_rank == 0

    @property
    def use_distributed_wrapper(self) -> bool:
        return self.data_parallel_world_size > 1

    @property
    def should_save_checkpoint_on_current_rank(self) -> bool:
        return self.is_data_parallel_master

    @property
    def checkpoint_suffix(self) -> str:
        return self.args.checkpoint_suffix or ""

    # -- components ---------------------------------------------------------------------------------------------------
    @property
    def model(self):
        return self.replica.get_model(self.data_parallel_process_group if self.use_distributed_wrapper else None)

    @property
    def loss(self):
        self.model  # noqa: B018  (the loss wrapper decision depends on the model's engine)
        return self.replica.get_loss(self.data_parallel_process_group if self.use_distributed_wrapper else None)

    # (``utils.validate_with_ema`` swaps the wrapped model in and out through these two names)
    @property
    def _wrapped_model(self):
        return self.replica.wrapped_model

    @_wrapped_model.setter
    def _wrapped_model(self, module):
        self.replica.wrapped_model = module

    @property
    def _model(self):
        return self.replica.model

    @property
    def _loss(self):
        return self.replica.loss

    @property
    def dp_engine(self):
        return self.replica.engine

    @property
    def optimizer(self):
        if self._optimizer is None:
            self._build_optimizer()
        return self._optimizer

    @property
    def lr_scheduler(self):
        if self._lr_scheduler is None:
            self._build_optimizer()
        return self._lr_scheduler

    def get_model(self):
        """The bare (unwrapped) model."""
        return self.replica.model

    def get_loss(self):
        return self.replica.loss

    def reinitialize(self):
        """Drop optimizer and wrappers so that they are rebuilt (e.g. after swapping the model)."""
        self._optimizer = self._lr_scheduler = None
        self.replica.drop_wrappers()

    def _tail_capable(self) -> bool:
        """Will this run use the fused optimizer tail (which needs the EMA shard on every rank)?"""
        return (self.data_parallel_world_size > 1 and getattr(self.dp_engine, "want_fused_tail", False)
                and (self.args.fp16 or self.args.bf16))

    def _trainable(self):
        named = chain(self.model.named_parameters(), self.loss.named_parameters())
        return [(n, p) for n, p in named if p.requires_grad]

    def _build_optimizer(self):
        args, engine = self.args, self.dp_engine
        legacy_only = args.ddp_backend in ("no_c10d", "legacy_ddp")
        if args.per_sample_clip_norm > 0 and not legacy_only:
            raise ValueError("--per-sample-clip-norm only supports --ddp-backend no_c10d")
        named = self._trainable()
        if args.fp16 or args.bf16:
            if args.allreduce_fp32_grad and not legacy_only:
                raise ValueError("--allreduce-fp32-grad requires --ddp-backend no_c10d")
            if hasattr(engine, "begin_optimizer_build"):
                engine.begin_optimizer_build()  # arenas of a previous build are reused, not leaked
            self._optimizer = optim.FP16Optimizer.build_optimizer(
                args, named, grad_alloc=getattr(engine, "alloc_grad_buffer", None),
                param_alloc=getattr(engine, "alloc_param_buffer", None),
            )
            if hasattr(engine, "attach_optimizer"):
                engine.attach_optimizer(self._optimizer, params=[p for _, p in named])
            self._configure_grad_sinks([p for _, p in named])
        else:
            self._optimizer = optim.build_optimizer(args, named)
        if hasattr(self._optimizer, "add_late_overflow_handler"):
            self._optimizer.add_late_overflow_handler(self._on_late_overflow)
        self._connect_ema()
        self._lr_scheduler = lr_scheduler.build_lr_scheduler(args, self._optimizer, self._total_train_steps)
        self._lr_scheduler.step_update(0)

    def _configure_grad_sinks(self, params):
        """Backward kernels may accumulate parameter gradients straight into the flat gradient arena
        (``unicore_b200/ops/grad_sink.py``: no temporaries, no AccumulateGrad add per parameter).  That bypasses autograd's
        gradient hooks, so it is enabled for a single process and for engines that take their "gradient ready" signal
        from the sink (``--ddp-backend b200`` registers itself in ``attach_optimizer``); torch DDP and the legacy engine
        keep plain autograd accumulation."""
        try:
            from unicore_b200.ops import grad_sink
        except ImportError:
            return
        engine = self.dp_engine
        if self.data_parallel_world_size == 1:
            grad_sink.enable(params) if self.cuda and not getattr(self.args, "no_grad_sinks", False) else grad_sink.disable(params)
        elif not hasattr(engine, "attach_optimizer") or getattr(self.args, "no_grad_sinks", False):
            grad_sink.disable(params)

    def _connect_ema(self):
        """Let the optimizer kernel carry the EMA update when it can (fused tail: on the shard; replicated fused Adam:
        in the same pass over the master weights)."""
        self.ema_in_optimizer = False
        if self.ema is None or not getattr(self.ema, "is_flattened", False):
            return
        opt = self._optimizer
        if getattr(opt, "uses_fused_tail", False):
            self.ema_in_optimizer = bool(opt.attach_ema(self.ema.flatten_params, self.ema.decay))
        elif getattr(opt, "is_fused", False) and hasattr(opt, "attach_replicated_ema"):
            self.ema_in_optimizer = bool(opt.attach_replicated_ema(self.ema.flatten_params, self.ema.decay))

    def sync_ema_shards(self):
        """COLLECTIVE (fused tail only): every rank keeps just its slices of the EMA arena current; before the EMA is
        saved or validated with, the slices are merged on all ranks."""
        opt = self._optimizer
        if self.ema is not None and self.ema_in_optimizer and getattr(opt, "uses_fused_tail", False):
            opt.gather_ema()

    # -- checkpoints ---------------------------------------------------------------------------------------------------
    def state_dict(self):
        return snapshot.assemble(self)

    def consolidate_optimizer_state(self):
        """COLLECTIVE, called on EVERY rank before a checkpoint is written (only the master goes on to assemble and
        save the state): sharded optimizer state and EMA slices are merged; no-op for replicated optimizers."""
        consolidate = getattr(self.optimizer, "consolidate_state", None)
        if consolidate is not None:
            consolidate()

    def save_checkpoint(self, filename, extra_state):
        """Write the full training state (every rank assembles it, rank 0 writes); tensors go to the CPU."""
        logger.info("Saving checkpoint to {}".format(filename))
        state = utils.move_to_cpu(self.state_dict())
        state["extra_state"].update(extra_state)
        if self.should_save_checkpoint_on_current_rank:
            from unicore import checkpoint_utils

            checkpoint_utils.torch_persistent_save(state, filename)
        logger.info("Finished saving checkpoint to {}".format(filename))

    def load_checkpoint(self, filename, reset_optimizer=False, reset_lr_scheduler=False, reset_dataloader=False,
                        optimizer_overrides=None, reset_meters=False, **passthrough_args):
        """Restore training state: rank 0 reads the file and broadcasts it to the other ranks."""
        logger.info("Preparing to load checkpoint {}".format(filename))
        self._optim_history = []
        state = snapshot.read_and_share(self, filename)
        loaded, extra_state, last_optim_state = None, None, None
        if state is not None:
            last_optim_state = state.get("last_optimizer_state", None)
            ema_state = state.get("ema", None)
            loaded = snapshot.restore_weights(self, state, filename)
            extra_state = state.get("extra_state", None)
            self._optim_history = state.get("optimizer_history", None)
            if self.ema is not None and loaded == "model":
                if ema_state is not None:
                    logger.info("Loading EMA state...")
                    self.ema.load_state_dict(ema_state)
                else:
                    logger.info("Cannot find EMA state in checkpoint, load model weight to ema directly")
                    self.ema.reset_from(self.get_model())

        epoch_itr = None
        if extra_state is not None:
            itr_state = extra_state["train_iterator"]
            if "previous_training_time" in extra_state:
                self.clock.resume_from(extra_state["previous_training_time"])
            snapshot.restore_meters(extra_state, reset_meters)
            if not reset_dataloader:
                epoch_itr = self.get_train_iterator(epoch=itr_state["epoch"], load_dataset=True, **passthrough_args)
                epoch_itr.load_state_dict(itr_state)
        resumed = epoch_itr is not None
        if epoch_itr is None:
            epoch_itr = self.get_train_iterator(epoch=1, load_dataset=True, **passthrough_args)
        self.init_total_train_steps(epoch_itr)

        if loaded is not None:
            # the 16-bit weights changed under the optimizer: its fp32 masters follow them (the arenas stay)
            if self._optimizer is None:
                self._build_optimizer()
            elif hasattr(self._optimizer, "sync_master_from_params"):
                self._optimizer.sync_master_from_params()
            else:
                self._build_optimizer()
        if last_optim_state is not None and not reset_optimizer:
            last = self._optim_history[-1]
            snapshot.check_compatible(last, self)
            if not reset_lr_scheduler:
                self.lr_scheduler.load_state_dict(last["lr_scheduler_state"])
            self.optimizer.load_state_dict(last_optim_state, optimizer_overrides)
            self.set_num_updates(last["num_updates"])

        if loaded == "model" and resumed:
            logger.info("Loaded checkpoint {} (epoch {} @ {} updates)".format(filename, epoch_itr.epoch, self.get_num_updates()))
        elif loaded == "model":
            logger.info("Loaded checkpoint {}".format(filename))
        elif loaded == "ema":
            logger.info("Loaded ema state from checkpoint {}".format(filename))
        else:
            logger.info("No existing checkpoint found {}".format(filename))
        self.lr_step(epoch_itr.epoch)
        return extra_state, epoch_itr

    # -- data ------------------------------------------------------------------------------------------------------
    def _iterator(self, dataset, batch_size, epoch, sharded=True, skip_invalid=True, disable_iterator_cache=False):
        a = self.args
        return self.task.get_batch_iterator(
            dataset=dataset, batch_size=batch_size, ignore_invalid_inputs=skip_invalid,
            required_batch_size_multiple=a.required_batch_size_multiple, seed=a.seed,
            num_shards=self.data_parallel_world_size if sharded else 1,
            shard_id=self.data_parallel_rank if sharded else 0,
            num_workers=a.num_workers, epoch=epoch, data_buffer_size=a.data_buffer_size,
            disable_iterator_cache=disable_iterator_cache,
        )

    def get_train_iterator(self, epoch, combine=True, load_dataset=True, data_selector=None, shard_batch_itr=True,
                           disable_iterator_cache=False):
        if load_dataset:
            logger.info("loading train data for epoch {}".format(epoch))
            self.task.load_dataset(self.args.train_subset, epoch=epoch, combine=combine, data_selector=data_selector)
        itr = self._iterator(self.task.dataset(self.args.train_subset), self.args.batch_size, epoch,
                             sharded=shard_batch_itr, disable_iterator_cache=disable_iterator_cache)
        self.reset_dummy_batch(itr.first_batch)
        return itr

    def get_valid_iterator(self, subset, disable_iterator_cache=False):
        # epoch fixed at 1 so that validation batches are identical across training epochs
        return self._iterator(self.task.dataset(subset), self.args.batch_size_valid, 1,
                              skip_invalid=self.args.skip_invalid_size_inputs_valid_test,
                              disable_iterator_cache=disable_iterator_cache)

    def init_total_train_steps(self, epoch_itr):
        if self.args.max_epoch > 0:
            self._total_train_steps = (len(epoch_itr) + 1) // self.args.update_freq[0] * self.args.max_epoch
        else:
            self._total_train_steps = self.args.max_update

    def begin_epoch(self, epoch):
        logger.info("begin training epoch {}".format(epoch))
        self.lr_step_begin_epoch(epoch)
        self.task.begin_epoch(epoch, self.get_model())

    def begin_valid_epoch(self, epoch):
        self.task.begin_valid_epoch(epoch, self.get_model())

    def reset_dummy_batch(self, batch):
        self._dummy_batch = batch

    def _prepare_sample(self, sample, is_dummy=False):
        """Returns ``(sample on the training device, is_dummy)``; an empty batch is replaced by the dummy batch (its
        gradients are ignored, it only keeps this rank inside the collectives)."""
        if isinstance(sample, str) and sample == "DUMMY":
            raise Exception(
                "Trying to use an uninitialized 'dummy' batch. This usually indicates that the total number of "
                "batches is smaller than the number of participating GPUs. Try reducing the batch size or using "
                "fewer GPUs."
            )
        if sample is None or len(sample) == 0:
            if self._dummy_batch is None or len(self._dummy_batch) == 0:
                raise RuntimeError("Invalid dummy batch: {}".format(self._dummy_batch))
            return self._prepare_sample(self._dummy_batch, is_dummy=True)[0], True
        if self.cuda:
            sample = utils.move_to_cuda(sample)
        if isinstance(self._dummy_batch, str) and self._dummy_batch == "DUMMY":
            self._dummy_batch = sample
        return sample, False

    # -- one update / one validation batch ---------------------------------------------------------------------------
    @metrics.aggregate("train")
    def train_step(self, samples, raise_oom=False):
        """Forward, backward and one parameter update over a list of micro-batches."""
        return UpdateStep(self, samples, raise_oom=raise_oom).run()

    @metrics.aggregate("valid")
    def valid_step(self, sample, raise_oom=False):
        """Forward in evaluation mode; the logging outputs are reduced across ranks."""
        with torch.no_grad():
            self.model.eval()
            self.loss.eval()
            sample, is_dummy = self._prepare_sample(sample)
            try:
                _loss, sample_size, log = self.task.valid_step(sample, self.model, self.loss)
            except RuntimeError as exc:
                if "out of memory" not in str(exc) or raise_oom:
                    raise
                self._log_oom(exc)
                logger.warning("ran out of memory in validation step, retrying batch")
                for p in self.model.parameters():
                    p.grad = None
                if self.cuda:
                    torch.cuda.empty_cache()
                return self.valid_step(sample, raise_oom=True)
            logs = [log]
            if is_dummy:
                sample_size = sample_size * 0.0
        if self.data_parallel_world_size > 1:
            logs, _ = self._sum_over_ranks(logs, [sample_size], ignore=is_dummy, is_train=False)
        return logs

    def _sum_over_ranks(self, logs, extras, ignore=False, is_train=False):
        """Cross-rank sum of logging outputs + extra scalars outside the training step."""
        group = self.data_parallel_process_group
        if not self.task.__class__.logging_outputs_can_be_summed(self.get_loss(), is_train=is_train):
            return gather_objects(logs, extras, group, max_size=getattr(self.args, "all_gather_list_size", 16384), ignore=ignore)
        ledger = StatLedger()
        keys = ledger.add_logging_outputs(logs, ignore=ignore)
        for i, v in enumerate(extras):
            ledger.add("extra:{}".format(i), v)
        reduce_ledger(ledger, self.device, group, engine=self.dp_engine)
        return ledger.logging_output(keys), [ledger.value("extra:{}".format(i)) for i in range(len(extras))]

    def zero_grad(self):
        self.optimizer.zero_grad()

    def clip_grad_norm(self, clip_norm):
        return self.optimizer.clip_grad_norm(clip_norm)

    # -- learning rate / progress ---------------------------------------------------------------------------------------
    def lr_step_begin_epoch(self, epoch):
        self.lr_scheduler.step_begin_epoch(epoch)
        return self.lr_step_update()

    def lr_step(self, epoch, val_loss=None):
        self.lr_scheduler.step(epoch, val_loss)
        return self.lr_step_update()

    def lr_step_update(self):
        rate = self.lr_scheduler.step_update(self.get_num_updates())
        if isinstance(rate, dict):
            for name, value in rate.items():
                metrics.log_scalar("lr_{}".format(name), value, weight=0, priority=300)
            return rate.get("default", next(iter(rate.values())))
        metrics.log_scalar("lr", rate, weight=0, priority=300)
        return rate

    def get_lr(self):
        return self.optimizer.get_lr()

    def get_num_updates(self):
        return self._progress

    def set_num_updates(self, num_updates):
        self._progress = num_updates
        self.lr_step_update()
        metrics.log_scalar("num_updates", self._progress, weight=0, priority=200)

    def cumulative_training_time(self):
        return self.clock.cumulative()

    def _on_late_overflow(self, message):
        """Deferred overflow check: the skipped update was counted optimistically; take it back."""
        logger.info("NOTE: gradient overflow detected (update skipped on the device), ignoring gradient, " + message)
        self.set_num_updates(max(0, self.get_num_updates() - 1))

    # -- diagnostics ----------------------------------------------------------------------------------------------------
    def _log_oom(self, exc):
        logger.warning("OOM: Ran out of memory with exception: {}".format(exc))
        if torch.cuda.is_available() and hasattr(torch.cuda, "memory_summary"):
            for idx in range(torch.cuda.device_count()):
                logger.warning(torch.cuda.memory_summary(device=idx))
        sys.stderr.flush()

    def _check_grad_norms(self, grad_norm):
        """Non-finite norm => ``FloatingPointError``; all replicas must report the same norm (they would have drifted
        apart otherwise).  Not used with the fused tail: there every rank derives the norm from the same exchanged rows
        in the same order, and parameters are re-broadcast from their owning shard every update."""
        args = self.args
        on_device = torch.is_tensor(grad_norm) and grad_norm.is_cuda
        deferred = getattr(args, "deferred_overflow_check", False)
        if on_device and deferred and getattr(self.optimizer, "scaler", None) is not None:
            return  # the optimizer reads the norm itself before the next backward (resolve_pending_overflow)
        world = self.data_parallel_world_size
        if world > 1 and not getattr(args, "no_grad_norm_check", False):
            rows = torch.zeros(world, dtype=torch.double, device=self.device)
            rows[self.data_parallel_rank] = torch.as_tensor(grad_norm, dtype=torch.double)
            engine = self.dp_engine
            if hasattr(engine, "reduce_stats") and self.device.type == "cuda" and world <= 64:
                rows = engine.reduce_stats(rows)
            else:
                rows = rows.to(distributed_utils._backend_device())  # noqa: SLF001
                distributed_utils.all_reduce(rows, group=self.data_parallel_process_group)
            norms = utils.tolist(rows)  # the single host read of this tail
            lead = norms[0]
            same = all(n == n and abs(n) != float("inf") and abs(n - lead) / (lead + 1e-6) < 1e-6 for n in norms)
            if not same:
                table = "\n".join("rank {:3d} = {:.8f}".format(r, n) for r, n in enumerate(norms))
                raise FloatingPointError(
                    "Fatal error: gradients are inconsistent between workers. Try --ddp-backend=legacy_ddp. "
                    "Or are you mixing up different generation of GPUs in training?\n"
                    + "-" * 80 + "\ngrad_norm across the workers:\n{}\n".format(table) + "-" * 80
                )
            return
        if on_device and deferred:
            # no loss scaler (bf16 / fp32): the norm is only inspected for NaN / Inf - look at the PREVIOUS update's
            # value, long since on the host, instead of waiting for this one
            previous, self._norm_in_flight = getattr(self, "_norm_in_flight", None), utils.AsyncHostRead(grad_norm)
            if previous is None:
                return
            value = float(previous.get())
        else:
            value = float(grad_norm)
        if value != value or abs(value) == float("inf"):
            raise FloatingPointError("gradients are Nan/Inf")

    def _reduce_and_log_stats(self, logging_outputs, sample_size, grad_norm=None):
        if grad_norm is not None:
            metrics.log_speed("ups", 1.0, priority=100, round=2)
            metrics.log_scalar("gnorm", grad_norm, priority=400, round=3)
            if self.args.clip_norm > 0:
                g = torch.as_tensor(grad_norm)
                metrics.log_scalar("clip", (g > self.args.clip_norm).to(g.dtype) * 100, priority=500, round=1)
        with metrics.aggregate() as agg:
            if logging_outputs is not None:
                # (single process: the scalars were staged to the host behind the forward pass - host numbers from here)
                logging_outputs = [utils.resolve_logging_output(log) for log in logging_outputs]
                self.task.reduce_metrics(logging_outputs, self.get_loss())
            if "loss" not in agg:
                if "loss" not in self._warned:
                    self._warned.add("loss")
                    logger.warning("Loss.reduce_metrics did not log a 'loss' value, which may break some functionality")
                metrics.log_scalar("loss", -1)
            return LazyStats(agg, sample_size)
