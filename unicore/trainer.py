"""Data-parallel training engine.

``Trainer`` owns the (wrapped) model, loss, optimizer, LR scheduler and EMA and implements one
optimizer update (``train_step``), one validation step, LR plumbing, metric aggregation across
ranks and checkpoint (de)serialisation.  Public surface and semantics follow the reference
``unicore/trainer.py`` (``Trainer:30``, ``train_step:571``, ``valid_step:805``,
``save_checkpoint:286``, ``load_checkpoint:299``, ``get_train_iterator:484``, ``lr_step*:853-874``,
stat sync ``:967-1049``, grad-norm consistency ``:1051-1084``), see SURVEY.md §3.2 for the call
stack and Appendix C for the numerics contract (grad normalisation ``world / sum(sample_size)``,
loss-scale state machine, overflow => skipped update, seeding discipline).

B200-first differences:
* device-agnostic (CPU/gloo plumbing runs work; the reference hard-codes CUDA in several places);
* the post-backward tail is ``norm kernel -> (fp16: one host read) -> fused Adam kernel``
  (see ``unicore/optim/fp16_optimizer.py``) instead of ~10 launches and >=4 host syncs;
* per-step stats + grad-norm consistency travel in ONE small fp64 all-reduce/all-gather pair;
* ``--ddp-backend b200`` provides symmetric-memory gradient buffers to the optimizer arenas so
  gradients are produced directly where the peer-memory reduction kernels read them;
* EMA update is a single fused pass over the flat fp32 master weights.
"""
import contextlib
import logging
import os
import sys
import time
from itertools import chain
from typing import Any, Dict, List

import torch

from unicore import checkpoint_utils, models, optim, utils
from unicore.distributed import utils as distributed_utils
from unicore.ema import ExponentialMovingAverageModel
from unicore.logging import meters, metrics
from unicore.nan_detector import NanDetector
from unicore.optim import lr_scheduler

logger = logging.getLogger(__name__)


class Trainer(object):
    """Main class for data parallel training (all-reduce of gradients across replicas)."""

    def __init__(self, args, task, model, loss):
        self.args = args
        self.task = task

        shared = _catalog_shared_params(model)
        self.cuda = torch.cuda.is_available() and not getattr(args, "cpu", False)
        self.device = torch.device("cuda") if self.cuda else torch.device("cpu")

        self._loss = loss
        self._model = model
        if args.fp16:
            self._loss = self._loss.half()
            self._model = self._model.half()
        elif args.bf16:
            self._loss = self._loss.bfloat16()
            self._model = self._model.bfloat16()
        if not self.use_distributed_wrapper:  # the DP wrapper moves the module itself
            self._loss = self._loss.to(device=self.device)
            self._model = self._model.to(device=self.device)

        # dtype/device moves re-create Parameters: re-tie the ones that were shared
        for paths in shared:
            anchor = _get_module_by_path(self._model, paths[0])
            for other in paths[1:]:
                logger.info("detected shared parameter: {} <- {}".format(paths[0], other))
                _set_module_by_path(self._model, other, anchor)

        self._dummy_batch = None
        self._total_train_steps = None
        self._lr_scheduler = None
        self._num_updates = 0
        self._optim_history = None
        self._optimizer = None
        self._warn_once = set()
        self._wrapped_loss = None
        self._wrapped_model = None
        self._grad_norm_buf = None

        if self.cuda:
            self.cuda_env = utils.CudaEnvironment()
            if self.data_parallel_world_size > 1:
                self.cuda_env_arr = distributed_utils.all_gather_list(
                    self.cuda_env, group=distributed_utils.get_global_group()
                )
            else:
                self.cuda_env_arr = [self.cuda_env]
            if self.data_parallel_rank == 0:
                utils.CudaEnvironment.pretty_print_cuda_env_list(self.cuda_env_arr)
        else:
            self.cuda_env = None
            self.cuda_env_arr = None

        if args.validate_with_ema and not args.ema_decay > 0:
            raise ValueError("--validate-with-ema requires --ema-decay > 0")
        model = self.model  # builds the DP wrapper if needed
        if args.ema_decay > 0 and (self.data_parallel_rank == 0 or args.validate_with_ema):
            self.ema = ExponentialMovingAverageModel(
                args, model, args.ema_decay, is_flattened=(args.fp16 or args.bf16)
            )
        else:
            self.ema = None

        metrics.log_start_time("wall", priority=790, round=2)
        self._start_time = time.time()
        self._previous_training_time = 0
        self._cumulative_training_time = None

    def reinitialize(self):
        """Drop optimizer / wrappers so they are rebuilt (e.g. after swapping the model)."""
        self._lr_scheduler = None
        self._optimizer = None
        self._wrapped_loss = None
        self._wrapped_model = None

    # -- topology -----------------------------------------------------------------------------------
    @property
    def data_parallel_world_size(self):
        if self.args.distributed_world_size == 1:
            return 1
        return distributed_utils.get_data_parallel_world_size()

    @property
    def data_parallel_process_group(self):
        return distributed_utils.get_data_parallel_group()

    @property
    def data_parallel_rank(self):
        if self.args.distributed_world_size == 1:
            return 0
        return distributed_utils.get_data_parallel_rank()

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

    # -- lazily built components -------------------------------------------------------------------------
    @property
    def loss(self):
        if self._wrapped_loss is None:
            if utils.has_parameters(self._loss) and self.use_distributed_wrapper:
                self._wrapped_loss = models.DistributedUnicoreModel(
                    self.args, self._loss, process_group=self.data_parallel_process_group, device=self.device
                )
            else:
                self._wrapped_loss = self._loss
        return self._wrapped_loss

    @property
    def model(self):
        if self._wrapped_model is None:
            if self.use_distributed_wrapper:
                self._wrapped_model = models.DistributedUnicoreModel(
                    self.args, self._model, process_group=self.data_parallel_process_group, device=self.device
                )
            else:
                self._wrapped_model = self._model
        return self._wrapped_model

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

    def _build_optimizer(self):
        named = [
            (n, p)
            for n, p in chain(self.model.named_parameters(), self.loss.named_parameters())
            if p.requires_grad
        ]
        if self.args.per_sample_clip_norm > 0 and self.args.ddp_backend not in ("no_c10d", "legacy_ddp"):
            raise ValueError("--per-sample-clip-norm only supports --ddp-backend no_c10d")
        if self.args.fp16 or self.args.bf16:
            grad_alloc = getattr(self._dp_engine(), "alloc_grad_buffer", None)
            param_alloc = getattr(self._dp_engine(), "alloc_param_buffer", None)
            self._optimizer = optim.FP16Optimizer.build_optimizer(
                self.args, named, grad_alloc=grad_alloc, param_alloc=param_alloc
            )
            if self.args.allreduce_fp32_grad and self.args.ddp_backend not in ("no_c10d", "legacy_ddp"):
                raise ValueError("--allreduce-fp32-grad requires --ddp-backend no_c10d")
            engine = self._dp_engine()
            if engine is not None and hasattr(engine, "attach_optimizer"):
                engine.attach_optimizer(self._optimizer)
        else:
            self._optimizer = optim.build_optimizer(self.args, named)
        if hasattr(self._optimizer, "add_late_overflow_handler"):
            self._optimizer.add_late_overflow_handler(self._on_late_overflow)
        self._lr_scheduler = lr_scheduler.build_lr_scheduler(self.args, self._optimizer, self._total_train_steps)
        self._lr_scheduler.step_update(0)

    def _dp_engine(self):
        """The innermost data-parallel wrapper object (or None when not distributed)."""
        wrapped = self._wrapped_model
        return getattr(wrapped, "module", None) if wrapped is not None and wrapped is not self._model else None

    # -- checkpoints ------------------------------------------------------------------------------------
    def state_dict(self):
        state = {
            "args": self.args,
            "model": self.model.state_dict(),
            "loss": self.loss.state_dict() if utils.has_parameters(self.loss) else None,
            "optimizer_history": (self._optim_history or [])
            + [
                {
                    "loss_name": self.get_loss().__class__.__name__,
                    "optimizer_name": self.optimizer.__class__.__name__,
                    "lr_scheduler_state": self.lr_scheduler.state_dict(),
                    "num_updates": self.get_num_updates(),
                }
            ],
            "task_state": self.task.state_dict() if self.task is not None else {},
            "extra_state": {
                "metrics": metrics.state_dict(),
                "previous_training_time": self.cumulative_training_time(),
            },
        }
        if not self.args.no_save_optimizer_state:
            state["last_optimizer_state"] = self.optimizer.state_dict()
        if self.ema is not None:
            state["ema"] = self.ema.state_dict()
        return state

    def consolidate_optimizer_state(self):
        """Collective, called on EVERY rank before a checkpoint is written: a sharded optimizer gathers its state
        (no-op for the replicated optimizers)."""
        consolidate = getattr(self._optimizer, "consolidate_state", None)
        if consolidate is not None:
            consolidate()

    def save_checkpoint(self, filename, extra_state):
        """Write the full training state (rank 0 only); tensors are stored as fp32 on CPU."""
        logger.info("Saving checkpoint to {}".format(filename))
        state = utils.move_to_cpu(self.state_dict())
        state["extra_state"].update(extra_state)
        if self.should_save_checkpoint_on_current_rank:
            checkpoint_utils.torch_persistent_save(state, filename)
        logger.info("Finished saving checkpoint to {}".format(filename))

    def load_checkpoint(
        self,
        filename,
        reset_optimizer=False,
        reset_lr_scheduler=False,
        reset_dataloader=False,
        optimizer_overrides=None,
        reset_meters=False,
        **passthrough_args,
    ):
        """Restore training state: rank 0 reads the file and broadcasts it to the other ranks."""
        extra_state, self._optim_history, last_optim_state = None, [], None
        logger.info("Preparing to load checkpoint {}".format(filename))
        distributed = self.data_parallel_world_size > 1
        master = self.data_parallel_rank == 0
        group = self.data_parallel_process_group

        exists = os.path.isfile(filename) if master else None
        if distributed:
            exists = distributed_utils.broadcast_object(exists, src_rank=0, group=group)

        loaded_model = loaded_ema_as_model = False
        if exists:
            state = checkpoint_utils.load_checkpoint_to_cpu(filename) if master else None
            if distributed:
                logger.info("Broadcast checkpoint from rank_0")
                state = distributed_utils.broadcast_object(state, src_rank=0, group=group)
            last_optim_state = state.get("last_optimizer_state", None)
            ema_state = state.get("ema", None)
            try:
                if self.args.load_from_ema:
                    logger.info("loading ema state to model")
                    report = self.model.load_state_dict(ema_state["params"], strict=False, model_args=self.args)
                    loaded_ema_as_model = True
                else:
                    report = self.model.load_state_dict(state["model"], strict=False, model_args=self.args)
                    del state["model"]  # free host memory early
                    loaded_model = True
                if report is not None:
                    if report.missing_keys:
                        logger.warning("Error in loading model state, missing_keys " + str(report.missing_keys))
                    if report.unexpected_keys:
                        logger.warning("Error in loading model state, unexpected_keys " + str(report.unexpected_keys))
                if utils.has_parameters(self.get_loss()):
                    self.get_loss().load_state_dict(state["loss"], strict=True)
                    del state["loss"]
            except Exception:
                raise Exception(
                    "Cannot load model parameters from checkpoint {}; "
                    "please ensure that the architectures match.".format(filename)
                )
            extra_state = state.get("extra_state", None)
            self._optim_history = state.get("optimizer_history", None)
            if ema_state is not None and self.ema is not None and not self.args.load_from_ema:
                logger.info("Loading EMA state...")
                self.ema.load_state_dict(ema_state)
            elif self.ema is not None and not loaded_ema_as_model:
                logger.info("Cannot find EMA state in checkpoint, load model weight to ema directly")
                self.ema = ExponentialMovingAverageModel(
                    self.args, self._model, decay=self.ema.decay, is_flattened=(self.args.fp16 or self.args.bf16)
                )

        epoch_itr = None
        if extra_state is not None:
            itr_state = extra_state["train_iterator"]
            if "previous_training_time" in extra_state:
                self._previous_training_time = extra_state["previous_training_time"]
                self._start_time = time.time()
            if itr_state.get("version", 1) >= 2 and itr_state["iterations_in_epoch"] == 0:
                reset_meters = True  # checkpoint taken at an epoch boundary
            if "metrics" in extra_state and not reset_meters:
                metrics.load_state_dict(extra_state["metrics"])
                for meter in metrics.get_meters("default").values():
                    if isinstance(meter, meters.TimeMeter):
                        meter.reset()  # wall-clock anchors of the old process are meaningless now
            if not reset_dataloader:
                epoch_itr = self.get_train_iterator(epoch=itr_state["epoch"], load_dataset=True, **passthrough_args)
                epoch_itr.load_state_dict(itr_state)
        resumed_iterator = epoch_itr is not None
        if epoch_itr is None:
            epoch_itr = self.get_train_iterator(epoch=1, load_dataset=True, **passthrough_args)
        self.init_total_train_steps(epoch_itr)

        if last_optim_state is not None and not reset_optimizer:
            self._build_optimizer()  # params may have changed: rebuild arenas from the loaded model
            last = self._optim_history[-1]
            if last["loss_name"] != self.get_loss().__class__.__name__:
                raise ValueError(
                    "Loss does not match; please reset the optimizer (--reset-optimizer). {} vs {}".format(
                        last["loss_name"], self.get_loss().__class__.__name__
                    )
                )
            if last["optimizer_name"] != self.optimizer.__class__.__name__:
                raise ValueError(
                    "Optimizer does not match; please reset the optimizer (--reset-optimizer). {} vs {}".format(
                        last["optimizer_name"], self.optimizer.__class__.__name__
                    )
                )
            if not reset_lr_scheduler:
                self.lr_scheduler.load_state_dict(last["lr_scheduler_state"])
            self.optimizer.load_state_dict(last_optim_state, optimizer_overrides)
            self.set_num_updates(last["num_updates"])
        elif self._optimizer is not None and (loaded_model or loaded_ema_as_model):
            # optimizer existed before the load: refresh its fp32 masters from the new weights
            self._build_optimizer()

        if loaded_model:
            if resumed_iterator:
                logger.info(
                    "Loaded checkpoint {} (epoch {} @ {} updates)".format(filename, epoch_itr.epoch, self.get_num_updates())
                )
            else:
                logger.info("Loaded checkpoint {}".format(filename))
        elif loaded_ema_as_model:
            logger.info("Loaded ema state from checkpoint {}".format(filename))
        else:
            logger.info("No existing checkpoint found {}".format(filename))

        self.lr_step(epoch_itr.epoch)
        return extra_state, epoch_itr

    # -- data ------------------------------------------------------------------------------------------
    def get_train_iterator(
        self, epoch, combine=True, load_dataset=True, data_selector=None, shard_batch_itr=True,
        disable_iterator_cache=False,
    ):
        if load_dataset:
            logger.info("loading train data for epoch {}".format(epoch))
            self.task.load_dataset(self.args.train_subset, epoch=epoch, combine=combine, data_selector=data_selector)
        itr = self.task.get_batch_iterator(
            dataset=self.task.dataset(self.args.train_subset),
            batch_size=self.args.batch_size,
            ignore_invalid_inputs=True,
            required_batch_size_multiple=self.args.required_batch_size_multiple,
            seed=self.args.seed,
            num_shards=self.data_parallel_world_size if shard_batch_itr else 1,
            shard_id=self.data_parallel_rank if shard_batch_itr else 0,
            num_workers=self.args.num_workers,
            epoch=epoch,
            data_buffer_size=self.args.data_buffer_size,
            disable_iterator_cache=disable_iterator_cache,
        )
        self.reset_dummy_batch(itr.first_batch)
        return itr

    def init_total_train_steps(self, epoch_itr):
        if self.args.max_epoch > 0:
            self._total_train_steps = (len(epoch_itr) + 1) // self.args.update_freq[0] * self.args.max_epoch
        else:
            self._total_train_steps = self.args.max_update

    def get_valid_iterator(self, subset, disable_iterator_cache=False):
        return self.task.get_batch_iterator(
            dataset=self.task.dataset(subset),
            batch_size=self.args.batch_size_valid,
            ignore_invalid_inputs=self.args.skip_invalid_size_inputs_valid_test,
            required_batch_size_multiple=self.args.required_batch_size_multiple,
            seed=self.args.seed,
            num_shards=self.data_parallel_world_size,
            shard_id=self.data_parallel_rank,
            num_workers=self.args.num_workers,
            epoch=1,  # fixed so that validation batches are identical across training epochs
            data_buffer_size=self.args.data_buffer_size,
            disable_iterator_cache=disable_iterator_cache,
        )

    def begin_epoch(self, epoch):
        logger.info("begin training epoch {}".format(epoch))
        self.lr_step_begin_epoch(epoch)
        self.task.begin_epoch(epoch, self.get_model())

    def begin_valid_epoch(self, epoch):
        self.task.begin_valid_epoch(epoch, self.get_model())

    def reset_dummy_batch(self, batch):
        self._dummy_batch = batch

    # -- one optimizer update -----------------------------------------------------------------------
    def _sync_context(self, i, n_micro):
        """``no_sync`` for every micro-batch except the last (gradient accumulation)."""
        if self.data_parallel_world_size > 1 and hasattr(self.model, "no_sync") and i < n_micro - 1:
            return self.model.no_sync()
        return contextlib.ExitStack()

    def _forward_backward(self, samples):
        """Run fwd+bwd over the micro-batches. Returns (logging_outputs, sample_size, ooms)."""
        logging_outputs, sample_size, ooms = [], 0, 0
        for i, sample in enumerate(samples):
            sample, is_dummy = self._prepare_sample(sample)
            try:
                with self._sync_context(i, len(samples)):
                    # per-(update, micro-batch, rank) dropout stream: reproducible and rank-distinct
                    with utils.torch_seed(self.args.seed, self.get_num_updates(), i, self.data_parallel_rank):
                        loss, sample_size_i, logging_output = self.task.train_step(
                            sample=sample,
                            model=self.model,
                            loss=self.loss,
                            optimizer=self.optimizer,
                            update_num=self.get_num_updates(),
                            ignore_grad=is_dummy,
                        )
                        del loss
                    if self.args.per_sample_clip_norm > 0:
                        self.optimizer.per_sample_clip_grad_norm(self.args.per_sample_clip_norm)
                logging_outputs.append(logging_output)
                sample_size = sample_size + sample_size_i
                if self.cuda and self.get_num_updates() == 0:
                    torch.cuda.empty_cache()  # first step allocates the high-water mark
            except RuntimeError as e:
                if "out of memory" not in str(e):
                    raise
                self._log_oom(e)
                if self.data_parallel_world_size > 1:
                    raise  # the collective schedule can no longer match the other ranks
                logger.warning("attempting to recover from OOM in forward/backward pass")
                ooms += 1
                self.zero_grad()
                if self.cuda:
                    torch.cuda.empty_cache()
                return None, 0, ooms
            if is_dummy:
                sample_size = sample_size * 0.0 if torch.is_tensor(sample_size) else 0.0
        return logging_outputs, sample_size, ooms

    @metrics.aggregate("train")
    def train_step(self, samples, raise_oom=False):
        """Forward, backward and one parameter update over a list of micro-batches."""
        # Module.train() walks the whole module tree (~1.3 ms for BERT-base); only flip when needed
        if not self.model.training:
            self.model.train()
        if not self.loss.training:
            self.loss.train()
        self.zero_grad()
        metrics.log_start_time("train_wall", priority=800, round=2)

        logging_outputs, sample_size, ooms = self._forward_backward(samples)
        if logging_outputs is None:  # single-process OOM: skip the step
            return None

        if torch.is_tensor(sample_size):
            sample_size = sample_size.float()
        else:
            sample_size = float(sample_size)

        local_sample_size = sample_size
        if self._sync_stats():
            train_time = self._local_cumulative_training_time()
            logging_outputs, (sample_size, ooms, total_train_time) = self._aggregate_logging_outputs(
                logging_outputs, sample_size, ooms, train_time, ignore=False, is_train=True
            )
            # kept as the (device) tensor the all-reduce produced: converting it here would drain the launch
            # queue right after backward on every multi-GPU step; cumulative_training_time() converts lazily
            self._cumulative_training_time = total_train_time / self.data_parallel_world_size

        overflow = False
        grad_norm = None
        try:
            with torch.autograd.profiler.record_function("reduce-grads"):
                self.optimizer.all_reduce_grads(self.model)
                if utils.has_parameters(self.loss):
                    self.optimizer.all_reduce_grads(self.loss)

            with torch.autograd.profiler.record_function("multiply-grads"):
                # DP engines average over ranks; we want sum(grads) / sum(sample_size)
                numer = self.data_parallel_world_size if self._sync_stats() else 1
                self.optimizer.multiply_grads(numer / (sample_size if _is_nonzero_static(sample_size) else 1.0))

            with torch.autograd.profiler.record_function("clip-grads"):
                grad_norm = self.clip_grad_norm(self.args.clip_norm)

            self._check_grad_norms(grad_norm)

            with torch.autograd.profiler.record_function("optimizer"):
                # rank-invariant RNG stream for the update (stochastic rounding must agree on all ranks)
                with utils.torch_seed(self.args.seed, self.get_num_updates()):
                    self.task.optimizer_step(self.optimizer, model=self.model, update_num=self.get_num_updates())

            if self.ema is not None:
                with torch.autograd.profiler.record_function("ema"):
                    if self.args.fp16 or self.args.bf16:
                        self.ema.update(self.optimizer.fp32_params)
                    else:
                        self.ema.update(self.model.named_parameters())
        except FloatingPointError:
            # non-finite or inconsistent grad norm: replay under the NaN detector for a useful message
            self.zero_grad()
            with NanDetector(self.get_model()):
                for _, sample in enumerate(samples):
                    sample, _ = self._prepare_sample(sample)
                    self.task.train_step(
                        sample, self.model, self.loss, self.optimizer, self.get_num_updates(), ignore_grad=False
                    )
            raise
        except OverflowError as e:
            overflow = True
            logger.info("NOTE: gradient overflow detected, ignoring gradient, {}".format(str(e)))
            grad_norm = torch.tensor(0.0, device=self.device)
            self.zero_grad()
        except RuntimeError as e:
            if "out of memory" in str(e):
                self._log_oom(e)
                logger.error("OOM during optimization, irrecoverable")
            raise

        logging_output = None
        if not overflow:
            self.set_num_updates(self.get_num_updates() + 1)
            if self.cuda and self.cuda_env is not None:
                gb_used = torch.cuda.max_memory_allocated() / 1024 / 1024 / 1024
                torch.cuda.reset_peak_memory_stats()
                gb_free = self.cuda_env.total_memory_in_GB - gb_used
                metrics.log_scalar("gb_free", gb_free, priority=1500, round=1, weight=0)
            logging_output = self._reduce_and_log_stats(logging_outputs, sample_size, grad_norm)
            if (
                self.cuda
                and self.args.empty_cache_freq > 0
                and (self.get_num_updates() + self.args.empty_cache_freq - 1) % self.args.empty_cache_freq == 0
            ):
                torch.cuda.empty_cache()

        if self.args.fp16:
            metrics.log_scalar("loss_scale", self.optimizer.scaler.loss_scale, priority=700, round=4, weight=0)
        metrics.log_stop_time("train_wall")
        return logging_output

    @metrics.aggregate("valid")
    def valid_step(self, sample, raise_oom=False):
        """Forward in evaluation mode; logging outputs are reduced across ranks."""
        with torch.no_grad():
            self.model.eval()
            self.loss.eval()
            sample, is_dummy = self._prepare_sample(sample)
            try:
                _loss, sample_size, logging_output = self.task.valid_step(sample, self.model, self.loss)
            except RuntimeError as e:
                if "out of memory" in str(e) and not raise_oom:
                    self._log_oom(e)
                    logger.warning("ran out of memory in validation step, retrying batch")
                    for p in self.model.parameters():
                        if p.grad is not None:
                            p.grad = None
                    if self.cuda:
                        torch.cuda.empty_cache()
                    return self.valid_step(sample, raise_oom=True)
                raise
            logging_outputs = [logging_output]
            if is_dummy:
                sample_size = sample_size * 0.0 if torch.is_tensor(sample_size) else 0.0
        if self.data_parallel_world_size > 1:
            logging_outputs, (sample_size,) = self._aggregate_logging_outputs(
                logging_outputs, sample_size, ignore=is_dummy, is_train=False
            )
        return logging_outputs

    def zero_grad(self):
        self.optimizer.zero_grad()

    # -- LR plumbing -------------------------------------------------------------------------------------
    def lr_step_begin_epoch(self, epoch):
        self.lr_scheduler.step_begin_epoch(epoch)
        return self.lr_step_update()

    def lr_step(self, epoch, val_loss=None):
        self.lr_scheduler.step(epoch, val_loss)
        return self.lr_step_update()

    def lr_step_update(self):
        new_lr = self.lr_scheduler.step_update(self.get_num_updates())
        if isinstance(new_lr, dict):
            for k, v in new_lr.items():
                metrics.log_scalar("lr_{}".format(k), v, weight=0, priority=300)
            new_lr = new_lr.get("default", next(iter(new_lr.values())))
        else:
            metrics.log_scalar("lr", new_lr, weight=0, priority=300)
        return new_lr

    def get_lr(self):
        return self.optimizer.get_lr()

    def get_model(self):
        """The bare (unwrapped) model."""
        return self._model

    def get_loss(self):
        return self._loss

    def get_num_updates(self):
        return self._num_updates

    def set_num_updates(self, num_updates):
        self._num_updates = num_updates
        self.lr_step_update()
        metrics.log_scalar("num_updates", self._num_updates, weight=0, priority=200)

    def clip_grad_norm(self, clip_norm):
        return self.optimizer.clip_grad_norm(clip_norm)

    def cumulative_training_time(self):
        if self._cumulative_training_time is None:
            return self._local_cumulative_training_time()
        if torch.is_tensor(self._cumulative_training_time):
            self._cumulative_training_time = float(utils.item(self._cumulative_training_time))
        return self._cumulative_training_time

    def _local_cumulative_training_time(self):
        return time.time() - self._start_time + self._previous_training_time

    # -- samples ------------------------------------------------------------------------------------------
    def _prepare_sample(self, sample, is_dummy=False):
        if isinstance(sample, str) and sample == "DUMMY":
            raise Exception(
                "Trying to use an uninitialized 'dummy' batch. This usually indicates that the total number of "
                "batches is smaller than the number of participating GPUs. Try reducing the batch size or using "
                "fewer GPUs."
            )
        if sample is None or len(sample) == 0:
            if self._dummy_batch is None or len(self._dummy_batch) == 0:
                raise RuntimeError("Invalid dummy batch: {}".format(self._dummy_batch))
            sample, _ = self._prepare_sample(self._dummy_batch, is_dummy=True)
            return sample, True
        if self.cuda:
            sample = utils.move_to_cuda(sample)
        if isinstance(self._dummy_batch, str) and self._dummy_batch == "DUMMY":
            self._dummy_batch = sample
        return sample, False

    # -- cross-rank statistics ----------------------------------------------------------------------------
    def _sync_stats(self):
        return self.data_parallel_world_size > 1

    def _log_oom(self, exc):
        logger.warning("OOM: Ran out of memory with exception: {}".format(exc))
        if torch.cuda.is_available() and hasattr(torch.cuda, "memory_summary"):
            for idx in range(torch.cuda.device_count()):
                logger.warning(torch.cuda.memory_summary(device=idx))
        sys.stderr.flush()

    def _aggregate_logging_outputs(self, logging_outputs: List[Dict[str, Any]], *extra_stats_to_sum,
                                   ignore=False, is_train=False):
        if self.task.__class__.logging_outputs_can_be_summed(self.get_loss(), is_train=is_train):
            return self._fast_stat_sync_sum(logging_outputs, *extra_stats_to_sum, ignore=ignore)
        return self._all_gather_list_sync(logging_outputs, *extra_stats_to_sum, ignore=ignore)

    def _all_gather_list_sync(self, logging_outputs, *extra_stats_to_sum, ignore=False):
        """Gather arbitrary (picklable) logging outputs from all ranks."""
        if ignore:
            logging_outputs = []
        gathered = distributed_utils.all_gather_list(
            [logging_outputs] + list(extra_stats_to_sum),
            max_size=getattr(self.args, "all_gather_list_size", 16384),
            group=self.data_parallel_process_group,
        )
        columns = list(zip(*gathered))
        logging_outputs = list(chain.from_iterable(columns[0]))
        extras = [sum(col) for col in columns[1:]]
        return logging_outputs, extras

    def _fast_stat_sync_sum(self, logging_outputs, *extra_stats_to_sum, ignore=False):
        """Sum scalar logging outputs and the extra stats over ranks with ONE fp64 all-reduce."""
        payload = {}
        for i, stat in enumerate(extra_stats_to_sum):
            payload["extra_stats_" + str(i)] = stat
        keys = None
        if len(logging_outputs) > 0:
            keys = list(logging_outputs[0].keys())
            for k in keys:
                if ignore:
                    v = logging_outputs[0][k]
                    v = torch.zeros_like(v) if torch.is_tensor(v) else 0
                else:
                    v = sum(log[k] for log in logging_outputs if k in log)
                payload["logging_outputs_" + k] = v
        reduced = distributed_utils.all_reduce_dict(payload, device=self.device, group=self.data_parallel_process_group)
        extras = [reduced["extra_stats_" + str(i)] for i in range(len(extra_stats_to_sum))]
        outputs = [{k: reduced["logging_outputs_" + k] for k in keys}] if keys is not None else []
        return outputs, extras

    def _on_late_overflow(self, message):
        """Deferred overflow check: the skipped update was counted optimistically; take it back."""
        logger.info("NOTE: gradient overflow detected (update skipped on the device), ignoring gradient, " + message)
        self.set_num_updates(max(0, self.get_num_updates() - 1))

    def _check_grad_norms(self, grad_norm):
        """Non-finite norm => FloatingPointError; all ranks must agree on the norm (replicas in sync)."""
        if (torch.is_tensor(grad_norm) and grad_norm.is_cuda and getattr(self.args, "deferred_overflow_check", False)
                and getattr(self.optimizer, "scaler", None) is not None):
            return  # nothing is read from the device on this path; the scaler sees the norm before the next backward
        if self.data_parallel_world_size > 1 and not getattr(self.args, "no_grad_norm_check", False):
            world = self.data_parallel_world_size
            mine = torch.as_tensor(grad_norm, dtype=torch.double).reshape(1)
            device = distributed_utils._backend_device()  # noqa: SLF001
            buf = torch.zeros(world, dtype=torch.double, device=device)
            buf[self.data_parallel_rank] = mine.to(device)[0]
            distributed_utils.all_reduce(buf, group=self.data_parallel_process_group)
            norms = utils.tolist(buf)  # the single host read of the multi-GPU tail
            head = norms[0]
            finite = all(n == n and abs(n) != float("inf") for n in norms)
            consistent = finite and all(abs(n - head) / (head + 1e-6) < 1e-6 for n in norms)
            if not consistent:
                detail = "\n".join("rank {:3d} = {:.8f}".format(r, n) for r, n in enumerate(norms))
                raise FloatingPointError(
                    "Fatal error: gradients are inconsistent between workers. Try --ddp-backend=legacy_ddp. "
                    "Or are you mixing up different generation of GPUs in training?\n"
                    + "-" * 80 + "\ngrad_norm across the workers:\n{}\n".format(detail) + "-" * 80
                )
            return
        if torch.is_tensor(grad_norm) and grad_norm.is_cuda and getattr(self.args, "deferred_overflow_check", False):
            # bf16 / fp32 runs have no loss scaler, the norm is only inspected for NaN/Inf: look at the PREVIOUS
            # step's value (long since copied to the host) instead of waiting for this one
            pending, self._pending_norm_check = getattr(self, "_pending_norm_check", None), utils.AsyncHostRead(grad_norm)
            if pending is None:
                return
            value = float(pending.get())
        else:
            value = float(grad_norm)
        if value != value or abs(value) == float("inf"):
            raise FloatingPointError("gradients are Nan/Inf")

    def _reduce_and_log_stats(self, logging_outputs, sample_size, grad_norm=None):
        if grad_norm is not None:
            metrics.log_speed("ups", 1.0, priority=100, round=2)
            metrics.log_scalar("gnorm", grad_norm, priority=400, round=3)
            if self.args.clip_norm > 0:
                gn = torch.as_tensor(grad_norm)
                metrics.log_scalar("clip", (gn > self.args.clip_norm).to(gn.dtype) * 100, priority=500, round=1)
        with metrics.aggregate() as agg:
            if logging_outputs is not None:
                self.task.reduce_metrics(logging_outputs, self.get_loss())
                del logging_outputs
            if "loss" not in agg:
                if "loss" not in self._warn_once:
                    self._warn_once.add("loss")
                    logger.warning("Loss.reduce_metrics did not log a 'loss' value, which may break some functionality")
                metrics.log_scalar("loss", -1)
            return _LazyStats(agg, sample_size)


class _LazyStats(object):
    """The logging output of one ``train_step``: a read-only mapping that is materialised on first access.

    Producing the smoothed values means bringing device-resident meters to the host, i.e. waiting for the
    step to finish on the GPU.  Callers that only test ``is not None`` (the CLI between log intervals,
    the device-timed benchmark) never pay for that; callers that read a value get exactly what the eager
    version returned.
    """

    def __init__(self, agg, sample_size):
        self._agg, self._sample_size, self._values = agg, sample_size, None

    def _get(self):
        if self._values is None:
            out = self._agg.get_smoothed_values()
            out["sample_size"] = self._sample_size
            for key in ("ppl", "wps", "wpb", "bsz"):
                out.pop(key, None)
            self._values, self._agg = out, None
        return self._values

    def __getitem__(self, key):
        return self._get()[key]

    def __contains__(self, key):
        return key in self._get()

    def __iter__(self):
        return iter(self._get())

    def __len__(self):
        return len(self._get())

    def get(self, key, default=None):
        return self._get().get(key, default)

    def keys(self):
        return self._get().keys()

    def values(self):
        return self._get().values()

    def items(self):
        return self._get().items()

    def __repr__(self):
        return repr(self._get())


def _is_nonzero_static(sample_size) -> bool:
    """Tensors are assumed non-zero (no host sync); python numbers are checked."""
    return True if torch.is_tensor(sample_size) else sample_size > 0


# -- shared-parameter bookkeeping ---------------------------------------------------------------------------
def _catalog_shared_params(module, memo=None, prefix=""):
    """Return lists of dotted paths that refer to the same Parameter object (len > 1 only)."""
    first_call = memo is None
    if first_call:
        memo = {}
    for name, param in module._parameters.items():
        if param is None:
            continue
        memo.setdefault(param, []).append((prefix + "." if prefix else "") + name)
    for name, child in module._modules.items():
        if child is None:
            continue
        _catalog_shared_params(child, memo, (prefix + "." if prefix else "") + name)
    if first_call:
        return [paths for paths in memo.values() if len(paths) > 1]


def _get_module_by_path(module, path):
    for name in path.split("."):
        module = getattr(module, name)
    return module


def _set_module_by_path(module, path, value):
    parts = path.split(".")
    for name in parts[:-1]:
        module = getattr(module, name)
    setattr(module, parts[-1], value)
