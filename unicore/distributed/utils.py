"""Process launch, rendezvous and collective helpers (one process per GPU, torch.distributed).

Parity map (reference ``unicore/distributed/utils.py``): ``infer_init_method:32`` (torchrun env /
SLURM / single-node spawn), ``distributed_init:109`` (+ NCCL warm-up all-reduce),
``distributed_main:147``, ``call_main:166``, group helpers ``:192-233``, tensor collectives
``:236-272``, ``all_gather_list:275``, ``all_reduce_dict:352``, ``broadcast_tensors:406``,
``broadcast_object:447``.

Differences by design:
* device-agnostic: buffers live on the backend's device (CUDA for NCCL, CPU for Gloo), so every
  helper also works in the CPU/gloo plumbing configuration (the reference hard-codes CUDA,
  SURVEY D10);
* ``all_gather_list`` is a real all-gather of length-prefixed pickles (the reference emulates it
  with an all-reduce over a zeroed ``world*max_size`` buffer);
* ``all_reduce_dict`` packs device-resident and host-resident values into ONE fp64 vector and
  issues a single collective (the reference issues up to two);
* checkpoints are broadcast as one flat byte tensor per dtype bucket instead of one broadcast per
  tensor.
"""
import io
import logging
import os
import pickle
import random
import socket
import struct
import subprocess
import warnings
from collections import OrderedDict
from dataclasses import dataclass
from datetime import timedelta
from typing import Any, Dict, List, Mapping, Optional

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------------
# rendezvous
# ------------------------------------------------------------------------------------------------
def is_master(args):
    return args.distributed_rank == 0


def _from_torchrun_env(args) -> bool:
    env = os.environ
    if not all(k in env for k in ("MASTER_ADDR", "MASTER_PORT", "WORLD_SIZE", "RANK")):
        return False
    args.distributed_init_method = "env://"
    args.distributed_world_size = int(env["WORLD_SIZE"])
    args.distributed_rank = int(env["RANK"])
    args.distributed_no_spawn = True  # the launcher already forked one process per device
    if "LOCAL_RANK" in env:
        args.device_id = int(env["LOCAL_RANK"])
    return True


def _from_slurm(args) -> bool:
    node_list = os.environ.get("SLURM_STEP_NODELIST") or os.environ.get("SLURM_JOB_NODELIST")
    if node_list is None or args.distributed_port <= 0:
        return False
    try:
        hostnames = subprocess.check_output(["scontrol", "show", "hostnames", node_list])
    except (subprocess.CalledProcessError, FileNotFoundError):
        return False
    args.distributed_init_method = "tcp://{host}:{port}".format(
        host=hostnames.split()[0].decode("utf-8"), port=args.distributed_port
    )
    nnodes = int(os.environ.get("SLURM_NNODES"))
    ntasks_per_node = os.environ.get("SLURM_NTASKS_PER_NODE")
    if ntasks_per_node is not None:
        ntasks_per_node = int(ntasks_per_node)
    else:
        ntasks = int(os.environ.get("SLURM_NTASKS"))
        if ntasks % nnodes != 0:
            raise RuntimeError("SLURM_NTASKS must be a multiple of SLURM_NNODES")
        ntasks_per_node = ntasks // nnodes
    if ntasks_per_node == 1:
        # one task per node: this process spawns one child per local GPU
        gpus_per_node = torch.cuda.device_count()
        if args.distributed_world_size % nnodes != 0:
            raise RuntimeError("world size must be a multiple of the node count")
        args.distributed_rank = int(os.environ.get("SLURM_NODEID")) * gpus_per_node
    else:
        if ntasks_per_node != args.distributed_world_size // nnodes:
            raise RuntimeError("SLURM tasks per node do not match --distributed-world-size")
        args.distributed_no_spawn = True
        args.distributed_rank = int(os.environ.get("SLURM_PROCID"))
        args.device_id = int(os.environ.get("SLURM_LOCALID"))
    return True


def _single_node(args) -> None:
    n_dev = torch.cuda.device_count()
    if args.distributed_world_size > max(1, n_dev) and not getattr(args, "cpu", False):
        raise RuntimeError(
            "world size is {} but only {} devices are visible".format(args.distributed_world_size, n_dev)
        )
    port = random.randint(10000, 20000)
    args.distributed_init_method = "tcp://127.0.0.1:{port}".format(port=port)


def infer_init_method(args, force_distributed=False):
    """Fill ``args.distributed_init_method`` / rank / world size from the environment."""
    if args.distributed_init_method is not None:
        return
    if _from_torchrun_env(args):
        return
    if _from_slurm(args):
        return
    if args.distributed_world_size > 1 or force_distributed:
        _single_node(args)


def _backend_device(args=None) -> torch.device:
    """Device on which collective buffers must live for the active backend."""
    backend = dist.get_backend() if dist.is_initialized() else getattr(args, "distributed_backend", "nccl")
    if "nccl" in str(backend) and torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def distributed_init(args):
    """Join the process group (idempotent) and return this process's rank."""
    if dist.is_available() and dist.is_initialized():
        warnings.warn("Distributed is already initialized, cannot initialize twice!")
    else:
        logger.info(
            "distributed init (rank {}): {}".format(args.distributed_rank, args.distributed_init_method)
        )
        backend = args.distributed_backend
        if backend == "nccl" and not torch.cuda.is_available():
            backend = "gloo"
            args.distributed_backend = backend
        dist.init_process_group(
            backend=backend,
            init_method=args.distributed_init_method,
            world_size=args.distributed_world_size,
            rank=args.distributed_rank,
            timeout=timedelta(seconds=int(os.environ.get("UNICORE_DIST_TIMEOUT", "90"))),
        )
        logger.info("initialized host {} as rank {}".format(socket.gethostname(), args.distributed_rank))
        # force communicator creation now (NCCL initialises lazily on the first collective)
        dist.all_reduce(torch.zeros(1, device=_backend_device(args)))

    args.distributed_rank = dist.get_rank()
    if is_master(args):
        logging.getLogger().setLevel(logging.INFO)
    else:
        logging.getLogger().setLevel(logging.WARNING)
    return args.distributed_rank


def distributed_main(i, main, args, kwargs):
    args.device_id = i
    if torch.cuda.is_available() and not getattr(args, "cpu", False):
        torch.cuda.set_device(args.device_id)
    if args.distributed_rank is None:  # spawned child
        args.distributed_rank = kwargs.pop("start_rank", 0) + i
    args.distributed_rank = distributed_init(args)
    after_init = kwargs.pop("after_distributed_init_fn", None)
    if after_init:
        args = after_init(args)
    main(args, **kwargs)
    if dist.is_initialized():
        dist.barrier(get_global_group())


def call_main(args, main, **kwargs):
    """Run ``main(args)`` in every worker: spawn children, or act as one (torchrun), or run single."""
    if args.distributed_init_method is None:
        infer_init_method(args)
    if args.distributed_init_method is not None:
        if not args.distributed_no_spawn:
            start_rank = args.distributed_rank
            args.distributed_rank = None  # assigned in the children
            kwargs["start_rank"] = start_rank
            nprocs = min(max(torch.cuda.device_count(), 1), args.distributed_world_size)
            if not torch.cuda.is_available() or getattr(args, "cpu", False):
                nprocs = args.distributed_world_size
            torch.multiprocessing.spawn(fn=distributed_main, args=(main, args, kwargs), nprocs=nprocs, join=True)
        else:
            local = int(os.environ.get("LOCAL_RANK", getattr(args, "device_id", 0)))
            distributed_main(local, main, args, kwargs)
    else:
        main(args, **kwargs)


# ------------------------------------------------------------------------------------------------
# groups
# ------------------------------------------------------------------------------------------------
def new_groups(grouped_ranks: List[List[int]]):
    groups = [dist.new_group(g) for g in grouped_ranks]
    mine = next(i for i, g in enumerate(grouped_ranks) if get_global_rank() in g)
    return groups[mine]


def get_rank(group=None):
    return dist.get_rank(group=group) if dist.is_initialized() else 0


def get_world_size(group=None):
    return dist.get_world_size(group=group) if dist.is_initialized() else 1


def get_global_group():
    """``None`` means the WORLD group to torch.distributed."""
    return None


def get_global_rank():
    return dist.get_rank() if dist.is_initialized() else 0


def get_global_world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def get_data_parallel_group():
    """Seam for future model-parallel layouts; data parallel == WORLD today."""
    return get_global_group()


def get_data_parallel_rank():
    return get_rank(get_data_parallel_group())


def get_data_parallel_world_size():
    return get_world_size(get_data_parallel_group())


# ------------------------------------------------------------------------------------------------
# tensor collectives
# ------------------------------------------------------------------------------------------------
_REDUCE_OPS = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}


def all_reduce(tensor, group=None, op="sum"):
    if op not in _REDUCE_OPS:
        raise NotImplementedError(op)
    dist.all_reduce(tensor, op=_REDUCE_OPS[op], group=group)
    return tensor


def broadcast(tensor, src, group=None):
    dist.broadcast(tensor, src=src, group=group)


def all_to_all(tensor, group=None):
    """Perform an all-to-all of equal splits of a 1-D tensor."""
    if tensor.dim() != 1:
        raise ValueError("all_to_all expects a 1-D tensor")
    if tensor.numel() % get_world_size(group=group) != 0:
        raise ValueError("tensor length must be divisible by the group size")
    output = torch.zeros_like(tensor)
    dist.all_to_all_single(output, tensor, group=group)
    return output


def all_gather(tensor, group=None, return_tensor=False):
    world = get_world_size(group=group)
    parts = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(parts, tensor, group=group)
    return torch.stack(parts, dim=0) if return_tensor else parts


# ------------------------------------------------------------------------------------------------
# python-object collectives
# ------------------------------------------------------------------------------------------------
_HEADER = struct.Struct(">I")


def all_gather_list(data, group=None, max_size=16384):
    """Gather arbitrary picklable objects from all ranks into a list (rank order).

    Each rank contributes at most ``max_size`` bytes (4-byte big-endian length header + pickle).
    """
    world = get_world_size(group=group)
    if world == 1:
        return [data]
    from unicore import utils

    payload = pickle.dumps(utils.move_to_cpu(data))
    need = _HEADER.size + len(payload)
    if need > max_size:
        raise ValueError("encoded data size ({}) exceeds max_size ({})".format(need, max_size))
    device = _backend_device()
    slot = torch.zeros(max_size, dtype=torch.uint8)
    slot[:need] = torch.frombuffer(bytearray(_HEADER.pack(len(payload)) + payload), dtype=torch.uint8)
    slot = slot.to(device)
    gathered = torch.empty(world * max_size, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, slot, group=group)
    gathered = gathered.cpu()
    out = []
    try:
        for r in range(world):
            chunk = gathered[r * max_size:(r + 1) * max_size]
            (n,) = _HEADER.unpack(bytes(chunk[: _HEADER.size].tolist()))
            if n > 0:
                out.append(pickle.loads(bytes(chunk[_HEADER.size:_HEADER.size + n].tolist())))
        return out
    except pickle.UnpicklingError:
        raise Exception(
            "Unable to unpickle data from other workers. all_gather_list requires all workers to enter the "
            "function together, so this error usually indicates that the workers have fallen out of sync "
            "somehow. Workers can fall out of sync if one of them runs out of memory, or if there are other "
            "conditions in your training script that can cause one worker to finish an epoch while other "
            "workers are still iterating over their portions of the data. Try rerunning with "
            "--ddp-backend=no_c10d and see if that helps."
        )


def all_reduce_dict(data: Mapping[str, Any], device, group=None) -> Dict[str, Any]:
    """Sum every value of ``data`` across ranks with a single fp64 all-reduce.

    Values may be python numbers or (scalar or small) tensors on any device; results are returned
    as tensors on ``device`` (host-origin values come back as 0-dim fp64 tensors).
    """
    keys = list(data.keys())
    if len(keys) == 0:
        return OrderedDict()
    comm_device = _backend_device()
    flat, shapes = [], []
    for k in keys:
        v = data[k]
        t = v if torch.is_tensor(v) else torch.tensor(v, dtype=torch.double)
        shapes.append(t.shape)
        flat.append(t.detach().to(device=comm_device, dtype=torch.double, non_blocking=True).reshape(-1))
    buf = torch.cat(flat)
    all_reduce(buf, group=group)
    out = OrderedDict()
    offset = 0
    for k, shape in zip(keys, shapes):
        n = 1
        for s in shape:
            n *= s
        piece = buf[offset:offset + n].reshape(shape)
        out[k] = piece.to(device) if device is not None else piece
        offset += n
    return out


@dataclass(frozen=True)
class _TensorPlaceholder:
    index: int


def _split_tensors(obj, bucket: List[torch.Tensor]):
    """Replace tensors inside ``obj`` by placeholders, collecting them into ``bucket``."""
    if torch.is_tensor(obj):
        bucket.append(obj)
        return _TensorPlaceholder(len(bucket) - 1)
    if isinstance(obj, dict):
        return type(obj)((k, _split_tensors(v, bucket)) for k, v in obj.items()) \
            if isinstance(obj, OrderedDict) else {k: _split_tensors(v, bucket) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_split_tensors(v, bucket) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_split_tensors(v, bucket) for v in obj)
    if isinstance(obj, set):
        return {_split_tensors(v, bucket) for v in obj}
    return obj


def _join_tensors(obj, bucket: List[torch.Tensor]):
    if isinstance(obj, _TensorPlaceholder):
        return bucket[obj.index]
    if isinstance(obj, dict):
        return type(obj)((k, _join_tensors(v, bucket)) for k, v in obj.items()) \
            if isinstance(obj, OrderedDict) else {k: _join_tensors(v, bucket) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_join_tensors(v, bucket) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_join_tensors(v, bucket) for v in obj)
    if isinstance(obj, set):
        return {_join_tensors(v, bucket) for v in obj}
    return obj


def _broadcast_pickle(obj, src_rank, group, dist_device):
    """Broadcast a picklable (tensor-free) object: length, then bytes."""
    if get_rank(group) == src_rank:
        blob = io.BytesIO()
        torch.save(obj, blob)
        data = torch.frombuffer(bytearray(blob.getbuffer()), dtype=torch.uint8).to(dist_device)
        length = torch.tensor([data.numel()], dtype=torch.long, device=dist_device)
        broadcast(length, src=src_rank, group=group)
        broadcast(data, src=src_rank, group=group)
        return obj
    length = torch.zeros(1, dtype=torch.long, device=dist_device)
    broadcast(length, src=src_rank, group=group)
    data = torch.empty(int(length.item()), dtype=torch.uint8, device=dist_device)
    broadcast(data, src=src_rank, group=group)
    return torch.load(io.BytesIO(data.cpu().numpy().tobytes()), map_location="cpu", weights_only=False)


def broadcast_tensors(tensors: Optional[List[torch.Tensor]], src_rank: int, group=None, dist_device=None):
    """Broadcast a list of tensors; metadata first, then ONE flat buffer per dtype."""
    if dist_device is None:
        dist_device = _backend_device()
    is_src = get_rank(group) == src_rank
    meta = [{"size": t.size(), "dtype": t.dtype, "device": t.device} for t in tensors] if is_src else None
    meta = _broadcast_pickle(meta, src_rank, group, dist_device)
    by_dtype: Dict[torch.dtype, List[int]] = OrderedDict()
    for i, m in enumerate(meta):
        by_dtype.setdefault(m["dtype"], []).append(i)
    out: List[Optional[torch.Tensor]] = [None] * len(meta)
    for dtype, idxs in by_dtype.items():
        total = sum(int(torch.Size(meta[i]["size"]).numel()) for i in idxs)
        if is_src:
            flat = torch.cat([tensors[i].detach().reshape(-1).to(dist_device) for i in idxs]) if total > 0 \
                else torch.empty(0, dtype=dtype, device=dist_device)
        else:
            flat = torch.empty(total, dtype=dtype, device=dist_device)
        if total > 0:
            broadcast(flat, src=src_rank, group=group)
        offset = 0
        for i in idxs:
            n = int(torch.Size(meta[i]["size"]).numel())
            piece = flat[offset:offset + n].view(meta[i]["size"])
            out[i] = tensors[i] if is_src else piece.to(meta[i]["device"]).clone()
            offset += n
    return out


def broadcast_object(obj: Any, src_rank: int, group=None, dist_device: Optional[torch.device] = None) -> Any:
    """Broadcast an arbitrary object (e.g. a checkpoint) from ``src_rank`` to all ranks."""
    if get_world_size(group) == 1:
        return obj
    if dist_device is None:
        dist_device = _backend_device()
    if get_rank(group) == src_rank:
        tensors: List[torch.Tensor] = []
        skeleton = _split_tensors(obj, tensors)
        _broadcast_pickle(skeleton, src_rank, group, dist_device)
        broadcast_tensors(tensors, src_rank, group, dist_device)
        return obj
    skeleton = _broadcast_pickle(None, src_rank, group, dist_device)
    tensors = broadcast_tensors(None, src_rank, group, dist_device)
    return _join_tensors(skeleton, tensors)
