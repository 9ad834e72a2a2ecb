"""Symmetric memory for one NVSwitch domain: the framework's own arena (``csrc/comm/symm_mem.cpp``).

Every rank creates a physical allocation with the CUDA virtual-memory API, exports it as a POSIX file descriptor,
hands the descriptor to every peer (``SCM_RIGHTS`` over abstract unix sockets whose names are exchanged through the
process group), imports and maps the peers' allocations, and - when the GPUs support NVLS - binds all of them to one
multicast object whose alias serves ``multimem.ld_reduce`` / ``multimem.st``.  The result is a ``SymmBuffer``: a local
tensor, the device addresses of all peers' copies, and the multicast address.

The reference has nothing comparable (it stops at ``dist.init_process_group``, ``unicore/distributed/utils.py:119-125``).
``torch.distributed._symmetric_memory`` remains available as a fallback provider
(``UNICORE_B200_SYMM_PROVIDER=torch``); the default ``auto`` uses the in-repo arena and falls back only if the driver
lacks the virtual-memory API or descriptor passing is not permitted in this container.
"""
import logging
import os
import socket
import struct
import uuid
from typing import List, Optional

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)

_KIND_PEER, _KIND_MULTICAST = 1, 2
_seq = 0


def provider_name() -> str:
    return os.environ.get("UNICORE_B200_SYMM_PROVIDER", "auto").lower()


class SymmBuffer:
    """One symmetric allocation: local tensor + peer addresses (+ multicast alias)."""

    def __init__(self, tensor, ptrs, multicast_ptr, rank, world, provider, keepalive=None):
        self.tensor = tensor
        self.ptrs = [int(p) for p in ptrs]
        self.multicast_ptr = int(multicast_ptr or 0)
        self.rank, self.world = int(rank), int(world)
        self.provider = provider
        self._keepalive = keepalive  # the native allocation / torch handle that owns the mappings


def _all_agree(flag: bool, group) -> bool:
    """True iff ``flag`` is true on every rank (collective)."""
    flags = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, bool(flag), group=group)
    return all(flags)


def _exchange_fds(group, rank: int, world: int, my_fd: int, mc_fd: Optional[int], expect_mc: bool):
    """Send ``my_fd`` to every peer (and ``mc_fd`` from rank 0); returns ({peer: fd}, multicast fd or None)."""
    global _seq
    _seq += 1
    name = "\0ub-symm-{}-{}-{}".format(uuid.uuid4().hex[:12], _seq, rank)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    try:
        srv.bind(name)
        srv.listen(world + 1)
        names: List[Optional[str]] = [None] * world
        dist.all_gather_object(names, name, group=group)  # also: everybody is listening from here on
        for peer in range(world):
            if peer == rank:
                continue
            with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                c.connect(names[peer])
                socket.send_fds(c, [struct.pack("ii", _KIND_PEER, rank)], [my_fd])
            if mc_fd is not None:
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:
                    c.connect(names[peer])
                    socket.send_fds(c, [struct.pack("ii", _KIND_MULTICAST, rank)], [mc_fd])
        peers, got_mc = {}, None
        want = (world - 1) + (1 if (expect_mc and rank != 0) else 0)
        srv.settimeout(120.0)
        for _ in range(want):
            conn, _addr = srv.accept()
            with conn:
                msg, fds, _flags, _a = socket.recv_fds(conn, 8, 1)
                kind, sender = struct.unpack("ii", msg)
                if kind == _KIND_PEER:
                    peers[sender] = fds[0]
                else:
                    got_mc = fds[0]
        return peers, got_mc
    finally:
        srv.close()


def _allocate_native(nbytes, numel, dtype, device, group) -> SymmBuffer:
    from unicore_b200.ops._native import native

    C = native()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev_index = device.index if device.index is not None else torch.cuda.current_device()
    want_mc = bool(C.symm_multicast_supported(dev_index)) and world > 1 and os.environ.get("UNICORE_B200_NO_NVLS", "0") != "1"
    alloc = C.SymmAllocation(int(nbytes), int(dev_index), int(rank), int(world), want_mc)
    use_mc = _all_agree(alloc.multicast_possible(), group)
    my_fd = int(alloc.export_fd())
    mc_fd = None
    if use_mc and rank == 0:
        try:
            mc_fd = int(alloc.multicast_create())
        except RuntimeError as exc:  # the multicast object could not be created: everybody learns it below
            logger.warning("NVLS multicast unavailable: %s", exc)
            mc_fd = None
    if use_mc:
        ok = [None]
        if rank == 0:
            ok[0] = mc_fd is not None
        dist.broadcast_object_list(ok, src=dist.get_global_rank(group, 0) if hasattr(dist, "get_global_rank") else 0,
                                   group=group)
        use_mc = bool(ok[0])
    try:
        peers, got_mc = _exchange_fds(group, rank, world, my_fd, mc_fd if use_mc else None, use_mc)
    finally:
        os.close(my_fd)
        if mc_fd is not None:
            os.close(mc_fd)
    for peer, fd in peers.items():
        alloc.import_peer(int(peer), int(fd))  # (closes fd)
    if use_mc:
        good = True
        try:
            if rank != 0:
                alloc.multicast_import(int(got_mc))
            alloc.multicast_add_device()
        except RuntimeError as exc:
            logger.warning("NVLS multicast unavailable on rank %d: %s", rank, exc)
            good = False
        if _all_agree(good, group):  # (also the barrier "every device has been added")
            try:
                alloc.multicast_bind_and_map()
            except RuntimeError as exc:
                logger.warning("cuMulticastBindMem failed on rank %d: %s", rank, exc)
                good = False
            if not _all_agree(good, group):
                raise RuntimeError("NVLS multicast binding failed on some ranks; set UNICORE_B200_NO_NVLS=1")
    tensor = alloc.tensor(int(numel), dtype)
    mc_ptr = int(alloc.multicast_ptr()) if alloc.has_multicast() else 0
    return SymmBuffer(tensor, alloc.ptrs(), mc_ptr, rank, world, "native", keepalive=alloc)


def _allocate_torch(numel, dtype, device, group) -> SymmBuffer:
    import torch.distributed._symmetric_memory as symm_mem

    tensor = symm_mem.empty(numel, dtype=dtype, device=device)
    handle = symm_mem.rendezvous(tensor, group)
    mc = 0
    try:
        if getattr(handle, "has_multicast_support", False):
            mc = int(handle.multicast_ptr)
    except Exception:  # noqa: BLE001
        mc = 0
    return SymmBuffer(tensor, [int(p) for p in handle.buffer_ptrs], mc, handle.rank, handle.world_size, "torch",
                      keepalive=handle)


_native_broken = False


def allocate(numel: int, dtype: torch.dtype, device: torch.device, group) -> SymmBuffer:
    """Collective: a zero-filled symmetric buffer of ``numel`` elements on every rank of ``group``."""
    global _native_broken
    want = provider_name()
    nbytes = max(16, int(numel) * torch.empty((), dtype=dtype).element_size())
    buf = None
    if want in ("auto", "native") and not _native_broken:
        from unicore_b200.ops._native import native

        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        usable = hasattr(native(), "SymmAllocation") and bool(native().symm_mem_supported(dev_index))
        if _all_agree(usable, group):
            try:
                buf = _allocate_native(nbytes, numel, dtype, device, group)
                err = None
            except (RuntimeError, OSError) as exc:
                err = exc
            if not _all_agree(err is None, group):
                if want == "native":
                    raise RuntimeError("in-repo symmetric arena failed: {!r}".format(err))
                logger.warning("in-repo symmetric arena unavailable (%r); using torch.distributed._symmetric_memory", err)
                _native_broken, buf = True, None
        elif want == "native":
            raise RuntimeError("the CUDA driver on some rank lacks the virtual-memory / descriptor-sharing API")
        else:
            _native_broken = True
    if buf is None:
        buf = _allocate_torch(numel, dtype, device, group)
    buf.tensor.zero_()
    return buf
