"""Multi-GPU plumbing: one process per GPU.  torch.distributed is used ONLY to pass the NCCL
unique id (128 bytes) from rank 0 to the other ranks and for host-side barriers; the data-path
collectives are NCCL calls made by libmellon_b200.so on its own stream."""

from __future__ import annotations

import os

import numpy as np


def _ensure_process_group():
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend="gloo")
    return dist


def exchange_unique_id(rank, world, make_id=None):
    """Rank 0 creates the NCCL unique id, everyone receives it (gloo broadcast on the host)."""
    import torch

    dist = _ensure_process_group()
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        if make_id is None:
            import ctypes as C

            from . import _native as nat

            raw = C.create_string_buffer(128)
            nat.check(nat.load_library().mb_comm_unique_id(raw), "mb_comm_unique_id")
            uid = raw.raw
        else:
            uid = make_id()
        buf = torch.from_numpy(np.frombuffer(uid, dtype=np.uint8).copy())
    dist.broadcast(buf, src=0)
    return bytes(buf.numpy().tobytes())


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def host_max(value):
    """Max of a host scalar over ranks (timing: max over ranks of the device-measured ms)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def host_sum(value):
    """Sum of a host scalar over ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
