"""Multi-GPU plumbing: one process per GPU, the vectorised settings sharded across ranks.

Every entry of the vector batch (a lattice setting and/or a beam) is independent -- no
reduction ever crosses it -- so the path needs NO per-step collective: rank r
takes a contiguous slice of every vectorised parameter, scalars and a shared beam are
replicated.  The only communication is one setup-time broadcast from rank 0 (NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests) so that all ranks hold identical lattice
scalars and the identical shared beam.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_settings: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of ``n_settings`` for ``rank``."""
    base, extra = divmod(n_settings, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_module(module: torch.nn.Module, src: int = 0) -> int:
    """Broadcast every buffer/parameter of ``module`` (a Segment or a beam) from ``src``.

    Returns the number of bytes broadcast.  A no-op without an initialised process group.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    total = 0
    for tensor in list(module.buffers()) + list(module.parameters()):
        if tensor.numel() == 0:
            continue
        dist.broadcast(tensor, src=src)
        total += tensor.numel() * tensor.element_size()
    return total


def shard_segment(segment, n_settings: int, rank: int, world_size: int):
    """Slice, in place, every parameter whose leading dimension is the settings batch."""
    begin, end = shard_bounds(n_settings, rank, world_size)
    from .lowering import flatten

    for element in flatten([segment]):
        for name, tensor in list(element.named_buffers(recurse=False)):
            if tensor.dim() >= 1 and tensor.shape[0] == n_settings and name != "misalignment":
                setattr(element, name, tensor[begin:end].contiguous())
            elif name == "misalignment" and tensor.dim() == 2 and tensor.shape[0] == n_settings:
                setattr(element, name, tensor[begin:end].contiguous())
    return begin, end
