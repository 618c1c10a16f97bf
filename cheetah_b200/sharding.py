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


# trailing (non-vector) dimensions of the fields that have any; every other field is a scalar
# per setting
_INNER_DIMS = {"misalignment": 1, "pixel_size": 1, "resolution": 1, "predefined_transfer_map": 2}


def shard_segment(segment, n_settings: int, rank: int, world_size: int):
    """Slice, in place, every buffer and parameter whose VECTOR shape starts with the settings
    batch.  The vector shape is what is left of a tensor's shape after the field's own trailing
    dimensions (2 for ``misalignment`` / ``pixel_size``, 7 x 7 for ``predefined_transfer_map``),
    so a (2,) ``pixel_size`` is not mistaken for two settings, nor a (7, 7) map for seven."""
    begin, end = shard_bounds(n_settings, rank, world_size)
    from .lowering import flatten

    for element in flatten([segment]):
        named = list(element.named_buffers(recurse=False)) + list(
            element.named_parameters(recurse=False))
        for name, tensor in named:
            vector_dims = tensor.dim() - _INNER_DIMS.get(name, 0)
            if vector_dims < 1 or tensor.shape[0] != n_settings:
                continue
            piece = tensor.detach()[begin:end].contiguous()
            if name in element._parameters:
                element._parameters[name] = torch.nn.Parameter(
                    piece, requires_grad=tensor.requires_grad)
            else:
                setattr(element, name, piece)
    return begin, end


def gather_settings(tensor: torch.Tensor, n_settings: int) -> torch.Tensor:
    """All ranks' slices of a per-setting tensor ``(shard, ...)`` -> ``(n_settings, ...)`` on
    every rank.  Shards differ by at most one setting (``shard_bounds``), so each is padded to
    the largest one for the collective and trimmed afterwards.  Meant for REDUCED observables
    (a few numbers per setting); the tracked particles ``(B, N, 7)`` are never gathered
    (SURVEY.md 8e)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    world = dist.get_world_size()
    largest = -(-n_settings // world)
    padded = tensor.new_zeros((largest, *tensor.shape[1:]))
    padded[: tensor.shape[0]] = tensor
    pieces = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(pieces, padded.contiguous())
    sizes = [end - begin for begin, end in (shard_bounds(n_settings, r, world) for r in range(world))]
    return torch.cat([piece[:size] for piece, size in zip(pieces, sizes)], dim=0)


def gather_moments(observed, n_settings: int):
    """``BeamMoments`` of this rank's settings -> ``BeamMoments`` of the whole batch on every rank
    (mu, sigma, survivors, covariance: at most 49 numbers per setting)."""
    from .tracking import BeamMoments

    def gather(tensor, inner_dims: int):
        # scalars shared by all settings (a common energy or s) are replicated, not gathered
        if tensor is None or tensor.dim() <= inner_dims:
            return tensor
        return gather_settings(tensor, n_settings)

    return BeamMoments(
        gather(observed.mu, 1), gather(observed.sigma, 1),
        gather(observed.num_particles_survived, 0), gather(observed.energy, 0),
        gather(observed.s, 0), cov=gather(observed.cov, 2),
    )
