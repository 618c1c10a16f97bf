"""``SpaceChargeKick.track`` and ``cloud_in_cell_charge_deposition`` on the CUDA library.

Host-side mirror of cheetah/accelerator/space_charge_kick.py:477-586: broadcast the beam's
vector dimensions, flatten them to one batch of B beams, run the seven-kernel chain (moments,
grid parameters, deposit, Green function, FFT Poisson solve, field, gather+kick) and restore
the vector shape.  All per-beam scalars stay on the device between kernels.
"""

from __future__ import annotations

import math
import os

import torch

from . import _capi
from .tracking import _bshape, _new_beam


def _flat(tensor: torch.Tensor, vector_shape: tuple, inner: tuple, dtype) -> tuple:
    """(contiguous tensor, batch stride in elements) of ``tensor`` viewed as [B, *inner]."""
    if tensor.dim() == len(inner) and tensor.dtype == dtype and tensor.is_contiguous():
        return tensor, 0  # nothing vectorised (the usual case): one value for every beam
    vshape = tuple(tensor.shape[: tensor.dim() - len(inner)])
    if tensor.dtype != dtype:
        tensor = tensor.to(dtype)
    if math.prod(vshape) == 1:
        return tensor.contiguous(), 0
    if vshape != tuple(vector_shape):
        tensor = tensor.expand(*vector_shape, *inner)
    return tensor.contiguous(), math.prod(inner)


def fft_len(n: int) -> int:
    """Padded transform length of a grid axis with ``n`` cells: the next power of two >= 2 n."""
    length = 8
    while length < 2 * n:
        length *= 2
    return length


def _scalar_ref(tensor: torch.Tensor, vector_shape: tuple, dtype) -> tuple:
    tensor, stride = _flat(tensor, vector_shape, (), dtype)
    return tensor, min(stride, 1)


class Prepared:
    """Grid parameters of a kick that the PREVIOUS kick's fused gather kernel already computed
    (``ch_sc_gather_kick_fused``): the kick can skip its moments pass."""

    def __init__(self, workspace, slot: int, element) -> None:
        self.workspace, self.slot, self.element = workspace, slot, element


class Workspace:
    """Scratch buffers of one kick (torch's caching allocator makes re-allocation cheap)."""

    @property
    def stats(self) -> torch.Tensor:
        return self.stats_slots[self.slot]

    @property
    def params(self) -> torch.Tensor:
        return self.params_slots[self.slot]

    def charge_grid(self) -> torch.Tensor:
        """Deposited charge per cell [B, nx, ny, nz] (sums the four quad-block parts, see
        ch_sc_deposit)."""
        quad = self.rho_quad  # [B, nx, 4, QY, QZ, 4]
        n_beams, nx, _, qy, qz, _ = quad.shape
        ny, nz = 2 * (qy - 1), 2 * (qz - 1)
        blocks = quad.reshape(n_beams, nx, 2, 2, qy, qz, 2, 2)  # py, pz, by, bz, dy, dz
        out = quad.new_zeros((n_beams, nx, ny + 3, nz + 3))  # index = cell + 1 (cells from -1)
        for py in (0, 1):
            for pz in (0, 1):
                # block (by, bz) covers cells y = 2 by - py + dy, z = 2 bz - pz + dz
                part = blocks[:, :, py, pz].permute(0, 1, 2, 4, 3, 5).reshape(
                    n_beams, nx, 2 * qy, 2 * qz)
                out[:, :, 1 - py : 1 - py + 2 * qy, 1 - pz : 1 - pz + 2 * qz] += part
        return out[:, :, 1 : ny + 1, 1 : nz + 1]

    def fixed_point_scratch(self) -> torch.Tensor:
        """int64 grid + scale word per beam for the deterministic deposit (allocated on demand)."""
        if self._fixed is None:
            n_beams, nx, ny, nz = self.phi.shape
            self._fixed = torch.empty((n_beams * (nx * ny * nz + 1),), dtype=torch.int64,
                                      device=self.phi.device)
        return self._fixed

    def __init__(self, n_beams: int, grid_shape: tuple, dtype, device) -> None:
        nx, ny, nz = grid_shape
        self._fixed = None
        cdtype = torch.complex64 if dtype == torch.float32 else torch.complex128
        # FFT length per axis: the next power of two >= 2 n (include/cheetah_b200.h)
        lx, ly, lz = (fft_len(v) for v in (nx, ny, nz))
        kx, ky, kz = lx // 2 + 1, ly // 2 + 1, lz // 2 + 1
        spectrum = (n_beams, lx, ly, kz)
        # two slots: the fused gather kernel of kick k writes the sums / grid parameters of kick
        # k + 1 while it is still reading those of kick k
        self.stats_slots = [
            torch.empty((n_beams, _capi.SC_STATS), dtype=torch.float64, device=device)
            for _ in range(2)
        ]
        self.params_slots = [
            torch.empty((n_beams, _capi.SC_PARAMS), dtype=torch.float64, device=device)
            for _ in range(2)
        ]
        self.slot = 0
        # quad blocks (see ch_sc_deposit); `charge_grid()` sums the four parts
        self.rho_quad = torch.empty(
            (n_beams, nx, 4, ny // 2 + 1, nz // 2 + 1, 4), dtype=dtype, device=device)
        self.lattice = torch.empty(
            (n_beams, nx + 1, ny + 1, nz + 1), dtype=torch.float64, device=device
        )
        self.green = None  # the mirrored (2n)^3 array is only built for the parity tests
        self.green_scratch = torch.empty(
            (n_beams * (nx * ny * kz + nx * ky * kz),), dtype=dtype, device=device
        )
        self.green_spectrum = torch.empty((n_beams, kx, ky, kz), dtype=dtype, device=device)
        self.rho_spectrum = torch.empty(spectrum, dtype=cdtype, device=device)
        self.phi = torch.empty((n_beams, nx, ny, nz), dtype=dtype, device=device)
        # float32: "bricks" (all 8 corners of a cell side by side, ch_sc_field_bricks);
        # float64: z corner pairs per node (ch_sc_field)
        self.bricks = dtype == torch.float32 and use_field_bricks
        if self.bricks:
            # one group of beams at a time (ch_sc_field_gather): the group's bricks stay in L2
            per_beam = nx * ny * nz * 96
            self.group_beams = max(1, min(n_beams, brick_group_bytes // per_beam))
            self.field = torch.empty((self.group_beams, nx, ny, nz, 3, 8), dtype=dtype,
                                     device=device)
        else:
            self.field = torch.empty((n_beams, nx, ny, nz, 2, 4), dtype=dtype, device=device)


# Bytes of field bricks built and consumed at a time: caps the scratch memory of a large batch.
# (Groups small enough to keep the bricks in L2 were measured SLOWER -- 128 beams: 5.1 ms with
# 1 beam per group, 4.4 with 2, 3.3 with all -- the gather is bound inside the SM, not by where
# its bricks come from, and short launches pay their tails.)
brick_group_bytes = int(os.environ.get("CH_BRICK_GROUP_MB", "4096")) << 20
# Set to False to gather from the node layout with float32 beams as well (tests compare both).
use_field_bricks = os.environ.get("CH_FIELD_NODES", "0") in ("", "0")

_workspace_cache: dict = {}
_side_streams: dict = {}


def _side_stream(device) -> torch.cuda.Stream:
    stream = _side_streams.get(device)
    if stream is None:
        # high priority: its (few, long-running) CTAs are placed before the deposit's
        stream = _side_streams[device] = torch.cuda.Stream(device, priority=-1)
    return stream


def _workspace(n_beams: int, grid_shape: tuple, dtype, device, stream: int) -> Workspace:
    """One cached workspace per (batch, grid, dtype, device, stream): kicks on a stream are
    ordered, so the scratch of the previous kick is free when the next one starts."""
    key = (n_beams, tuple(grid_shape), dtype, device, stream, use_field_bricks, brick_group_bytes)
    ws = _workspace_cache.get(key)
    if ws is None:
        if len(_workspace_cache) >= 4:
            _workspace_cache.clear()
        ws = _workspace_cache[key] = Workspace(n_beams, grid_shape, dtype, device)
    return ws


def kick(particles, energy, charges, survival, mass_eV, effect_length, extents, grid_shape,
         want_intermediates: bool = False, prepared: Prepared | None = None, element=None,
         fuse_records=None, next_element=None, records_ready=None, next_tensors=None):
    """Run one kick on already-broadcast inputs; returns (particles_out [B,N,7], workspace).

    ``prepared``: the grid parameters were computed by the previous kick (skip the moments pass).
    ``fuse_records`` (``[1 or B, record_len]`` compose records): apply that linear map to the
    kicked particles in the same pass (``records_ready``: CUDA event after which they are
    valid, when they were composed on another stream).  ``next_element``: also compute the moments
    and grid parameters of that following SpaceChargeKick; ``workspace.prepared_next`` is then
    set."""
    device, dtype = particles.device, particles.dtype
    lib = _capi.lib()
    code = _capi.dtype_code(dtype)
    nx, ny, nz = (int(v) for v in grid_shape)
    for v in (nx, ny, nz):
        if v < 2 or v > 256:
            raise NotImplementedError(
                f"cheetah_b200 SpaceChargeKick supports grid sizes in [2, 256], got "
                f"{tuple(grid_shape)}"
            )

    vector_shape = _bshape(
        particles.shape[:-2], energy.shape, charges.shape[:-1], survival.shape[:-1],
        effect_length.shape, *(e.shape for e in extents), (1,),
    )
    n_beams = math.prod(vector_shape)
    n = particles.shape[-2]
    p, p_stride = _flat(particles, vector_shape, (n, 7), dtype)
    q, q_stride = _flat(charges, vector_shape, (n,), dtype)
    w, w_stride = _flat(survival, vector_shape, (n,), dtype)
    e, e_stride = _scalar_ref(energy, vector_shape, dtype)
    length, l_stride = _scalar_ref(effect_length, vector_shape, dtype)
    ext = [_scalar_ref(x, vector_shape, dtype) for x in extents]
    mass = mass_eV if mass_eV.dtype in (torch.float32, torch.float64) else mass_eV.to(dtype)

    stream = _capi.current_stream(device)
    ws = _workspace(n_beams, (nx, ny, nz), dtype, device, stream)
    out = torch.empty((n_beams, n, 7), dtype=dtype, device=device)
    forces = torch.empty((n_beams, n, 3), dtype=dtype, device=device) if want_intermediates else None
    ws.prepared_next = None
    # torch.use_deterministic_algorithms: fixed-order moment sums and fixed-point deposit, and no
    # moments fused into the gather (its cross-CTA sums are float64 atomics)
    deterministic = torch.are_deterministic_algorithms_enabled()
    if deterministic:
        next_element = None
    reuse = (
        prepared is not None and prepared.workspace is ws and prepared.element is element
        and element is not None
    )
    with _capi.device_guard(device):
        if reuse:
            ws.slot = prepared.slot
        else:
            moment_args = (
                p.data_ptr(), p_stride, w.data_ptr(), w_stride, n, n_beams,
                e.data_ptr(), e_stride, _capi.dtype_code(e.dtype),
                mass.data_ptr(), _capi.dtype_code(mass.dtype),
                length.data_ptr(), l_stride, _capi.dtype_code(length.dtype),
                ext[0][0].data_ptr(), ext[0][1], ext[1][0].data_ptr(), ext[1][1],
                ext[2][0].data_ptr(), ext[2][1], code, nx, ny, nz, code,
            )
            if deterministic:
                partials = torch.empty((n_beams * ((n + 1023) // 1024) * 8,), dtype=torch.float64,
                                       device=device)
                _capi.check(lib.ch_sc_moments_and_params_deterministic(
                    *moment_args, partials.data_ptr(), ws.stats.data_ptr(), ws.params.data_ptr(),
                    stream))
            else:
                _capi.check(lib.ch_sc_moments_and_params(
                    *moment_args, ws.stats.data_ptr(), ws.params.data_ptr(), stream))
        if deterministic or want_intermediates:
            main = torch.cuda.current_stream(device)
            # the Green-function chain only needs the grid parameters: it runs on a side stream
            # concurrently with the deposit and the first two FFT passes of the charge
            side = _side_stream(device)
            forked = torch.cuda.Event()
            forked.record(main)
            side.wait_event(forked)
            ws.green = (
                torch.empty((n_beams, 2 * nx, 2 * ny, 2 * nz), dtype=dtype, device=device)
                if want_intermediates else None
            )
            _capi.check(lib.ch_sc_green_function(
                ws.params.data_ptr(), n_beams, nx, ny, nz, code, ws.lattice.data_ptr(),
                _capi.ptr(ws.green), side.cuda_stream))
            _capi.check(lib.ch_sc_green_spectrum(
                ws.lattice.data_ptr(), ws.params.data_ptr(), n_beams, nx, ny, nz, code,
                ws.green_scratch.data_ptr(),
                ws.green_spectrum.data_ptr(), side.cuda_stream))
            joined = torch.cuda.Event()
            joined.record(side)
            if deterministic:
                # fixed-point accumulation: bit-identical from run to run
                _capi.check(lib.ch_sc_deposit_deterministic(
                    p.data_ptr(), p_stride, q.data_ptr(), q_stride, w.data_ptr(), w_stride,
                    ws.params.data_ptr(), n, n_beams, nx, ny, nz, code,
                    ws.fixed_point_scratch().data_ptr(), ws.rho_quad.data_ptr(), stream))
            else:
                _capi.check(lib.ch_sc_deposit(
                    p.data_ptr(), p_stride, q.data_ptr(), q_stride, w.data_ptr(), w_stride,
                    ws.params.data_ptr(), n, n_beams, nx, ny, nz, code, ws.rho_quad.data_ptr(),
                    stream))
            main.wait_event(joined)
            _capi.check(lib.ch_sc_poisson_solve(
                ws.rho_quad.data_ptr(), ws.green_spectrum.data_ptr(), ws.params.data_ptr(),
                n_beams, nx, ny, nz, code, ws.rho_spectrum.data_ptr(), ws.phi.data_ptr(), stream))
        else:
            # the same four stages in one C call (its own side stream and events)
            ws.green = None
            _capi.check(lib.ch_sc_solve(
                p.data_ptr(), p_stride, q.data_ptr(), q_stride, w.data_ptr(), w_stride,
                ws.params.data_ptr(), n, n_beams, nx, ny, nz, code, ws.rho_quad.data_ptr(),
                ws.lattice.data_ptr(), ws.green_scratch.data_ptr(),
                ws.green_spectrum.data_ptr(), ws.rho_spectrum.data_ptr(), ws.phi.data_ptr(),
                stream))
        if records_ready is not None:
            torch.cuda.current_stream(device).wait_event(records_ready)
        nxt = None
        next_slot = 1 - ws.slot
        if next_element is not None:
            following = next_tensors or KickTensors(next_element)
            nlen, nl_stride = _scalar_ref(following.effect_length, vector_shape, dtype)
            next = [_scalar_ref(x, vector_shape, dtype) for x in following.extents]
            nxt = (nlen, nl_stride, next)
        record_stride = 0
        if fuse_records is not None and fuse_records.shape[0] > 1:
            record_stride = fuse_records.shape[1]
        fusion_args = (
            _capi.ptr(fuse_records), record_stride, w.data_ptr(), w_stride,
            ws.stats_slots[next_slot].data_ptr() if nxt else None,
            ws.params_slots[next_slot].data_ptr() if nxt else None,
            e.data_ptr(), e_stride, _capi.dtype_code(e.dtype),
            mass.data_ptr(), _capi.dtype_code(mass.dtype),
            nxt[0].data_ptr() if nxt else None, nxt[1] if nxt else 0,
            _capi.dtype_code(nxt[0].dtype) if nxt else code,
            nxt[2][0][0].data_ptr() if nxt else None, nxt[2][0][1] if nxt else 0,
            nxt[2][1][0].data_ptr() if nxt else None, nxt[2][1][1] if nxt else 0,
            nxt[2][2][0].data_ptr() if nxt else None, nxt[2][2][1] if nxt else 0, code,
            nx, ny, nz,
        )
        if ws.bricks:
            # float32: field bricks and gather interleaved per group of beams
            _capi.check(lib.ch_sc_field_gather(
                p.data_ptr(), p_stride, ws.phi.data_ptr(), ws.field.data_ptr(), ws.group_beams,
                ws.params.data_ptr(), n, n_beams, nx, ny, nz, code, *fusion_args,
                out.data_ptr(), _capi.ptr(forces), stream))
        else:
            _capi.check(lib.ch_sc_field(
                ws.phi.data_ptr(), ws.params.data_ptr(), n_beams, nx, ny, nz, code,
                ws.field.data_ptr(), stream))
            if fuse_records is None and next_element is None:
                _capi.check(lib.ch_sc_gather_kick(
                    p.data_ptr(), p_stride, ws.field.data_ptr(), _capi.SC_FIELD_NODES,
                    ws.params.data_ptr(), n, n_beams, nx, ny, nz, code, out.data_ptr(),
                    _capi.ptr(forces), stream))
            else:
                _capi.check(lib.ch_sc_gather_kick_fused(
                    p.data_ptr(), p_stride, ws.field.data_ptr(), _capi.SC_FIELD_NODES,
                    ws.params.data_ptr(), n, n_beams, nx, ny, nz, code, *fusion_args,
                    out.data_ptr(), stream))
        if nxt is not None:
            ws.prepared_next = Prepared(ws, next_slot, next_element)
    ws.forces = forces
    return out.reshape(*vector_shape, n, 7), ws


class KickTensors:
    """The tensors of one SpaceChargeKick element, read once per lowered program (every attribute
    of an element is an ``nn.Module.__getattr__`` call, and a kick needs them a dozen times; a
    re-assigned attribute re-lowers the lattice, so the lowered stage may hold them like the slot
    table holds the parameters of linear elements)."""

    __slots__ = ("effect_length", "extents", "grid_shape", "shapes")

    def __init__(self, element) -> None:
        self.effect_length = element.effect_length
        self.extents = (element.grid_extent_x, element.grid_extent_y, element.grid_extent_tau)
        self.grid_shape = tuple(int(v) for v in element.grid_shape)
        self.shapes = (self.effect_length.shape, *(e.shape for e in self.extents))


def kick_vector_shape(element, incoming, tensors: KickTensors | None = None) -> tuple:
    """Vector shape one kick works on (space_charge_kick.py:493-528, without the (1,) helper)."""
    tensors = tensors or KickTensors(element)
    return tuple(_bshape(
        incoming.particles.shape[:-2], incoming.energy.shape,
        incoming.particle_charges.shape[:-1], incoming.survival_probabilities.shape[:-1],
        *tensors.shapes,
    ))


def track(element, incoming):
    """``SpaceChargeKick.track`` (space_charge_kick.py:477-586)."""
    return track_fused(element, incoming)[0]


def track_fused(element, incoming, prepared: Prepared | None = None, fuse_records=None,
                next_element=None, records_ready=None, tensors: KickTensors | None = None,
                next_tensors: KickTensors | None = None, s=None, species=None):
    """``SpaceChargeKick.track`` plus, optionally, the linear map of the following section and
    the moments of the following kick in the same particle pass.  Returns (beam, Prepared|None);
    the caller adds the section length to ``s`` when it passed ``fuse_records``."""
    assert type(incoming).__name__ == "ParticleBeam", (
        "SpaceChargeKick tracking is currently only supported for `ParticleBeam`."
    )
    particles = incoming.particles
    if particles.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"cheetah_b200 tracks float32/float64 beams, got {particles.dtype}")
    tensors = tensors or KickTensors(element)
    names = ("effect_length", "grid_extent_x", "grid_extent_y", "grid_extent_tau")
    for name, tensor in zip(names, (tensors.effect_length, *tensors.extents)):
        if tensor.device != particles.device:
            raise ValueError(
                f"{name} of element {element.name!r} lives on {tensor.device} but the beam is on "
                f"{particles.device}; move the lattice with `segment.to(device)` first"
            )
    if next_element is not None and next_tensors is None:
        next_tensors = KickTensors(next_element)
    out, ws = kick(
        particles, incoming.energy, incoming.particle_charges, incoming.survival_probabilities,
        incoming.species.mass_eV, tensors.effect_length, tensors.extents,
        tensors.grid_shape, prepared=prepared, element=element, fuse_records=fuse_records,
        next_element=next_element, records_ready=records_ready, next_tensors=next_tensors,
    )
    # the reference drops the (1,) helper dimension again when nothing is vectorised
    out_shape = _bshape(
        particles.shape[:-2], incoming.energy.shape, incoming.particle_charges.shape[:-1],
        incoming.survival_probabilities.shape[:-1], *tensors.shapes,
    )
    out = out.reshape(*out_shape, particles.shape[-2], 7)
    # (`s`, `species`: the caller fused a linear section into the kick and passes the path length
    # after it, so that the outgoing beam is built once)
    outgoing = _new_beam(
        incoming, out, incoming.energy, incoming.particle_charges,
        incoming.survival_probabilities, incoming.s if s is None else s,
        incoming.species if species is None else species,
        getattr(incoming, "_unit_seventh", None),
    )
    return outgoing, ws.prepared_next


def cloud_in_cell_charge_deposition(positions, bins, extent=None, charges=None):
    """Cloud-in-cell deposit with the signature of cheetah/utils/cloud_in_cell.py:8-64 for 1, 2
    or 3 position dimensions: ``positions (..., N, d)`` -> ``(..., *bins)``."""
    dims = positions.shape[-1]
    if dims not in (1, 2, 3):
        raise NotImplementedError(
            "cheetah_b200 accelerates cloud-in-cell deposits in 1, 2 and 3 dimensions"
        )
    if not positions.is_cuda:
        raise RuntimeError("cheetah_b200: positions must be on a CUDA device (no CPU fallback)")
    dtype, device = positions.dtype, positions.device
    shape = [bins] * dims if isinstance(bins, int) else list(bins)
    assert len(shape) == dims, "Number of bin values must match number of position dimensions."
    if extent is None:
        extent = torch.stack([positions.amin(dim=-2), positions.amax(dim=-2)], dim=-1)
    vector_shape = tuple(positions.shape[:-2])
    n = positions.shape[-2]
    n_beams = max(1, math.prod(vector_shape))
    pos = positions.reshape(n_beams, n, dims).contiguous()
    ext = extent.to(dtype).expand(*vector_shape, dims, 2).reshape(n_beams, dims, 2).contiguous()
    q = None
    if charges is not None:
        q = charges.to(dtype).expand(*vector_shape, n).reshape(n_beams, n).contiguous()
    grid = torch.empty((n_beams, *shape), dtype=dtype, device=device)
    padded = shape + [1] * (3 - dims)
    with _capi.device_guard(device):
        if torch.are_deterministic_algorithms_enabled():
            scratch = torch.empty((n_beams * (math.prod(shape) + 1),), dtype=torch.int64,
                                  device=device)
            _capi.check(_capi.lib().ch_cic_deposit_deterministic(
                pos.data_ptr(), ext.data_ptr(), _capi.ptr(q), n, n_beams, dims, *padded,
                _capi.dtype_code(dtype), scratch.data_ptr(), grid.data_ptr(),
                _capi.current_stream(device)))
        else:
            _capi.check(_capi.lib().ch_cic_deposit(
                pos.data_ptr(), ext.data_ptr(), _capi.ptr(q), n, n_beams, dims, *padded,
                _capi.dtype_code(dtype), grid.data_ptr(), _capi.current_stream(device)))
    return grid.reshape(*vector_shape, *shape)
