"""``Segment.track`` / ``Element.track`` / ``first_order_transfer_map`` on the CUDA library.

Host-side mirror of cheetah/accelerator/segment.py:534-574 and element.py:159-193: beam
bookkeeping (``s``, ``energy``, ``species.clone()``, pass-through of untouched tensors,
NumPy-style broadcasting of vector dimensions) stays in Python; every floating-point
operation on maps and particles happens in ``libcheetah_b200.so``.  No CPU fallback: a beam
that is not on a CUDA device raises.
"""

from __future__ import annotations

import math
import warnings

import torch

from . import _capi, lowering


# When set to a list, every ch_apply_maps launch appends a (start, stop) CUDA-event pair
# recorded on the launching stream (bench.py uses it for the roofline line).
apply_events: list | None = None


def _bshape(*shapes) -> tuple:
    """``torch.broadcast_shapes`` for the common cases without its per-call overhead."""
    result = ()
    for shape in shapes:
        shape = tuple(shape)
        if shape == result or not shape:
            continue
        if not result:
            result = shape
            continue
        return tuple(torch.broadcast_shapes(*shapes))
    return result


def _new_beam(like, particles, energy, particle_charges, survival_probabilities, s, species,
              unit_seventh=None):
    """Outgoing beam of the same class as ``like`` (the reference's ParticleBeam or ours)."""
    cls = like.__class__
    fast = getattr(cls, "_from_tracking", None)
    if fast is not None:
        return fast(particles, energy, particle_charges, survival_probabilities, s, species,
                    unit_seventh)
    beam = cls(particles, energy, particle_charges=particle_charges,
               survival_probabilities=survival_probabilities, s=s, species=species)
    try:
        beam._unit_seventh = unit_seventh
    except Exception:
        pass
    return beam


def _is_particle_beam(beam) -> bool:
    return type(beam).__name__ == "ParticleBeam"


def _is_parameter_beam(beam) -> bool:
    return type(beam).__name__ == "ParameterBeam"


def _require_cuda(tensor: torch.Tensor, what: str) -> None:
    if not tensor.is_cuda:
        raise RuntimeError(
            f"cheetah_b200: {what} is on {tensor.device}. This backend only runs on CUDA devices "
            "(B200, sm_100a) and has no CPU fallback; move the beam and lattice to 'cuda'."
        )


def _plan(elements, device: torch.device, target_shape: tuple, cache_owner=None):
    """Lowered program for ``elements``, cached on ``cache_owner`` (a Segment).

    Fast test: nothing anywhere was edited (global epoch) and no value-dependent decision went
    stale.  When the epoch moved, the cached program is kept if THIS lattice still consists of the
    same element objects in the same order with unchanged edit counters (``lattice_signature``):
    attribute traffic on unrelated elements does not force a re-lowering, while removed, inserted
    or reordered elements do."""
    from .elements import lattice_epoch, lattice_signature

    epoch = lattice_epoch()
    key = (device, tuple(target_shape))
    cache = getattr(cache_owner, "_plan_cache", None) if cache_owner is not None else None
    if cache is not None and cache[0] == key and not cache[2].is_stale():
        if cache[1] == epoch:
            return cache[2]
        if cache[3] == lattice_signature(elements):
            object.__setattr__(cache_owner, "_plan_cache", (key, epoch, cache[2], cache[3]))
            return cache[2]
    program = lowering.lower(elements, device, target_shape)
    if cache_owner is not None and hasattr(cache_owner, "_plan_cache"):
        object.__setattr__(cache_owner, "_plan_cache",
                           (key, epoch, program, lattice_signature(elements)))
    return program


def _index_table(source_shape: tuple, out_shape: tuple, device) -> torch.Tensor | None:
    """int32 table mapping a flat index over ``out_shape`` to one over ``source_shape``
    (None when the mapping is the identity or a constant)."""
    if math.prod(source_shape) == 1 or tuple(source_shape) == tuple(out_shape):
        return None
    index = torch.arange(math.prod(source_shape), dtype=torch.int32, device=device)
    return index.reshape(source_shape).expand(out_shape).reshape(-1).contiguous()


def _compose(program, section, energy: torch.Tensor, species, dtype):
    """Run ``ch_compose_maps`` for one section -> records ``(*map_shape, record_len)``."""
    device = energy.device
    mass_eV, charge = species.mass_eV, species.num_elementary_charges
    if section.cavity is not None:
        # the reference switches to the energy-gain second-order terms when ANY setting of the
        # batch gains energy (cavity.py:155); evaluated on the device, no host sync
        element, gain_flag = section.cavity
        delta_energy = element.voltage * element.phase.deg2rad().cos() * charge * -1
        gain_flag.copy_((delta_energy > 0).any())
    map_shape = tuple(_bshape(section.lattice_shape, energy.shape))
    n_settings = math.prod(map_shape)
    if energy.dtype not in (torch.float32, torch.float64):
        energy = energy.to(dtype)
    if energy.numel() == 1:
        energy_stride = 0
    else:
        energy = energy.expand(map_shape).contiguous()
        energy_stride = 1
    rec_len = _capi.record_len(section.n_apertures, section.cavity is not None)
    records = torch.empty((n_settings, rec_len), dtype=dtype, device=device)
    with _capi.device_guard(device):
        _capi.check(
            _capi.lib().ch_compose_maps(
                program.native, section.op_begin, section.op_end, n_settings,
                energy.data_ptr(), energy_stride, _capi.dtype_code(energy.dtype),
                mass_eV.data_ptr(), _capi.dtype_code(mass_eV.dtype),
                charge.data_ptr(), _capi.dtype_code(charge.dtype),
                records.data_ptr(), rec_len, _capi.dtype_code(dtype),
                _capi.current_stream(device),
            )
        )
    return records, map_shape


def _section_length(records: torch.Tensor, map_shape: tuple, length_shape: tuple,
                    column: int = 1) -> torch.Tensor:
    """Sum of element lengths with the reference's (un-expanded) vector shape."""
    total = records[:, column].reshape(map_shape)
    # drop the vector dims the lengths do not carry
    lead = len(map_shape) - len(length_shape)
    index = [0] * lead + [slice(None) if n > 1 else 0 for n in length_shape]
    total = total[tuple(index)] if index else total
    return total.reshape(length_shape)


def _maps_from_records(records: torch.Tensor, map_shape: tuple) -> torch.Tensor:
    """(..., 7, 7) maps from compose records."""
    n = records.shape[0]
    tm = torch.zeros((n, 7, 7), dtype=records.dtype, device=records.device)
    tm[:, :6, :] = records[:, _capi.RECORD_HEADER : _capi.RECORD_HEADER + 42].reshape(n, 6, 7)
    tm[:, 6, 6] = 1.0
    return tm.reshape(*map_shape, 7, 7)


def first_order_transfer_map(elements, energy: torch.Tensor, species) -> torch.Tensor:
    """Merged first-order map of skippable ``elements`` (segment.py:534-541)."""
    _require_cuda(energy, "energy")
    program = _plan(elements, energy.device, tuple(energy.shape))
    sections = [s for s in program.stages if isinstance(s, lowering.LinearSection)]
    if len(program.stages) == 0:
        return torch.eye(7, dtype=energy.dtype, device=energy.device).repeat(*energy.shape, 1, 1)
    if len(sections) != 1 or len(program.stages) != 1 or sections[0].n_apertures:
        raise ValueError("first_order_transfer_map needs a run of skippable elements")
    dtype = energy.dtype if energy.dtype in (torch.float32, torch.float64) else torch.float32
    records, map_shape = _compose(program, sections[0], energy, species, dtype)
    return _maps_from_records(records, map_shape)


def _unit_seventh(beam) -> bool:
    flag = getattr(beam, "_unit_seventh", None)
    if flag is None:
        flag = bool((beam.particles[..., 6] == 1).all())
        try:
            beam._unit_seventh = flag
        except Exception:
            pass
    return flag


def _track_linear_section(program, section, beam, moments: str | None = None,
                          covariance: bool = False):
    """One ``ch_compose_maps`` + one ``ch_apply_maps`` for a ParticleBeam.

    ``moments``: None (particles only), "with" (particles + fused moments) or "only" (the
    outgoing particles are never written to HBM; returns ``(None, BeamMoments)``)."""
    particles = beam.particles
    device, dtype = particles.device, particles.dtype
    if dtype not in (torch.float32, torch.float64):
        raise TypeError(f"cheetah_b200 tracks float32/float64 beams, got {dtype}")
    n = particles.shape[-2]
    vp = tuple(particles.shape[:-2])

    records, vm = _compose(program, section, beam.energy, beam.species, dtype)
    new_s = beam.s + _section_length(records, vm, section.length_shape)
    new_energy = beam.energy
    if section.cavity is not None:
        cavity = section.cavity[0]
        new_energy = beam.energy + (
            cavity.voltage * cavity.phase.deg2rad().cos()
            * beam.species.num_elementary_charges * -1
        ).to(beam.energy.dtype)

    if not section.has_maps and section.n_apertures == 0 and moments is None:
        # identity elements only: the reference still multiplies by eye(7).repeat(energy.shape)
        widened = tuple(_bshape(vp, beam.energy.shape))
        if widened != vp:
            particles = particles.expand(*widened, n, 7)
        return _new_beam(beam, particles, beam.energy, beam.particle_charges,
                         beam.survival_probabilities, new_s, beam.species.clone(),
                         getattr(beam, "_unit_seventh", None))

    vo = tuple(_bshape(vm, vp))
    vs = tuple(beam.survival_probabilities.shape[:-1])
    if section.n_apertures or moments is not None:
        # survival probabilities may carry vector dims of their own (tests/test_vectorized.py:
        # 339-371): the kernel then runs on the wider batch and the particles are narrowed again
        vo = tuple(_bshape(vo, vs))
    n_out = math.prod(vo)
    if not particles.is_contiguous():
        particles = particles.contiguous()
    particle_index = _index_table(vp, vo, device)
    record_index = _index_table(vm, vo, device)
    out = None if moments == "only" else torch.empty((*vo, n, 7), dtype=dtype, device=device)
    sums = None
    n_sums = _capi.MOMENTS_COV if covariance else _capi.MOMENTS
    if moments is not None:
        sums = torch.empty((n_out, n_sums), dtype=torch.float64, device=device)

    survival_in = beam.survival_probabilities
    survival_out = None
    survival_index = None
    if section.n_apertures:
        if survival_in.dtype != dtype or not survival_in.is_contiguous():
            survival_in = survival_in.to(dtype).contiguous()
        # the kernel works on the full output batch; the reference's survival tensor only
        # carries the vector dims that reached the last aperture
        survival_index = _index_table(vs, vo, device)
        if moments != "only":
            survival_out = torch.empty((*vo, n), dtype=dtype, device=device)

    events = None
    if apply_events is not None:
        events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        events[0].record(torch.cuda.current_stream(device))
    if moments is not None and not section.n_apertures:
        # the sums are weighted with the survival probabilities even without apertures
        if survival_in.dtype != dtype or not survival_in.is_contiguous():
            survival_in = survival_in.to(dtype).contiguous()
        survival_index = _index_table(vs, vo, device)
    common = (
        particles.data_ptr(), 0 if math.prod(vp) == 1 else n * 7, _capi.ptr(particle_index),
        survival_in.data_ptr() if (section.n_apertures or moments is not None) else None,
        0 if math.prod(vs) == 1 else n, _capi.ptr(survival_index),
        records.data_ptr(), 0 if math.prod(vm) == 1 else records.shape[1],
        _capi.ptr(record_index),
        records.shape[1], section.n_apertures, section.elliptical_mask,
        n, n_out, _capi.ptr(out), _capi.ptr(survival_out),
    )
    tail = (_capi.dtype_code(dtype), int(_unit_seventh(beam)), _capi.current_stream(device))
    with _capi.device_guard(device):
        if moments is None:
            _capi.check(_capi.lib().ch_apply_maps(*common, *tail))
        elif covariance:
            _capi.check(_capi.lib().ch_apply_maps_covariance(*common, sums.data_ptr(), *tail))
        else:
            _capi.check(_capi.lib().ch_apply_maps_moments(*common, sums.data_ptr(), *tail))
    if events is not None:
        events[1].record(torch.cuda.current_stream(device))
        apply_events.append(events)

    observed = None
    if sums is not None:
        observed = BeamMoments.from_sums(sums.reshape(*vo, n_sums), new_energy, new_s, dtype)
    if moments == "only":
        return None, observed

    energy_shape = tuple(beam.energy.shape)
    if survival_out is not None:
        # reference shape: broadcast(survival_in, particles vector dims, everything up to and
        # including the last aperture); later (post-aperture) vectorised elements do not widen
        # it, and the beam energy only enters through a map in front of the aperture
        keep = tuple(_bshape(vs, vp, section.survival_shape,
                             energy_shape if section.maps_before_last_aperture else ()))
        new_survival = _narrowed(survival_out, vo, keep, (n,))
    else:
        new_survival = beam.survival_probabilities
    if out is not None:
        # particles: vectorised aperture limits do not widen them (aperture.py:108-132)
        keep = tuple(_bshape(vp, section.map_shape, energy_shape if section.n_map_ops else ()))
        out = _narrowed(out, vo, keep, (n, 7))

    outgoing = _new_beam(beam, out, new_energy, beam.particle_charges, new_survival, new_s,
                         beam.species.clone(), _unit_seventh(beam))
    return outgoing if moments is None else (outgoing, observed)


def _narrowed(tensor, vo: tuple, keep: tuple, inner: tuple):
    """The kernels work on the full broadcast batch ``vo``; the reference's tensors only carry
    the vector dims of what actually shaped them."""
    if keep == vo:
        return tensor
    lead = len(vo) - len(keep)
    index = [0] * lead + [slice(None) if k == v else 0 for k, v in zip(keep, vo[lead:])]
    return tensor[tuple(index)].reshape(*keep, *inner)


def _track_nonlinear_run(program, run, beam):
    """One ``ch_nonlinear_constants`` + one ``ch_track_nonlinear`` for a run of
    drift_kick_drift / second_order elements (drift.py:106-154, quadrupole.py:168-251,
    dipole.py:183-370, transverse_deflecting_cavity.py:122-209, element.py:195-225)."""
    particles = beam.particles
    device, dtype = particles.device, particles.dtype
    if dtype not in (torch.float32, torch.float64):
        raise TypeError(f"cheetah_b200 tracks float32/float64 beams, got {dtype}")
    n = particles.shape[-2]
    vp = tuple(particles.shape[:-2])
    lib = _capi.lib()
    species = beam.species
    energy = beam.energy
    vm = tuple(_bshape(run.lattice_shape, energy.shape))
    n_settings = math.prod(vm)
    if energy.dtype not in (torch.float32, torch.float64):
        energy = energy.to(dtype)
    if energy.numel() == 1:
        energy_stride = 0
    else:
        energy = energy.expand(vm).contiguous()
        energy_stride = 1
    with _capi.device_guard(device):
        n_consts = int(lib.ch_nonlinear_constants_len(program.native, run.op_begin, run.op_end))
        if n_consts < 0:
            _capi.check(-1)
        constants = torch.empty((n_settings, n_consts), dtype=torch.float64, device=device)
        mass_eV, charge = species.mass_eV, species.num_elementary_charges
        _capi.check(
            lib.ch_nonlinear_constants(
                program.native, run.op_begin, run.op_end, n_settings,
                energy.data_ptr(), energy_stride, _capi.dtype_code(energy.dtype),
                mass_eV.data_ptr(), _capi.dtype_code(mass_eV.dtype),
                charge.data_ptr(), _capi.dtype_code(charge.dtype),
                constants.data_ptr(), _capi.current_stream(device),
            )
        )
        new_s = beam.s + _section_length(constants, vm, run.length_shape, column=5).to(beam.s.dtype)
        vo = tuple(_bshape(vm, vp))
        if not particles.is_contiguous():
            particles = particles.contiguous()
        particle_index = _index_table(vp, vo, device)
        constants_index = _index_table(vm, vo, device)
        out = torch.empty((*vo, n, 7), dtype=dtype, device=device)
        _capi.check(
            lib.ch_track_nonlinear(
                program.native, run.op_begin, run.op_end,
                constants.data_ptr(), 0 if n_settings == 1 else n_consts,
                _capi.ptr(constants_index),
                particles.data_ptr(), 0 if math.prod(vp) == 1 else n * 7,
                _capi.ptr(particle_index),
                n, math.prod(vo), out.data_ptr(), _capi.dtype_code(dtype),
                _capi.current_stream(device),
            )
        )
    # the reference hands back ref_energy = sqrt(p0c^2 + m^2) == energy up to rounding, the same
    # species object and the untouched charges / survival probabilities
    return _new_beam(beam, out, beam.energy, beam.particle_charges, beam.survival_probabilities,
                     new_s, species, _unit_seventh(beam))


# Layout of a SECOND_ORDER block of the constants table (include/cheetah_b200.h,
# "second-order block"): frame, edge kicks, the nine distinct entries of the body R and the 39
# non-zero T_ijk in the order of track_methods.py:147-279.
_SO_COS, _SO_SIN, _SO_OX, _SO_OY, _SO_KX1, _SO_KY1, _SO_KX2, _SO_KY2, _SO_MX, _SO_MY = range(10)
_SO_R, _SO_T = 10, 19
_SO_R_PLACES = {0: [(0, 0), (1, 1)], 1: [(0, 1)], 2: [(1, 0)], 3: [(2, 2), (3, 3)], 4: [(2, 3)],
                5: [(3, 2)], 6: [(0, 5), (4, 1)], 7: [(1, 5), (4, 0)], 8: [(4, 5)]}
_SO_T_PLACES = [
    (0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 0, 5), (0, 1, 5), (0, 5, 5), (0, 2, 2), (0, 2, 3), (0, 3, 3),
    (1, 0, 0), (1, 0, 1), (1, 1, 1), (1, 0, 5), (1, 1, 5), (1, 5, 5), (1, 2, 2), (1, 2, 3), (1, 3, 3),
    (2, 0, 2), (2, 0, 3), (2, 1, 2), (2, 1, 3), (2, 2, 5), (2, 3, 5),
    (3, 0, 2), (3, 0, 3), (3, 1, 2), (3, 1, 3), (3, 2, 5), (3, 3, 5),
    (4, 0, 0), (4, 0, 1), (4, 1, 1), (4, 0, 5), (4, 1, 5), (4, 5, 5), (4, 2, 2), (4, 2, 3), (4, 3, 3),
]


def second_order_transfer_map(element, energy: torch.Tensor, species) -> torch.Tensor:
    """Dense ``T (..., 7, 7, 7)`` of ``out_i = sum_jk T_ijk in_j in_k`` with the first-order map in
    ``T[:, 6, :]`` and the element's frame changes folded in, as the reference returns it
    (drift.py:68-82, quadrupole.py:113-144, sextupole.py:91-116, dipole.py:397-428).

    The coefficients come from ``ch_nonlinear_constants`` (fp64, the same table the tracking
    kernel applies in sparse form); only the scatter into the dense tensor and the contraction
    with the 7 x 7 entry / exit frames -- a few hundred numbers per setting -- are torch ops."""
    if "second_order" not in element.supported_tracking_methods:
        raise NotImplementedError
    _require_cuda(energy, "energy")
    twin = element.clone()
    twin.tracking_method = "second_order"
    device, dtype = energy.device, element.length.dtype
    program = lowering.lower([twin], device, tuple(energy.shape))
    run = program.stages[0]
    assert isinstance(run, lowering.NonlinearRun)
    vm = tuple(_bshape(run.lattice_shape, energy.shape))
    n_settings = math.prod(vm)
    lib = _capi.lib()
    if energy.dtype not in (torch.float32, torch.float64):
        energy = energy.to(dtype)
    energy_c = energy if energy.numel() == 1 else energy.expand(vm).contiguous()
    with _capi.device_guard(device):
        n_consts = int(lib.ch_nonlinear_constants_len(program.native, run.op_begin, run.op_end))
        table = torch.empty((n_settings, n_consts), dtype=torch.float64, device=device)
        mass_eV, charge = species.mass_eV, species.num_elementary_charges
        _capi.check(lib.ch_nonlinear_constants(
            program.native, run.op_begin, run.op_end, n_settings, energy_c.data_ptr(),
            0 if energy_c.numel() == 1 else 1, _capi.dtype_code(energy_c.dtype),
            mass_eV.data_ptr(), _capi.dtype_code(mass_eV.dtype),
            charge.data_ptr(), _capi.dtype_code(charge.dtype),
            table.data_ptr(), _capi.current_stream(device),
        ))
    c = table[:, _capi.NL_HEADER:]
    eye = torch.eye(7, dtype=torch.float64, device=device).expand(n_settings, 7, 7)
    body = torch.zeros((n_settings, 7, 7, 7), dtype=torch.float64, device=device)
    for k, (i, j, l) in enumerate(_SO_T_PLACES):
        body[:, i, j, l] = c[:, _SO_T + k]
    first = eye.clone()
    for k, places in _SO_R_PLACES.items():
        for i, j in places:
            first[:, i, j] = c[:, _SO_R + k]
    body[:, :, 6, :] = first
    cs, sn = c[:, _SO_COS], c[:, _SO_SIN]
    entry, leave = eye.clone(), eye.clone()
    # entry: rotate into the element frame (+ offset), then the entrance edge kick
    entry[:, 0, 0], entry[:, 0, 2], entry[:, 0, 6] = cs, sn, c[:, _SO_OX]
    entry[:, 2, 0], entry[:, 2, 2], entry[:, 2, 6] = -sn, cs, c[:, _SO_OY]
    entry[:, 1, 1], entry[:, 1, 3] = cs, sn
    entry[:, 3, 1], entry[:, 3, 3] = -sn, cs
    entry[:, 1, :] = entry[:, 1, :] + c[:, _SO_KX1, None] * entry[:, 0, :]
    entry[:, 3, :] = entry[:, 3, :] + c[:, _SO_KY1, None] * entry[:, 2, :]
    # exit: the exit edge kick, then rotate back (+ offset)
    kick = eye.clone()
    kick[:, 1, 0], kick[:, 3, 2] = c[:, _SO_KX2], c[:, _SO_KY2]
    leave[:, 0, 0], leave[:, 0, 2], leave[:, 0, 6] = cs, -sn, c[:, _SO_MX]
    leave[:, 2, 0], leave[:, 2, 2], leave[:, 2, 6] = sn, cs, c[:, _SO_MY]
    leave[:, 1, 1], leave[:, 1, 3] = cs, -sn
    leave[:, 3, 1], leave[:, 3, 3] = sn, cs
    leave = leave @ kick
    dense = torch.einsum("bij,bjkl,bkn,blm->binm", leave, body, entry, entry)
    return dense.reshape(*vm, 7, 7, 7).to(dtype)


# Set to False to run every stage as its own pass (tests compare the two paths).
fuse_space_charge = True

_compose_streams: dict = {}


def _compose_stream(device) -> torch.cuda.Stream:
    stream = _compose_streams.get(device)
    if stream is None:
        stream = _compose_streams[device] = torch.cuda.Stream(device)
    return stream


def _track_space_charge(program, stages: list, i: int, beam, prepared):
    """SpaceChargeKick stage ``i`` with its neighbours fused into the gather pass when the lattice
    allows it: the following linear section (skippable run, no apertures / cavity, one map per
    beam) and the moments of the kick after that.  Returns (beam, prepared, stages consumed)."""
    from . import space_charge

    def kick_tensors(stage):  # read once per lowered program, see space_charge.KickTensors
        cached = stage.__dict__.get("_kick_tensors")
        if cached is None:
            cached = stage.__dict__["_kick_tensors"] = space_charge.KickTensors(stage.element)
        return cached

    element = stages[i].element
    tensors = kick_tensors(stages[i])
    section = records = next_element = next_tensors = None
    if fuse_space_charge and type(beam).__name__ == "ParticleBeam":
        vs = space_charge.kick_vector_shape(element, beam, tensors)
        j = i + 1
        if j < len(stages) and isinstance(stages[j], lowering.LinearSection):
            candidate = stages[j]
            vm = tuple(_bshape(candidate.lattice_shape, beam.energy.shape))
            if (candidate.has_maps and candidate.n_apertures == 0 and candidate.cavity is None
                    and (math.prod(vm) == 1 or vm == vs)):
                section = candidate
                j += 1
            else:
                j = len(stages)  # the next kick sees particles we do not produce here
        if j < len(stages) and isinstance(stages[j], lowering.Barrier) \
                and stages[j].kind == "space_charge":
            candidate = stages[j].element
            following = kick_tensors(stages[j])
            if (following.grid_shape == tensors.grid_shape
                    and tuple(_bshape(vs, *following.shapes)) == vs
                    and all(t.device == beam.particles.device for t in (
                        following.effect_length, *following.extents))):
                next_element, next_tensors = candidate, following
    records_ready = None
    if section is not None:
        # The maps of the following section only depend on the lattice and the beam energy: they
        # are composed on a side stream while the kick's deposit and Poisson solve run, and the
        # gather pass waits for them (one beam: 20 us of 128 per kick off the critical path).
        # Only while a CUDA graph is being captured (GraphedTrack): an eager call with one beam is
        # bound by the host's launch rate, where the extra events cost more than they save
        # (config 4 eager 17.9 -> 22.6 ms, graph replay 12.8 -> 12.4 ms), and with many beams the
        # 17 us of a compose launch do not matter.
        device = beam.particles.device
        if torch.cuda.is_current_stream_capturing():
            main = torch.cuda.current_stream(device)
            side = _compose_stream(device)
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                records, vm = _compose(program, section, beam.energy, beam.species,
                                       beam.particles.dtype)
                records_ready = torch.cuda.Event()
                records_ready.record(side)
            records.record_stream(main)
        else:
            records, vm = _compose(program, section, beam.energy, beam.species,
                                   beam.particles.dtype)
            # same stream as the kick: the outgoing beam is built once, with its final s
            new_s = beam.s + _section_length(records, vm, section.length_shape)
            outgoing, prepared = space_charge.track_fused(
                element, beam, prepared=prepared, fuse_records=records,
                next_element=next_element, tensors=tensors, next_tensors=next_tensors,
                s=new_s, species=beam.species.clone(),
            )
            return outgoing, prepared, 2
    outgoing, prepared = space_charge.track_fused(
        element, beam, prepared=prepared, fuse_records=records, next_element=next_element,
        records_ready=records_ready, tensors=tensors, next_tensors=next_tensors,
    )
    if section is not None:
        # (the kick made the main stream wait for the records)
        new_s = beam.s + _section_length(records, vm, section.length_shape)
        outgoing = _new_beam(
            outgoing, outgoing.particles, outgoing.energy, outgoing.particle_charges,
            outgoing.survival_probabilities, new_s, outgoing.species.clone(),
            getattr(outgoing, "_unit_seventh", None),
        )
    return outgoing, prepared, 1 if section is None else 2


def _track_monitor(stage, beam):
    """Active BPM / Screen between two sections (bpm.py:77-86, screen.py:187-239)."""
    from . import diagnostics

    if stage.kind == "bpm":
        return diagnostics.track_bpm(stage.element, beam)
    return diagnostics.track_screen(stage.element, beam)


class BeamMoments:
    """What ``ParticleBeam.mu_*`` / ``sigma_*`` / ``num_particles_survived`` return on the
    outgoing beam, computed in the epilogue of the apply kernel (no (B, N, 7) array needed).

    ``mu``, ``sigma``: ``(..., 6)`` for (x, px, y, py, tau, p); survival-weighted mean and
    unbiased weighted standard deviation (cheetah/utils/statistics.py:30-62)."""

    names = ("x", "px", "y", "py", "tau", "p")

    def __init__(self, mu, sigma, num_particles_survived, energy, s, cov=None) -> None:
        self.mu, self.sigma = mu, sigma
        self.num_particles_survived = num_particles_survived
        self.energy, self.s = energy, s
        # (..., 6, 6) unbiased weighted covariance matrix (statistics.py:65-88) when requested
        self.cov = cov

    @classmethod
    def from_sums(cls, sums: torch.Tensor, energy, s, dtype) -> "BeamMoments":
        s0, sww = sums[..., 0], sums[..., 1]
        s1, s2, pilot = sums[..., 2:8], sums[..., 8:14], sums[..., 14:20]
        mu = pilot + s1 / s0.unsqueeze(-1)
        centred = s2 - s1.square() / s0.unsqueeze(-1)
        correction = (s0 - sww / s0).unsqueeze(-1)
        sigma = (centred.clamp_min(0.0) / correction).sqrt()
        cov = None
        if sums.shape[-1] == _capi.MOMENTS_COV:
            rows, cols = torch.triu_indices(6, 6, offset=1, device=sums.device)
            second = torch.diag_embed(s2)
            second[..., rows, cols] = sums[..., 20:35]
            second[..., cols, rows] = sums[..., 20:35]
            cov = (
                second - s1.unsqueeze(-1) * s1.unsqueeze(-2) / s0[..., None, None]
            ) / correction.unsqueeze(-1)
            cov = cov.to(dtype)
        return cls(mu.to(dtype), sigma.to(dtype), s0.to(dtype), energy, s, cov)

    def __getattr__(self, name: str):
        for prefix, source in (("mu_", "mu"), ("sigma_", "sigma")):
            if name.startswith(prefix) and name[len(prefix):] in self.names:
                return getattr(self, source)[..., self.names.index(name[len(prefix):])]
        raise AttributeError(name)

    def __repr__(self) -> str:
        return f"BeamMoments(mu={self.mu!r}, sigma={self.sigma!r})"


def track_moments(elements, incoming, cache_owner=None, keep_particles: bool = False,
                  covariance: bool = False):
    """Track ``incoming`` and return the first/second moments of the outgoing beam.

    The last linear section of the lattice computes them in the kernel epilogue; with
    ``keep_particles=False`` (default) its outgoing particles are never written to HBM, which
    removes the 32 B per (particle, setting) write and the 131 GB capacity wall of large
    vectorised scans.  Returns ``BeamMoments`` (or ``(beam, BeamMoments)``).
    """
    if not _is_particle_beam(incoming):
        raise TypeError(f"Parameter incoming is of invalid type {type(incoming)}")
    _require_cuda(incoming.particles, "ParticleBeam")
    program = _plan(elements, incoming.particles.device, tuple(incoming.energy.shape), cache_owner)
    stages = program.stages
    if not stages or not isinstance(stages[-1], lowering.LinearSection):
        raise NotImplementedError("track_moments needs a lattice that ends with a linear section")
    beam = incoming
    for stage in stages[:-1]:
        if isinstance(stage, lowering.LinearSection):
            beam = _track_linear_section(program, stage, beam)
        elif isinstance(stage, lowering.NonlinearRun):
            beam = _track_nonlinear_run(program, stage, beam)
        elif stage.kind == "space_charge":
            from . import space_charge

            beam = space_charge.track(stage.element, beam)
        elif stage.kind in ("bpm", "screen"):
            beam = _track_monitor(stage, beam)
        else:
            raise NotImplementedError(
                f"cheetah_b200: element {stage.element.name!r} is outside the accelerated hot path"
            )
    outgoing, observed = _track_linear_section(
        program, stages[-1], beam, moments="with" if keep_particles else "only",
        covariance=covariance,
    )
    return (outgoing, observed) if keep_particles else observed


class _IdentityLattice:
    """A one-marker lattice with its own plan cache: lets the moments kernel observe a beam as it
    is (``ParticleBeam.second_moments``)."""

    def __init__(self) -> None:
        from .elements import Marker

        self.elements = [Marker(name="beam_moments")]
        self._plan_cache = None


_identity_lattice: _IdentityLattice | None = None


def beam_moments(beam) -> BeamMoments:
    """Survival-weighted mean, sigma and full 6 x 6 covariance of ``beam`` itself: one pass of
    the covariance kernel with the identity map (``ParticleBeam.mu_* / sigma_* / cov_*``,
    particle_beam.py:1699-1943; ``unbiased_weighted_covariance_matrix``, statistics.py:65-88)."""
    global _identity_lattice
    if _identity_lattice is None:
        _identity_lattice = _IdentityLattice()
    return track_moments(_identity_lattice.elements, beam, cache_owner=_identity_lattice,
                         covariance=True)


def _track_parameter_beam(program, beam):
    """tm @ mu, tm @ cov @ tm^T with the composed maps (element.py:166-179)."""
    mu, cov, s, energy = beam.mu, beam.cov, beam.s, beam.energy
    for stage in program.stages:
        if isinstance(stage, lowering.NonlinearRun):
            if "drift_kick_drift" in stage.methods:
                raise AssertionError(
                    "Drift-kick-drift tracking is currently only supported for `ParticleBeam`."
                )
            raise AssertionError(
                "Second-order tracking is currently only supported for `ParticleBeam`."
            )
        if isinstance(stage, lowering.Barrier) and stage.kind in ("bpm", "screen"):
            beam = _track_monitor(stage, beam.__class__(
                mu, cov, energy, total_charge=beam.total_charge, s=s, species=beam.species,
            ))
            mu, cov, s = beam.mu, beam.cov, beam.s
            continue
        if isinstance(stage, lowering.Barrier):
            if stage.kind == "space_charge":
                raise AssertionError(
                    "SpaceChargeKick tracking is currently only supported for `ParticleBeam`."
                )
            raise NotImplementedError(
                f"cheetah_b200: element {stage.element.name!r} of type "
                f"{type(stage.element).__name__} is outside the accelerated hot path"
            )
        if stage.n_apertures:
            warnings.warn(
                "Aperture tracking is currently only supported for `ParticleBeam`.",
                _physics_warning(), stacklevel=3,
            )
        dtype = mu.dtype
        records, vm = _compose(program, stage, energy, beam.species, dtype)
        cavity_offset = -1
        if stage.cavity is not None:
            cavity_offset = _capi.record_len(stage.n_apertures, False)
        # mu' = M mu, cov' = M cov M^T on the device library (element.py:166-179)
        device = mu.device
        vb = tuple(_bshape(mu.shape[:-1], cov.shape[:-2]))
        vo = tuple(_bshape(vm, vb))
        n_out = math.prod(vo)
        mu_c = mu.expand(*vb, 7).contiguous()
        cov_c = cov.expand(*vb, 7, 7).contiguous()
        beam_index = _index_table(vb, vo, device)
        record_index = _index_table(vm, vo, device)
        mu_out = torch.empty((*vo, 7), dtype=dtype, device=device)
        cov_out = torch.empty((*vo, 7, 7), dtype=dtype, device=device)
        with _capi.device_guard(device):
            _capi.check(_capi.lib().ch_apply_maps_parameter(
                mu_c.data_ptr(), 0 if math.prod(vb) == 1 else 7, _capi.ptr(beam_index),
                cov_c.data_ptr(), 0 if math.prod(vb) == 1 else 49,
                records.data_ptr(), 0 if math.prod(vm) == 1 else records.shape[1],
                _capi.ptr(record_index), cavity_offset, n_out, mu_out.data_ptr(),
                cov_out.data_ptr(), _capi.dtype_code(dtype), _capi.current_stream(device),
            ))
        # vectorised aperture limits do not shape a ParameterBeam (aperture.py:108-113)
        keep = tuple(_bshape(vb, stage.map_shape, tuple(energy.shape) if stage.n_map_ops else ()))
        mu, cov = _narrowed(mu_out, vo, keep, (7,)), _narrowed(cov_out, vo, keep, (7, 7))
        if stage.cavity is not None:
            cavity = stage.cavity[0]
            energy = energy + (
                cavity.voltage * cavity.phase.deg2rad().cos()
                * beam.species.num_elementary_charges * -1
            ).to(energy.dtype)
        s = s + _section_length(records, vm, stage.length_shape)
    return beam.__class__(
        mu, cov, energy, total_charge=beam.total_charge, s=s, species=beam.species.clone()
    )


def _physics_warning():
    from .elements import PhysicsWarning

    return PhysicsWarning


def track(elements, incoming, cache_owner=None):
    """Track ``incoming`` through ``elements`` (segment.py:545-574)."""
    if _is_parameter_beam(incoming):
        _require_cuda(incoming.mu, "ParameterBeam")
        program = _plan(elements, incoming.mu.device, tuple(incoming.energy.shape), cache_owner)
        return _track_parameter_beam(program, incoming)
    if not _is_particle_beam(incoming):
        raise TypeError(f"Parameter incoming is of invalid type {type(incoming)}")
    _require_cuda(incoming.particles, "ParticleBeam")

    program = _plan(elements, incoming.particles.device, tuple(incoming.energy.shape), cache_owner)
    return track_program(program, incoming)


def track_program(program, incoming):
    """Run an already lowered ``program`` on ``incoming`` (the stage loop of ``track``)."""
    if _is_parameter_beam(incoming):
        _require_cuda(incoming.mu, "ParameterBeam")
        return _track_parameter_beam(program, incoming)
    if not _is_particle_beam(incoming):
        raise TypeError(f"Parameter incoming is of invalid type {type(incoming)}")
    _require_cuda(incoming.particles, "ParticleBeam")
    beam = incoming
    stages = program.stages
    prepared = None  # grid parameters a kick already computed for the kick that follows it
    i = 0
    while i < len(stages):
        stage = stages[i]
        if isinstance(stage, lowering.Barrier) and stage.kind == "space_charge":
            beam, prepared, consumed = _track_space_charge(program, stages, i, beam, prepared)
            i += consumed
            continue
        prepared = None
        if isinstance(stage, lowering.LinearSection):
            beam = _track_linear_section(program, stage, beam)
        elif isinstance(stage, lowering.NonlinearRun):
            beam = _track_nonlinear_run(program, stage, beam)
        elif stage.kind in ("bpm", "screen"):
            beam = _track_monitor(stage, beam)
        else:
            raise NotImplementedError(
                f"cheetah_b200: element {stage.element.name!r} of type "
                f"{type(stage.element).__name__} (tracking_method="
                f"{getattr(stage.element, 'tracking_method', None)!r}) is outside the accelerated "
                "hot path of this build (SURVEY.md 8f)"
            )
        i += 1
    if beam is incoming:  # empty lattice: still hand back a new beam object
        beam = incoming.__class__(
            incoming.particles, incoming.energy, particle_charges=incoming.particle_charges,
            survival_probabilities=incoming.survival_probabilities, s=incoming.s,
            species=incoming.species.clone(),
        )
    return beam
