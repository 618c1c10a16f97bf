"""Host-buffer entry points: track a CPU-resident beam/lattice on the GPU and stream the
results back to host memory.

This is the call a user of the reference makes when everything lives in host memory (the
reference's own mode of operation): inputs are CPU tensors, outputs land in host buffers.

``HostTracker`` (lattices that lower to ONE linear section -- drifts, magnets, apertures: the
ARES case -- with many vectorised settings): every call uploads the lattice settings and the
beam (H2D), runs the CUDA path in chunks of settings, and downloads the outgoing beam (D2H)
through a ring of pinned staging buffers, the copy of chunk c overlapping the kernels of chunk
c+1 on a second stream.  The path is PCIe-bound, so the bytes are what matters: the seventh
phase-space column is the constant 1 (particle_beam.py:60-106) and never crosses the bus, and
when the incoming survival probabilities are all ones the outgoing ones are 0/1 masks and travel
as one byte each -- 25 instead of 32 bytes per (particle, setting) (``ch_apply_maps_compact``).
Outputs are delivered chunk by chunk to ``consumer(begin, end, coordinates_host,
survival_host)`` with ``coordinates_host (count, N, 6)`` and ``survival_host (count, N)`` (uint8
mask or beam dtype); the pinned buffers are a ring that is reused, so a consumer must finish
with (or copy) its views before returning.

``track_host`` is the general entry point for any lattice the package tracks (space-charge
kicks, non-linear runs, several sections): host tensors in, ``Segment.track`` on the device,
host tensors out.
"""

from __future__ import annotations

import math

import torch

from . import _capi, lowering, tracking


def _pinned_like(shape, dtype) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype).pin_memory()


class _NumaLocal:
    """While active, the calling thread runs on the CPUs NVML reports as closest to ``device``, so
    that pinned staging buffers allocated inside are first-touched on the GPU's NUMA node (the
    D2H stream is PCIe-bound; a remote node adds an inter-socket hop).  Best effort: without
    NVML, or on a single-node host, it does nothing."""

    def __init__(self, device: torch.device) -> None:
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.previous = None

    def __enter__(self):
        import os

        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            physical = int(visible.split(",")[self.index]) if visible else self.index
            handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
            self.previous = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(handle)
        except Exception:
            self.previous = None
        return self

    def __exit__(self, *exc) -> None:
        import os

        if self.previous is not None:
            try:
                os.sched_setaffinity(0, self.previous)
            except OSError:
                pass


class HostTracker:
    def __init__(self, segment_cpu, n_particles: int, n_settings: int, device="cuda",
                 dtype=torch.float32, chunk_settings: int = 64, ring: int = 4,
                 consumer_threads: int = 2) -> None:
        import copy

        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        self.n_particles = n_particles
        self.n_settings = n_settings
        self.chunk = min(chunk_settings, n_settings)
        self.host_segment = segment_cpu
        self.device_segment = copy.deepcopy(segment_cpu).to(self.device)
        # Lattice tensors travel as ONE flat pinned buffer -> ONE flat device buffer per call: the
        # device lattice's buffers are re-pointed to 16-byte aligned views of the device block
        # (ARES has ~400 small tensors; one copy each costs more than the whole kernel).  The host
        # lattice's CURRENT tensors are looked up by name at every call (`quad.k1 =
        # torch.tensor(...)` rebinds a buffer) and only re-staged when (data_ptr, _version) changed.
        self.layout = []  # (name, offset, nbytes, shape, dtype)
        offset = 0
        for name, tensor in list(self.device_segment.named_buffers()) + list(
                self.device_segment.named_parameters()):
            nbytes = tensor.numel() * tensor.element_size()
            self.layout.append((name, offset, nbytes, tuple(tensor.shape), tensor.dtype))
            offset += (nbytes + 15) // 16 * 16
        self.flat_dev = torch.zeros(max(offset, 16), dtype=torch.uint8, device=self.device)
        self.flat_pinned = torch.zeros(max(offset, 16), dtype=torch.uint8).pin_memory()
        for name, off, nbytes, shape, tdtype in self.layout:
            if nbytes == 0:
                continue
            view = self.flat_dev[off:off + nbytes].view(tdtype).reshape(shape)
            view.copy_(self.device_segment.get_parameter(name) if name in dict(
                self.device_segment.named_parameters()) else self.device_segment.get_buffer(name))
            path, _, attr = name.rpartition(".")
            owner = self.device_segment.get_submodule(path) if path else self.device_segment
            if attr in owner._buffers:
                owner._buffers[attr] = view
            else:
                owner._parameters[attr] = torch.nn.Parameter(view, requires_grad=False)
        self.staged_versions = {}
        # Tensors whose VALUES shape the lowering (a cavity is an active element only when its
        # voltage is non-zero, lowering.LatticeProgram.watched) must not share the version counter
        # of the flat block -- every upload would look like an edit and force a re-lowering
        # (14 ms of Python for ARES).  They keep storage of their own and are copied only when the
        # host tensor changed.
        self.separate = {}
        program = tracking._plan(list(self.device_segment.elements), self.device, (),
                                 self.device_segment)
        watched = {t.data_ptr() for t, _ in program.watched}
        for name, off, nbytes, shape, tdtype in self.layout:
            if nbytes and self.flat_dev.data_ptr() + off in watched:
                own = self.flat_dev[off:off + nbytes].view(tdtype).reshape(shape).clone()
                path, _, attr = name.rpartition(".")
                owner = self.device_segment.get_submodule(path) if path else self.device_segment
                if attr in owner._buffers:
                    owner._buffers[attr] = own
                else:
                    owner._parameters[attr] = torch.nn.Parameter(own, requires_grad=False)
                self.separate[name] = own
        # the buffers were re-pointed behind Element.__setattr__: drop the cached plan
        object.__setattr__(self.device_segment, "_plan_cache", None)
        n, c = n_particles, self.chunk
        self.beam_dev = torch.empty((n, 7), dtype=dtype, device=self.device)
        self.survival_dev = torch.empty((n,), dtype=dtype, device=self.device)
        self.energy_dev = torch.empty((), dtype=dtype, device=self.device)
        self.out_dev = [torch.empty((c, n, 6), dtype=dtype, device=self.device) for _ in range(2)]
        self.mask_dev = [torch.empty((c, n), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.surv_dev = None  # (c, n) beam dtype, allocated when a beam needs it
        with _NumaLocal(self.device):
            self.beam_pinned = _pinned_like((n, 7), dtype)
            self.survival_pinned = _pinned_like((n,), dtype)
            self.out_host = [_pinned_like((c, n, 6), dtype) for _ in range(ring)]
            self.mask_host = [_pinned_like((c, n), torch.uint8) for _ in range(ring)]
        self.surv_host = None
        self.ring = ring
        self.copy_stream = torch.cuda.Stream(self.device)
        from concurrent.futures import ThreadPoolExecutor

        self.workers = ThreadPoolExecutor(max_workers=consumer_threads)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.bytes_per_particle_setting = 6 * self.beam_dev.element_size() + 1

    # ---- uploads -------------------------------------------------------------------------
    def _host_tensors(self) -> dict:
        tensors = dict(self.host_segment.named_buffers())
        tensors.update(dict(self.host_segment.named_parameters()))
        return tensors

    def _upload(self, beam_cpu) -> bool:
        """H2D of the beam and of every lattice tensor; returns whether the incoming survival
        probabilities are all ones (then the outgoing ones are exact 0/1 masks)."""
        n = self.n_particles
        key = (beam_cpu.particles.data_ptr(), beam_cpu.particles._version)
        if getattr(self, "_checked_beam", None) != key:  # once per beam tensor version
            if not bool((beam_cpu.particles[..., 6] == 1).all()):
                raise ValueError("HostTracker needs particles[..., 6] == 1 (the reference's layout)")
            self._checked_beam = key
        survival = beam_cpu.survival_probabilities
        unit_survival = bool((survival == 1).all())
        # a beam that already lives in pinned memory is uploaded from where it is
        source = beam_cpu.particles
        if not (source.is_pinned() and source.is_contiguous() and source.dtype == self.dtype):
            self.beam_pinned.copy_(source)
            source = self.beam_pinned
        self.beam_dev.copy_(source, non_blocking=True)
        source = survival
        if not (source.is_pinned() and source.shape == (n,) and source.dtype == self.dtype):
            self.survival_pinned.copy_(survival.expand(n))
            source = self.survival_pinned
        self.survival_dev.copy_(source, non_blocking=True)
        self.energy_dev.copy_(beam_cpu.energy.to(self.dtype), non_blocking=True)
        self.h2d_bytes += (self.beam_dev.numel() + self.survival_dev.numel() + 1) \
            * self.beam_dev.element_size()
        host = self._host_tensors()
        for name, off, nbytes, shape, tdtype in self.layout:
            src = host[name]
            if nbytes == 0:
                continue
            if tuple(src.shape) != shape:
                raise ValueError(
                    f"lattice tensor {name!r} changed shape on the host ({shape} -> "
                    f"{tuple(src.shape)}); build a new HostTracker"
                )
            key = (src.data_ptr(), src._version)
            if self.staged_versions.get(name) == key:
                continue
            self.staged_versions[name] = key
            if name in self.separate:  # changed on the host: the plan sees the new version
                self.separate[name].copy_(src.detach(), non_blocking=False)
                self.h2d_bytes += nbytes
                continue
            self.flat_pinned[off:off + nbytes].view(tdtype).reshape(shape).copy_(src.detach())
        # the lattice block crosses the bus every call (the host may have changed any of it)
        self.flat_dev.copy_(self.flat_pinned, non_blocking=True)
        self.h2d_bytes += self.flat_dev.numel()
        return unit_survival

    def _device_beam(self, beam_cpu):
        from .beam import ParticleBeam

        beam = ParticleBeam(
            self.beam_dev, self.energy_dev, particle_charges=None,
            survival_probabilities=self.survival_dev,
            species=self._species(beam_cpu),
        )
        beam._unit_seventh = True
        return beam

    def _species(self, beam_cpu):
        species = beam_cpu.species
        if species.name in species.known:
            return species.__class__(species.name, device=self.device, dtype=self.dtype)
        return species.__class__(
            species.name, num_elementary_charges=species.num_elementary_charges.to(self.device),
            mass_eV=species.mass_eV.to(self.device))

    # ---- observables only --------------------------------------------------------------
    def track_moments(self, beam_cpu):
        """Host in, host out, but only the outgoing-beam moments come back: uploads beam and
        settings, runs the fused-epilogue kernel over all settings and downloads
        ``n_settings x 13`` numbers (``tracking.BeamMoments`` on the CPU)."""
        device = self.device
        self.h2d_bytes = self.d2h_bytes = 0
        self._upload(beam_cpu)
        beam = self._device_beam(beam_cpu)
        observed = self.device_segment.track_moments(beam)
        packed = torch.cat(
            [observed.mu, observed.sigma, observed.num_particles_survived.unsqueeze(-1)], dim=-1)
        if getattr(self, "moments_pinned", None) is None or \
                self.moments_pinned.shape != packed.shape:
            self.moments_pinned = _pinned_like(packed.shape, packed.dtype)
        self.moments_pinned.copy_(packed, non_blocking=True)
        self.d2h_bytes += packed.numel() * packed.element_size()
        torch.cuda.current_stream(device).synchronize()
        host = self.moments_pinned
        return tracking.BeamMoments(host[..., :6], host[..., 6:12], host[..., 12],
                                    beam_cpu.energy, observed.s.cpu())

    # ---- all outgoing particles ---------------------------------------------------------
    def track(self, beam_cpu, consumer=None) -> None:
        device, dtype, n = self.device, self.dtype, self.n_particles
        compute = torch.cuda.current_stream(device)
        self.h2d_bytes = self.d2h_bytes = 0
        unit_survival = self._upload(beam_cpu)
        if not unit_survival and self.surv_dev is None:
            c = self.chunk
            self.surv_dev = [torch.empty((c, n), dtype=dtype, device=device) for _ in range(2)]
            self.surv_host = [_pinned_like((c, n), dtype) for _ in range(self.ring)]

        program = tracking._plan(list(self.device_segment.elements), device, (), self.device_segment)
        sections = program.stages
        if len(sections) != 1 or not isinstance(sections[0], lowering.LinearSection) \
                or sections[0].cavity is not None:
            raise NotImplementedError(
                "HostTracker handles lattices that lower to one linear section without an active "
                "cavity; use cheetah_b200.host.track_host for anything else")
        section = sections[0]
        species = self._species(beam_cpu)
        records, vm = tracking._compose(program, section, self.energy_dev, species, dtype)
        n_settings = math.prod(vm)
        assert n_settings == self.n_settings, (n_settings, self.n_settings)
        rec_len = records.shape[1]
        lib = _capi.lib()
        element_size = self.beam_dev.element_size()

        done = [None, None]  # D2H-finished events per device buffer
        slot_busy = [None] * self.ring  # future of the consumer still reading a ring slot
        for index, begin in enumerate(range(0, n_settings, self.chunk)):
            end = min(begin + self.chunk, n_settings)
            count = end - begin
            buf = index & 1
            if done[buf] is not None:
                compute.wait_event(done[buf])  # previous download of this device buffer
            out = self.out_dev[buf]
            mask = self.mask_dev[buf] if unit_survival else None
            surv = None if unit_survival else self.surv_dev[buf]
            with _capi.device_guard(device):
                _capi.check(
                    lib.ch_apply_maps_compact(
                        self.beam_dev.data_ptr(), 0, None,
                        self.survival_dev.data_ptr(), 0, None,
                        records.data_ptr() + begin * rec_len * records.element_size(), rec_len,
                        None, rec_len, section.n_apertures, section.elliptical_mask,
                        n, count, out.data_ptr(), _capi.ptr(surv), _capi.ptr(mask),
                        _capi.dtype_code(dtype), compute.cuda_stream,
                    )
                )
            ready = torch.cuda.Event()
            ready.record(compute)
            slot = index % self.ring
            # the host consumer must be done with this ring slot before it is overwritten
            if slot_busy[slot] is not None:
                slot_busy[slot].result()
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                self.out_host[slot][:count].copy_(out[:count], non_blocking=True)
                if unit_survival:
                    self.mask_host[slot][:count].copy_(mask[:count], non_blocking=True)
                else:
                    self.surv_host[slot][:count].copy_(surv[:count], non_blocking=True)
                finished = torch.cuda.Event()
                finished.record(self.copy_stream)
            done[buf] = finished
            self.d2h_bytes += count * n * (6 * element_size + (1 if unit_survival else element_size))
            # consumers run on worker threads (event wait + host reads release the GIL), so the
            # GPU and the PCIe link keep working on the following chunks meanwhile
            slot_busy[slot] = self.workers.submit(
                self._deliver, (finished, begin, end, slot, unit_survival), consumer)
        for future in slot_busy:
            if future is not None:
                future.result()
        compute.wait_stream(self.copy_stream)

    def _deliver(self, item, consumer) -> None:
        finished, begin, end, slot, unit_survival = item
        finished.synchronize()
        if consumer is None:
            return
        count = end - begin
        survival = (self.mask_host if unit_survival else self.surv_host)[slot][:count]
        consumer(begin, end, self.out_host[slot][:count], survival)


def track_host(segment_device, particles, energy, particle_charges=None,
               survival_probabilities=None, device="cuda", buffers: dict | None = None):
    """Host tensors in, host tensors out, for ANY lattice ``Segment.track`` handles (space-charge
    kicks, non-linear runs, several sections, one or many beams).

    ``particles (..., N, 7)``, ``particle_charges (..., N)`` and ``survival_probabilities (..., N)``
    are CPU tensors; they are staged through pinned memory, tracked on ``device`` by
    ``segment_device`` (a Segment whose tensors live there) and the outgoing particles and
    survival probabilities come back in pinned host tensors.  ``buffers`` (returned by a previous
    call) reuses the pinned staging and device input buffers.  Returns
    ``(particles_host, survival_host, buffers)``."""
    from .beam import ParticleBeam, Species

    device = torch.device(device)
    dtype = particles.dtype
    if buffers is None or buffers["particles_in"].shape != particles.shape:
        buffers = {
            "particles_in": _pinned_like(particles.shape, dtype),
            "particles_dev": torch.empty(particles.shape, dtype=dtype, device=device),
            "particles_out": None,
        }
    buffers["particles_in"].copy_(particles)
    buffers["particles_dev"].copy_(buffers["particles_in"], non_blocking=True)

    def staged(name, tensor):
        if tensor is None:
            return None
        key = name + "_in"
        if key not in buffers or buffers[key].shape != tensor.shape:
            buffers[key] = _pinned_like(tensor.shape, tensor.dtype)
            buffers[name + "_dev"] = torch.empty(tensor.shape, dtype=tensor.dtype, device=device)
        buffers[key].copy_(tensor)
        buffers[name + "_dev"].copy_(buffers[key], non_blocking=True)
        return buffers[name + "_dev"]

    beam = ParticleBeam(
        buffers["particles_dev"], torch.as_tensor(energy, dtype=dtype).to(device),
        particle_charges=staged("charges", particle_charges),
        survival_probabilities=staged("survival", survival_probabilities),
        species=Species("electron", device=device, dtype=dtype),
    )
    out = segment_device.track(beam)
    if buffers["particles_out"] is None or buffers["particles_out"].shape != out.particles.shape:
        buffers["particles_out"] = _pinned_like(out.particles.shape, dtype)
        buffers["survival_out"] = _pinned_like(out.survival_probabilities.shape,
                                               out.survival_probabilities.dtype)
    buffers["particles_out"].copy_(out.particles, non_blocking=True)
    buffers["survival_out"].copy_(out.survival_probabilities, non_blocking=True)
    torch.cuda.current_stream(device).synchronize()
    return buffers["particles_out"], buffers["survival_out"], buffers
