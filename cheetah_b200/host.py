"""Host-buffer entry point: track a CPU-resident beam/lattice on the GPU and stream the
results back to host memory.

This is the call a user of the reference makes when everything lives in host memory (the
reference's own mode of operation): inputs are CPU tensors, outputs land in host buffers.
Every call uploads the lattice settings and the beam (H2D), runs the CUDA path in chunks of
settings, and downloads particles and survival probabilities (D2H) through pinned staging
buffers, with the copy of chunk c overlapping the kernels of chunk c+1 on a second stream.

``HostTracker`` supports lattices that lower to ONE linear section (drifts, magnets,
apertures: the ARES case).  Outputs are delivered chunk by chunk to ``consumer(begin, end,
particles_host, survival_host)``; the pinned buffers are a ring that is reused, so a consumer
must finish with (or copy) its views before returning.
"""

from __future__ import annotations

import math

import torch

from . import _capi, lowering, tracking


class HostTracker:
    def __init__(self, segment_cpu, n_particles: int, n_settings: int, device="cuda",
                 dtype=torch.float32, chunk_settings: int = 64, ring: int = 2) -> None:
        import copy

        self.device = torch.device(device)
        self.dtype = dtype
        self.n_particles = n_particles
        self.n_settings = n_settings
        self.chunk = min(chunk_settings, n_settings)
        self.host_segment = segment_cpu
        self.device_segment = copy.deepcopy(segment_cpu).to(self.device)
        self.pairs = [
            (dst, src)
            for dst, src in zip(self.device_segment.buffers(), self.host_segment.buffers())
        ]
        self.host_settings = [src.pin_memory() for _, src in self.pairs]
        n, c = n_particles, self.chunk
        self.beam_dev = torch.empty((n, 7), dtype=dtype, device=self.device)
        self.survival_dev = torch.empty((n,), dtype=dtype, device=self.device)
        self.energy_dev = torch.empty((), dtype=dtype, device=self.device)
        self.beam_pinned = torch.empty((n, 7), dtype=dtype).pin_memory()
        self.survival_pinned = torch.empty((n,), dtype=dtype).pin_memory()
        self.out_dev = [torch.empty((c, n, 7), dtype=dtype, device=self.device) for _ in range(2)]
        self.surv_dev = [torch.empty((c, n), dtype=dtype, device=self.device) for _ in range(2)]
        self.out_host = [torch.empty((c, n, 7), dtype=dtype).pin_memory() for _ in range(ring)]
        self.surv_host = [torch.empty((c, n), dtype=dtype).pin_memory() for _ in range(ring)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def track_moments(self, beam_cpu):
        """Host in, host out, but only the outgoing-beam moments come back: uploads beam and
        settings, runs the fused-epilogue kernel over all settings and downloads
        ``n_settings x 20`` doubles (``tracking.BeamMoments`` on the CPU)."""
        device, dtype = self.device, self.dtype
        self.h2d_bytes = self.d2h_bytes = 0
        self._upload(beam_cpu)
        beam = self._device_beam(beam_cpu)
        observed = self.device_segment.track_moments(beam)
        if not hasattr(self, "moments_pinned"):
            self.moments_pinned = [
                torch.empty(t.shape, dtype=t.dtype).pin_memory()
                for t in (observed.mu, observed.sigma, observed.num_particles_survived)
            ]
        for dst, src in zip(
            self.moments_pinned, (observed.mu, observed.sigma, observed.num_particles_survived)
        ):
            dst.copy_(src, non_blocking=True)
            self.d2h_bytes += src.numel() * src.element_size()
        torch.cuda.current_stream(device).synchronize()
        return tracking.BeamMoments(*self.moments_pinned, beam_cpu.energy, observed.s.cpu())

    def _device_beam(self, beam_cpu):
        from .beam import ParticleBeam

        beam = ParticleBeam(
            self.beam_dev, self.energy_dev, particle_charges=None,
            survival_probabilities=self.survival_dev,
            species=beam_cpu.species.__class__(beam_cpu.species.name, device=self.device,
                                               dtype=self.dtype)
            if beam_cpu.species.name in beam_cpu.species.known else None,
        )
        beam._unit_seventh = True
        return beam

    def _upload(self, beam_cpu) -> None:
        n, dtype = self.n_particles, self.dtype
        self.beam_pinned.copy_(beam_cpu.particles)
        self.survival_pinned.copy_(beam_cpu.survival_probabilities.expand(n))
        self.beam_dev.copy_(self.beam_pinned, non_blocking=True)
        self.survival_dev.copy_(self.survival_pinned, non_blocking=True)
        self.energy_dev.copy_(beam_cpu.energy.to(dtype), non_blocking=True)
        self.h2d_bytes += self.beam_dev.numel() * 4 + self.survival_dev.numel() * 4 + 4
        for (dst, src), pinned in zip(self.pairs, self.host_settings):
            pinned.copy_(src)
            dst.copy_(pinned, non_blocking=True)  # in place: the lowered program stays valid
            self.h2d_bytes += dst.numel() * dst.element_size()

    def track(self, beam_cpu, consumer=None) -> None:
        device, dtype, n = self.device, self.dtype, self.n_particles
        compute = torch.cuda.current_stream(device)
        self.h2d_bytes = self.d2h_bytes = 0

        # ---- H2D: beam + every lattice parameter (settings may have changed on the host) ----
        self.beam_pinned.copy_(beam_cpu.particles)
        self.survival_pinned.copy_(beam_cpu.survival_probabilities.expand(n))
        self.beam_dev.copy_(self.beam_pinned, non_blocking=True)
        self.survival_dev.copy_(self.survival_pinned, non_blocking=True)
        self.energy_dev.copy_(beam_cpu.energy.to(dtype), non_blocking=True)
        self.h2d_bytes += self.beam_dev.numel() * 4 + self.survival_dev.numel() * 4 + 4
        for (dst, src), pinned in zip(self.pairs, self.host_settings):
            pinned.copy_(src)
            dst.copy_(pinned, non_blocking=True)  # in place: the lowered program stays valid
            self.h2d_bytes += dst.numel() * dst.element_size()

        program = tracking._plan(list(self.device_segment.elements), device, (), self.device_segment)
        sections = program.stages
        if len(sections) != 1 or not isinstance(sections[0], lowering.LinearSection):
            raise NotImplementedError("HostTracker handles lattices with one linear section")
        section = sections[0]
        species = beam_cpu.species.__class__(beam_cpu.species.name, device=device, dtype=dtype)
        records, vm = tracking._compose(program, section, self.energy_dev, species, dtype)
        n_settings = math.prod(vm)
        assert n_settings == self.n_settings, (n_settings, self.n_settings)
        rec_len = records.shape[1]
        lib = _capi.lib()

        done = [None, None]  # D2H-finished events per device buffer
        for index, begin in enumerate(range(0, n_settings, self.chunk)):
            end = min(begin + self.chunk, n_settings)
            count = end - begin
            buf = index & 1
            if done[buf] is not None:
                compute.wait_event(done[buf])  # previous download of this device buffer
            out, surv = self.out_dev[buf], self.surv_dev[buf]
            with _capi.device_guard(device):
                _capi.check(
                    lib.ch_apply_maps(
                        self.beam_dev.data_ptr(), 0, None,
                        self.survival_dev.data_ptr(), 0, None,
                        records.data_ptr() + begin * rec_len * records.element_size(), rec_len,
                        None, rec_len, section.n_apertures, section.elliptical_mask,
                        n, count, out.data_ptr(), surv.data_ptr(),
                        _capi.dtype_code(dtype), 1, compute.cuda_stream,
                    )
                )
            ready = torch.cuda.Event()
            ready.record(compute)
            slot = index % len(self.out_host)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                self.out_host[slot][:count].copy_(out[:count], non_blocking=True)
                self.surv_host[slot][:count].copy_(surv[:count], non_blocking=True)
                finished = torch.cuda.Event()
                finished.record(self.copy_stream)
            done[buf] = finished
            self.d2h_bytes += count * n * 32
            if consumer is not None:
                finished.synchronize()
                consumer(begin, end, self.out_host[slot][:count], self.surv_host[slot][:count])
        compute.wait_stream(self.copy_stream)
