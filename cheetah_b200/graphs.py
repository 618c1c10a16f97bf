"""CUDA-graph replay of a whole ``Segment.track`` call.

A lattice with space-charge kicks is a chain of hundreds of short kernels (config 4: 100 kicks
x 15 launches + 101 linear sections); eager Python dispatch leaves the GPU idle between them.
``GraphedTrack`` captures one ``segment.track(beam)`` into a CUDA graph and replays it with a
single launch.  Semantics differ from the eager call in one documented way: the outgoing beam
lives in static buffers owned by the graph and is overwritten by the next ``replay``; the
incoming beam is read from a static input buffer that ``replay`` refreshes with a device copy.
Magnet settings are read from the live parameter tensors at replay time, so in-place updates
(`quad.k1.fill_(...)`, `copy_`) are honoured without re-capturing.
"""

from __future__ import annotations

import torch

from .elements import lattice_epoch, lattice_signature


class GraphedTrack:
    def __init__(self, segment, example_beam, warmup: int = 2) -> None:
        self.segment = segment
        self.static_in = example_beam.clone()
        self.static_in._unit_seventh = getattr(example_beam, "_unit_seventh", None)
        device = self.static_in.particles.device
        stream = torch.cuda.Stream(device)
        stream.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(stream):
            for _ in range(warmup):  # lowers the lattice, uploads the program, sizes the pools
                segment.track(self.static_in)
        torch.cuda.current_stream(device).wait_stream(stream)
        torch.cuda.synchronize(device)
        self.epoch = lattice_epoch()
        self.signature = lattice_signature([segment])
        self.unit_seventh = bool(getattr(self.static_in, "_unit_seventh", False))
        from .lowering import flatten

        self.screens = [e for e in flatten([segment]) if type(e).__name__ == "Screen"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = segment.track(self.static_in)
        if lattice_signature([segment]) != self.signature:
            raise RuntimeError("the lattice changed while it was being captured")
        self.epoch = lattice_epoch()

    def replay(self, beam=None):
        """Track ``beam`` (or the captured input again); returns the graph-owned outgoing beam."""
        if lattice_epoch() != self.epoch:
            # something was (re)assigned somewhere: only edits of THIS lattice invalidate the graph
            if lattice_signature([self.segment]) != self.signature:
                raise RuntimeError(
                    "an element attribute of this lattice was re-assigned since capture; build a "
                    "new GraphedTrack (in-place updates of parameter tensors do not need this)"
                )
            self.epoch = lattice_epoch()
        if beam is not None and beam is not self.static_in:
            if self.unit_seventh:
                # the captured kernels add column 6 of the map instead of multiplying by the
                # seventh coordinate: the new beam must satisfy that too
                flag = getattr(beam, "_unit_seventh", None)
                if flag is None:
                    flag = bool((beam.particles[..., 6] == 1).all())
                if not flag:
                    raise ValueError(
                        "this graph was captured for beams with particles[..., 6] == 1; build a "
                        "new GraphedTrack for a beam whose seventh coordinate is not 1"
                    )
            self.static_in.particles.copy_(beam.particles)
            self.static_in.energy.copy_(beam.energy)
            self.static_in.particle_charges.copy_(beam.particle_charges)
            self.static_in.survival_probabilities.copy_(beam.survival_probabilities)
            self.static_in.s.copy_(beam.s)
        self.graph.replay()
        self._invalidate_screen_readings()
        return self.static_out

    def _invalidate_screen_readings(self) -> None:
        """Active screens share the graph-owned beam buffers: an image rendered from an earlier
        replay must not be served for this one."""
        for element in self.screens:
            if element.__dict__.get("_cached_reading") is not None:
                object.__setattr__(element, "_cached_reading", None)
