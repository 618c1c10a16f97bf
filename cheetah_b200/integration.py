"""The binding of INTEGRATION.md section 2 as code: a dispatch inside the REFERENCE's own
``Segment.track`` that sends CUDA beams through libcheetah_b200.so and everything else through
the reference's implementation.

    import cheetah, cheetah_b200.integration
    cheetah_b200.integration.install(cheetah)       # patches cheetah.Segment.track
    outgoing = segment.track(beam)                  # cheetah objects in, cheetah.ParticleBeam out

The accelerated path is taken when (cheetah/utils/cache.py:16-21 for the autograd
rule)
  * the beam is a ``ParticleBeam`` / ``ParameterBeam`` on a CUDA device in float32 / float64,
  * nothing that enters the result asks for gradients (autograd is enabled AND a beam tensor or
    a lattice tensor has ``requires_grad``) -- the kernels are forward-only,
  * every element of the lattice lowers (an element type or tracking method outside the hot
    path raises ``NotImplementedError`` during lowering: the call then falls through).
Anything else runs the reference's own ``Segment.track`` unchanged.  This module never imports
the reference: ``install`` is handed the module object.
"""

from __future__ import annotations

import threading

import torch

from . import lowering, tracking

_ORIGINAL = "_cheetah_b200_original_track"
_FLOATS = (torch.float32, torch.float64)

# dispatch statistics (tests and users can see which path a call took)
counters = {"accelerated": 0, "fallback": 0}
_state = threading.local()


def _beam_tensors(beam):
    for name in ("particles", "mu", "cov", "energy", "particle_charges", "total_charge",
                 "survival_probabilities", "s"):
        tensor = beam.__dict__.get("_buffers", {}).get(name)
        if tensor is None:
            tensor = getattr(beam, name, None) if name in ("particles", "mu", "cov") else None
        if isinstance(tensor, torch.Tensor):
            yield tensor
    species = getattr(beam, "species", None)
    for name in ("mass_eV", "num_elementary_charges"):
        tensor = getattr(species, name, None)
        if isinstance(tensor, torch.Tensor):
            yield tensor


def _lattice_tensors(elements):
    for element in lowering.flatten(elements):
        yield from element.parameters(recurse=True)
        yield from element.buffers(recurse=True)


def eligible(segment, incoming) -> bool:
    """True when ``segment.track(incoming)`` can run on the CUDA library (see module docstring);
    lowering may still refuse the lattice."""
    kind = type(incoming).__name__
    if kind == "ParticleBeam":
        state = incoming.particles
    elif kind == "ParameterBeam":
        state = incoming.mu
    else:
        return False
    if not (isinstance(state, torch.Tensor) and state.is_cuda and state.dtype in _FLOATS):
        return False
    if torch.is_grad_enabled():
        if any(t.requires_grad for t in _beam_tensors(incoming)):
            return False
        if any(t.requires_grad for t in _lattice_tensors(segment.elements)):
            return False
    return True


def _validity_key(elements) -> tuple:
    """What a cached lowering of reference elements is valid for: the element objects, in order,
    and identity + in-place version of every tensor they hold (cheetah/utils/cache.py:28-41 keys
    its transfer-map cache the same way) plus the non-tensor settings the lowering reads."""
    key = []
    for element in lowering.flatten(elements):
        tensors = tuple(
            (name, id(t), t._version)
            for name, t in list(element._buffers.items()) + list(element._parameters.items())
            if t is not None
        )
        plain = tuple(
            (name, value) for name, value in vars(element).items()
            if isinstance(value, (str, bool, int, float, tuple)) and not name.startswith("_")
        )
        key.append((id(element), tensors, plain, getattr(element, "tracking_method", None)))
    return tuple(key)


def _plan(segment, device, energy_shape):
    key = (device, tuple(energy_shape), _validity_key(segment.elements))
    cached = segment.__dict__.get("_cheetah_b200_plan")
    if cached is not None and cached[0] == key and not cached[1].is_stale():
        return cached[1]
    program = lowering.lower(list(segment.elements), device, tuple(energy_shape))
    object.__setattr__(segment, "_cheetah_b200_plan", (key, program))
    return program


def track(segment, incoming):
    """``Segment.track`` on the CUDA library for the reference's objects; raises
    ``NotImplementedError`` when the lattice does not lower."""
    state = incoming.particles if type(incoming).__name__ == "ParticleBeam" else incoming.mu
    program = _plan(segment, state.device, tuple(incoming.energy.shape))
    return tracking.track_program(program, incoming)


def install(cheetah) -> None:
    """Patch ``cheetah.Segment.track`` with the dispatcher (idempotent)."""
    segment_class = cheetah.Segment
    if hasattr(segment_class, _ORIGINAL):
        return
    original = segment_class.track

    def dispatch(self, incoming):
        # the reference's Segment.track builds sub-segments and tracks them in turn
        # (segment.py:548-572): one decision per user call, the nested calls follow it
        if getattr(_state, "inside_reference", False):
            return original(self, incoming)
        if eligible(self, incoming):
            try:
                outgoing = track(self, incoming)
            except NotImplementedError:
                pass  # an element outside the hot path: the reference tracks this lattice
            else:
                counters["accelerated"] += 1
                return outgoing
        counters["fallback"] += 1
        _state.inside_reference = True
        try:
            return original(self, incoming)
        finally:
            _state.inside_reference = False

    dispatch.__doc__ = original.__doc__
    setattr(segment_class, _ORIGINAL, original)
    segment_class.track = dispatch


def uninstall(cheetah) -> None:
    segment_class = cheetah.Segment
    original = getattr(segment_class, _ORIGINAL, None)
    if original is not None:
        segment_class.track = original
        delattr(segment_class, _ORIGINAL)
