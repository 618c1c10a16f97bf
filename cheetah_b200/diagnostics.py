"""Active ``BPM`` and ``Screen`` elements on the CUDA library.

Host-side mirror of cheetah/accelerator/bpm.py:77-86 and screen.py:187-344: the elements let the
beam through (a blocking screen zeroes the survival probabilities) and record a reading.  The BPM
reading comes from one ``ch_sc_beam_moments`` pass (fp64 sums about a pilot particle); the screen
image is computed lazily, on the first access of ``Screen.reading``, by ``ch_screen_image``.

Difference from the reference, on purpose: the beam handed on (and the beam a screen remembers)
SHARES the incoming tensors instead of deep-copying them (``incoming.clone()`` costs a full
56 B/particle pass); nothing in this backend mutates a beam in place.
"""

from __future__ import annotations

import math

import torch

from . import _capi


def _require_device(tensor: torch.Tensor, device, name: str, element) -> None:
    if tensor.device != device:
        raise ValueError(
            f"{name} of element {element.name!r} lives on {tensor.device} but the beam is on "
            f"{device}; move the lattice with `segment.to(device)` first"
        )


def _pass_through(beam, survival=None):
    outgoing = beam.__class__(
        beam.particles, beam.energy, particle_charges=beam.particle_charges,
        survival_probabilities=beam.survival_probabilities if survival is None else survival,
        s=beam.s, species=beam.species.clone(),
    )
    try:
        outgoing._unit_seventh = getattr(beam, "_unit_seventh", None)
    except Exception:
        pass
    return outgoing


def beam_centroid(beam) -> tuple[torch.Tensor, torch.Tensor]:
    """Survival-weighted (mu_x, mu_y) of a ParticleBeam (particle_beam.py:1699-1707, :1735-1743)
    with the vector shape the reference's properties have."""
    particles, survival = beam.particles, beam.survival_probabilities
    device, dtype = particles.device, particles.dtype
    n = particles.shape[-2]
    vp, vs = tuple(particles.shape[:-2]), tuple(survival.shape[:-1])
    vo = tuple(torch.broadcast_shapes(vp, vs))
    n_beams = max(1, math.prod(vo))
    if vp != vo and math.prod(vp) != 1:
        particles = particles.expand(*vo, n, 7)
    if vs != vo and math.prod(vs) != 1:
        survival = survival.expand(*vo, n)
    particles = particles.contiguous()
    survival = survival.to(dtype).contiguous()
    stats = torch.empty((n_beams, _capi.SC_STATS), dtype=torch.float64, device=device)
    with _capi.device_guard(device):
        _capi.check(_capi.lib().ch_sc_beam_moments(
            particles.data_ptr(), 0 if math.prod(vp) == 1 else n * 7,
            survival.data_ptr(), 0 if math.prod(vs) == 1 else n,
            n, n_beams, _capi.dtype_code(dtype), stats.data_ptr(), _capi.current_stream(device),
        ))
    mean = stats[:, 8:10] + stats[:, 2:4] / stats[:, 0:1]
    mean = mean.to(dtype).reshape(*vo, 2)
    return mean[..., 0], mean[..., 1]


def track_bpm(element, incoming):
    """bpm.py:77-86."""
    if element.is_active:
        if type(incoming).__name__ == "ParameterBeam":
            mu_x, mu_y = incoming.mu[..., 0], incoming.mu[..., 2]
            _require_device(element.misalignment, incoming.mu.device, "misalignment", element)
        else:
            _require_device(element.misalignment, incoming.particles.device, "misalignment", element)
            mu_x, mu_y = beam_centroid(incoming)
        # stored without going through Element.__setattr__: a reading is not a lattice change
        object.__setattr__(element, "_reading", torch.stack(
            [mu_x - element.misalignment[..., 0], mu_y - element.misalignment[..., 1]], dim=-1
        ))
    if type(incoming).__name__ == "ParameterBeam":
        return incoming.__class__(
            incoming.mu, incoming.cov, incoming.energy, total_charge=incoming.total_charge,
            s=incoming.s, species=incoming.species.clone(),
        )
    return _pass_through(incoming)


def track_screen(element, incoming):
    """screen.py:187-239."""
    parameter = type(incoming).__name__ == "ParameterBeam"
    if element.is_active:
        device = incoming.mu.device if parameter else incoming.particles.device
        _require_device(element.misalignment, device, "misalignment", element)
        _require_device(element.pixel_size, device, "pixel_size", element)
        element.set_read_beam(incoming)
    if element.is_active and element.is_blocking:
        if parameter:
            return incoming.__class__(
                incoming.mu, incoming.cov, incoming.energy,
                total_charge=torch.zeros_like(incoming.total_charge), s=incoming.s,
                species=incoming.species.clone(),
            )
        return _pass_through(incoming, torch.zeros_like(incoming.survival_probabilities))
    if parameter:
        return incoming.__class__(
            incoming.mu, incoming.cov, incoming.energy, total_charge=incoming.total_charge,
            s=incoming.s, species=incoming.species.clone(),
        )
    return _pass_through(incoming)


def screen_image(element, beam) -> torch.Tensor:
    """``Screen.reading`` for a ParticleBeam: ``(..., height, width)`` (screen.py:296-340)."""
    if type(beam).__name__ == "ParameterBeam":
        return _parameter_beam_image(element, beam)
    method = element.method
    particles = beam.particles
    device, dtype = particles.device, particles.dtype
    charges, survival = beam.particle_charges, beam.survival_probabilities
    misalignment = element.misalignment.to(dtype)
    if method == "histogram" and (
        particles.dim() > 2 or charges.dim() > 1 or beam.energy.dim() > 0
    ):
        raise NotImplementedError(
            "The `'histogram'` method of `Screen` does not support vectorization. Use `'kde'` "
            "instead. If this is a feature you would like to see, please open an issue on GitHub."
        )
    n = particles.shape[-2]
    shapes = [tuple(particles.shape[:-2]), tuple(charges.shape[:-1]), tuple(survival.shape[:-1]),
              tuple(misalignment.shape[:-1])]
    vo = tuple(torch.broadcast_shapes(*shapes))
    n_beams = max(1, math.prod(vo))

    def flat(tensor, inner):
        vector = tuple(tensor.shape[: tensor.dim() - len(inner)])
        if math.prod(vector) == 1:
            return tensor.to(dtype).contiguous(), 0
        if vector != vo:
            tensor = tensor.expand(*vo, *inner)
        return tensor.to(dtype).contiguous(), math.prod(inner)

    particles, particle_stride = flat(particles, (n, 7))
    charges, charge_stride = flat(charges, (n,))
    survival, survival_stride = flat(survival, (n,))
    misalignment, misalignment_stride = flat(misalignment, (2,))
    pixel_size = element.pixel_size.to(dtype).contiguous()
    resolution, binning = element.resolution, int(element.binning)
    nx, ny = element.effective_resolution
    edges_x = edges_y = None
    if method == "histogram":
        edges_x, edges_y = (e.to(dtype).contiguous() for e in element.pixel_bin_edges)
    image = torch.empty((n_beams, ny, nx), dtype=dtype, device=device)
    if method == "kde":
        # Gaussian kernel density estimate on the pixel centres (screen.py:312-326, kde.py:160-204)
        centers_x, centers_y = (c.to(dtype).contiguous() for c in element.pixel_bin_centers)
        bandwidth = element.kde_bandwidth
        if bandwidth is None:
            bandwidth = element.pixel_size[0]
        bandwidth = torch.as_tensor(bandwidth, dtype=dtype, device=device)
        if bandwidth.dim() != 0:
            raise ValueError(f"Input sigma must be a of the shape (1,). Got {bandwidth.shape}")
        bandwidth = bandwidth.reshape(1).contiguous()
        totals = torch.empty(n_beams, dtype=torch.float64, device=device)
        with _capi.device_guard(device):
            _capi.check(_capi.lib().ch_screen_kde(
                particles.data_ptr(), particle_stride, charges.data_ptr(), charge_stride,
                survival.data_ptr(), survival_stride, misalignment.data_ptr(),
                misalignment_stride, centers_x.data_ptr(), nx, centers_y.data_ptr(), ny,
                bandwidth.data_ptr(), n, n_beams, _capi.dtype_code(dtype), image.data_ptr(),
                totals.data_ptr(), _capi.current_stream(device),
            ))
        return image.reshape(*vo, ny, nx)
    with _capi.device_guard(device):
        _capi.check(_capi.lib().ch_screen_image(
            particles.data_ptr(), particle_stride, charges.data_ptr(), charge_stride,
            survival.data_ptr(), survival_stride, misalignment.data_ptr(), misalignment_stride,
            pixel_size.data_ptr(), int(resolution[0]), int(resolution[1]), binning,
            1 if method == "histogram" else 0, _capi.ptr(edges_x), _capi.ptr(edges_y),
            n, n_beams, _capi.dtype_code(dtype), image.data_ptr(), _capi.current_stream(device),
        ))
    return image.reshape(*vo, ny, nx)


def _parameter_beam_image(element, beam) -> torch.Tensor:
    """``Screen.reading`` for a ParameterBeam: the bivariate normal density of (x, y) on the pixel
    grid ``arange(left, right, step)`` (screen.py:251-289, ``ch_screen_gaussian``).  The read beam
    is shifted by the misalignment (:187-198)."""
    if torch.numel(beam.mu[..., 0]) > 1:
        raise NotImplementedError(
            "`Screen` does not support vectorization of `ParameterBeam`. Please use "
            "`ParticleBeam` instead. If this is a feature you would like to see, please open an "
            "issue on GitHub."
        )
    device, dtype = beam.mu.device, beam.mu.dtype
    mu = beam.mu.reshape(7).contiguous()
    cov = beam.cov.reshape(7, 7).to(dtype).contiguous()
    misalignment = element.misalignment.to(dtype).reshape(2).contiguous()
    # torch.arange(left, right, step) with tensor bounds: ceil((right - left) / step) points in
    # double, float32 values (the kernel reproduces the rounding); needs the bounds on the host
    extent = [float(v) for v in element.extent.to(dtype)]
    steps = [float(v) for v in (element.pixel_size.to(dtype) * int(element.binning))]
    nx = max(0, math.ceil((extent[1] - extent[0]) / steps[0]))
    ny = max(0, math.ceil((extent[3] - extent[2]) / steps[1]))
    image = torch.empty((ny, nx), dtype=dtype, device=device)
    with _capi.device_guard(device):
        _capi.check(_capi.lib().ch_screen_gaussian(
            mu.data_ptr(), cov.data_ptr(), misalignment.data_ptr(), extent[0], steps[0], nx,
            extent[2], steps[1], ny, _capi.dtype_code(dtype), image.data_ptr(),
            _capi.current_stream(device),
        ))
    return image
