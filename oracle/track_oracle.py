"""CPU oracle for the ``Segment.track(ParticleBeam)`` hot path.

TEST INFRASTRUCTURE ONLY -- this module is the *checker*, never the product.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package (``cheetah_b200``) never
imports anything under ``oracle/`` and has no CPU fallback.

It is a functional restatement, in plain PyTorch CPU ops, of the reference algorithm
(desy-ml/cheetah @ 60d1053, paths relative to /root/reference):

* linear transfer maps   cheetah/track_methods.py:17-77, :284-382 and the per-element
                         wrappers in cheetah/accelerator/*.py (cited per function)
* map merging + apply    cheetah/accelerator/segment.py:534-574,
                         cheetah/accelerator/element.py:159-191
* aperture               cheetah/accelerator/aperture.py:90-132
* space charge           cheetah/accelerator/space_charge_kick.py:103-586,
                         cheetah/utils/cloud_in_cell.py:244-384,
                         cheetah/particles/particle_beam.py:1262-1346,
                         cheetah/utils/statistics.py:30-62

Parity pin: ``tests/test_oracle_golden.py`` checks this oracle against (a) the
reference's own golden pickles ``tests/resources/consistency_expected_outgoing/*``
(converted to ``tests/golden/consistency.npz`` by ``oracle/make_golden.py``) at the
reference's own tolerance (rtol 1e-5, atol 1e-8, float64) and (b) outputs of the
unmodified reference run in the build container on ARES / FODO+space-charge cases
(``tests/golden/ares_*.npz``, ``tests/golden/space_charge_*.npz``).

A lattice is a list of plain dicts ``{"type": "Quadrupole", "length": t, "k1": t, ...}``
whose values are tensors (any broadcastable leading "vector" shape).  A beam is the
dict ``{"particles", "energy", "particle_charges", "survival_probabilities", "s",
"mass_eV", "num_elementary_charges"}``.
"""

from __future__ import annotations

import math

import torch

# CODATA values as used by the reference through scipy.constants
# (cheetah/particles/species.py:5-9, cheetah/accelerator/space_charge_kick.py:3).
SPEED_OF_LIGHT = 299792458.0
ELEMENTARY_CHARGE = 1.602176634e-19
EPSILON_0 = 8.8541878188e-12
ELECTRON_MASS_EV = 510998.95069
PROTON_MASS_EV = 938272089.4300001
EV_TO_KG = 1.7826619216278975e-36

LINEAR_TYPES = {
    "Drift",
    "Quadrupole",
    "Dipole",
    "RBend",
    "HorizontalCorrector",
    "VerticalCorrector",
    "CombinedCorrector",
    "Solenoid",
    "Undulator",
    "Cavity",
    "Marker",
    "BPM",
    "Screen",
    "Aperture",
    "CustomTransferMap",
    "Sextupole",
}


def _t(value, like: torch.Tensor) -> torch.Tensor:
    return torch.as_tensor(value, dtype=like.dtype, device=like.device)


def relativistic_factors(energy: torch.Tensor, mass_eV) -> tuple:
    """gamma, 1/gamma^2, beta  (cheetah/utils/physics.py:4-19)."""
    gamma = energy / mass_eV
    igamma2 = gamma.square().reciprocal()
    beta = (1.0 - igamma2).sqrt()
    return gamma, igamma2, beta


def _eye(vector_shape, like: torch.Tensor) -> torch.Tensor:
    return torch.eye(7, dtype=like.dtype, device=like.device).repeat(
        *vector_shape, 1, 1
    )


def si1mdiv(x: torch.Tensor) -> torch.Tensor:
    """(1 - sinc(sqrt(x))) / x with the 1/6 limit (cheetah/utils/autograd.py:108-128)."""
    sx = torch.complex(x, torch.zeros_like(x)).sqrt()
    safe = torch.where(x != 0, x, torch.ones_like(x))
    val = (1 - (sx / torch.pi).sinc().real) / safe
    return torch.where(x != 0, val, torch.full_like(x, 1.0 / 6.0))


def log1pdiv(x: torch.Tensor) -> torch.Tensor:
    """log(1+x)/x with limit 1 (cheetah/utils/autograd.py:77-105)."""
    safe = torch.where(x != 0, x, torch.ones_like(x))
    return torch.where(x != 0, torch.log1p(x) / safe, torch.ones_like(x))


def base_rmatrix(length, k1, hx, energy, mass_eV) -> torch.Tensor:
    """Thick quadrupole / sector-bend body map (cheetah/track_methods.py:17-77)."""
    zero = length.new_zeros(())
    _, igamma2, beta = relativistic_factors(energy, mass_eV)

    kx2 = k1 + hx.square()
    ky2 = -k1
    kx = torch.complex(kx2, zero.expand_as(kx2)).sqrt()
    ky = torch.complex(ky2, zero.expand_as(ky2)).sqrt()
    cx = (kx * length).cos().real
    cy = (ky * length).cos().real
    sx = ((kx * length / torch.pi).sinc() * length).real
    sy = ((ky * length / torch.pi).sinc() * length).real
    half = (0.5 * kx * length / torch.pi).sinc()
    dx = hx * 0.5 * length.square() * half.square().real
    r56 = (
        hx.square() * length.pow(3) * si1mdiv(kx2 * length.square()) / beta.square()
        - length / beta.square() * igamma2
    )

    shape = torch.broadcast_shapes(length.shape, k1.shape, hx.shape, energy.shape)
    R = _eye(shape, length)
    R[..., 0, 0] = cx
    R[..., 0, 1] = sx
    R[..., 0, 5] = dx / beta
    R[..., 1, 0] = -kx2 * sx
    R[..., 1, 1] = cx
    R[..., 1, 5] = sx * hx / beta
    R[..., 2, 2] = cy
    R[..., 2, 3] = sy
    R[..., 3, 2] = -ky2 * sy
    R[..., 3, 3] = cy
    R[..., 4, 0] = sx * hx / beta
    R[..., 4, 1] = dx / beta
    R[..., 4, 5] = r56
    return R


def drift_map(length, energy, mass_eV, extra=()) -> torch.Tensor:
    """Drift map (cheetah/track_methods.py:284-299); ``extra`` only widens the shape."""
    _, igamma2, beta = relativistic_factors(energy, mass_eV)
    shape = torch.broadcast_shapes(length.shape, igamma2.shape, *[e.shape for e in extra])
    R = _eye(shape, length)
    R[..., 0, 1] = length
    R[..., 2, 3] = length
    R[..., 4, 5] = -length / beta.square() * igamma2
    return R


def rotation_map(angle) -> torch.Tensor:
    """x-y rotation (cheetah/track_methods.py:302-323)."""
    cs, sn = angle.cos(), angle.sin()
    R = _eye(angle.shape, angle)
    R[..., 0, 0] = cs
    R[..., 0, 2] = sn
    R[..., 1, 1] = cs
    R[..., 1, 3] = sn
    R[..., 2, 0] = -sn
    R[..., 2, 2] = cs
    R[..., 3, 1] = -sn
    R[..., 3, 3] = cs
    return R


def tilt_misalignment_maps(angle, misalignment) -> tuple:
    """Entry/exit maps of a tilted, misaligned magnet (cheetah/track_methods.py:345-382)."""
    shape = torch.broadcast_shapes(angle.shape, misalignment.shape[:-1])
    entry = rotation_map(angle.expand(shape).clone())
    exit_ = entry.clone().mT.contiguous()
    cs, sn = angle.cos(), angle.sin()
    entry[..., 0, 6] = -misalignment[..., 0] * cs - misalignment[..., 1] * sn
    entry[..., 2, 6] = misalignment[..., 0] * sn - misalignment[..., 1] * cs
    exit_[..., 0, 6] = misalignment[..., 0]
    exit_[..., 2, 6] = misalignment[..., 1]
    return entry, exit_


def misalignment_maps(misalignment) -> tuple:
    """Pure shift entry/exit maps (cheetah/track_methods.py:326-342)."""
    shape = misalignment.shape[:-1]
    entry = _eye(shape, misalignment)
    exit_ = _eye(shape, misalignment)
    entry[..., 0, 6] = -misalignment[..., 0]
    entry[..., 2, 6] = -misalignment[..., 1]
    exit_[..., 0, 6] = misalignment[..., 0]
    exit_[..., 2, 6] = misalignment[..., 1]
    return entry, exit_


def _get(el: dict, key: str, like: torch.Tensor, default=0.0) -> torch.Tensor:
    value = el.get(key)
    if value is None:
        return _t(default, like)
    return torch.as_tensor(value, dtype=like.dtype, device=like.device)


def _dipole_edge(hx, e, fint, gap) -> torch.Tensor:
    """Dipole pole-face map (cheetah/accelerator/dipole.py:430-466)."""
    sec_e = e.cos().reciprocal()
    phi = fint * hx * gap * sec_e * (1 + e.sin().square())
    R = _eye(phi.shape, hx)
    R[..., 1, 0] = hx * e.tan()
    R[..., 3, 2] = -hx * (e - phi).tan()
    return R


def _cavity_map(el, energy, mass_eV, num_elementary_charges) -> torch.Tensor:
    """Cavity R-matrix (cheetah/accelerator/cavity.py:253-358)."""
    length = el["length"]
    voltage = _get(el, "voltage", length)
    phase = _get(el, "phase", length)
    frequency = _get(el, "frequency", length)
    cavity_type = el.get("cavity_type", "standing_wave")

    phi = phase.deg2rad()
    effective_voltage = -voltage * num_elementary_charges
    delta_energy = effective_voltage * phi.cos()
    Ei = energy / mass_eV
    dE = delta_energy / mass_eV
    Ef = Ei + dE
    Ep = dE / length
    k = 2 * torch.pi * frequency / SPEED_OF_LIGHT

    if cavity_type == "standing_wave":
        alpha = (
            math.sqrt(0.125) * effective_voltage / energy * log1pdiv(delta_energy / energy)
        )
        beta0 = (1 - Ei.square().reciprocal()).sqrt()
        beta1 = (1 - Ef.square().reciprocal()).sqrt()
        r11 = alpha.cos() - math.sqrt(2.0) * phi.cos() * alpha.sin()
        r12 = (alpha / torch.pi).sinc() * log1pdiv(delta_energy / energy) * length
        r21 = -(
            effective_voltage
            / ((energy + delta_energy) * math.sqrt(2.0) * length)
            * (0.5 + phi.cos().square())
            * alpha.sin()
        )
        r22 = Ei / Ef * (alpha.cos() + math.sqrt(2.0) * phi.cos() * alpha.sin())
        r55 = 1.0 + (
            k * length * beta0 * phi.tan() * (Ei * Ef * (beta0 * beta1 - 1) + 1)
            / (beta1 * Ef * dE)
        ).where(dE != 0.0, 0.0)
        r56 = -length / (Ef.square() * Ei * beta1) * (Ef + Ei) / (beta1 + beta0)
        r65 = k * phi.sin() * effective_voltage / (beta1 * (energy + delta_energy))
        r66 = Ei / Ef * beta0 / beta1
    elif cavity_type == "traveling_wave":
        body12 = length * log1pdiv(dE / Ei)
        body22 = Ei / Ef
        f_in = -Ep / (2 * Ei)
        f_out = Ep / (2 * Ef)
        # exit-fringe @ body @ entry-fringe, written out for the 2x2 case
        r11 = 1 + body12 * f_in
        r12 = body12 + torch.zeros_like(f_in)
        r21 = f_out * (1 + body12 * f_in) + body22 * f_in
        r22 = f_out * body12 + body22
        r55 = length.new_ones(())
        r56 = length.new_zeros(())
        r65 = k * torch.sin(phi) * effective_voltage / (energy + delta_energy)
        r66 = r22
    else:
        raise ValueError(f"Invalid cavity type: {cavity_type}")

    r11, r12, r21, r22, r55, r56, r65, r66 = torch.broadcast_tensors(
        r11, r12, r21, r22, r55, r56, r65, r66
    )
    R = _eye(r11.shape, length)
    R[..., 0, 0] = r11
    R[..., 0, 1] = r12
    R[..., 1, 0] = r21
    R[..., 1, 1] = r22
    R[..., 2, 2] = r11
    R[..., 2, 3] = r12
    R[..., 3, 2] = r21
    R[..., 3, 3] = r22
    R[..., 4, 4] = r55
    R[..., 4, 5] = r56
    R[..., 5, 4] = r65
    R[..., 5, 5] = r66
    return R


def element_length(el: dict, like: torch.Tensor) -> torch.Tensor:
    if el["type"] == "Segment":
        total = _t(0.0, like)
        for sub in el["elements"]:
            total = total + element_length(sub, like)
        return total
    return _get(el, "length", like)


def first_order_map(el: dict, energy: torch.Tensor, mass_eV, num_elementary_charges=-1.0):
    """7x7 first-order map of one element (per-type wrappers)."""
    kind = el["type"]
    like = energy
    length = _get(el, "length", like)

    if kind in ("Drift", "Sextupole"):
        # cheetah/accelerator/drift.py:61-65, sextupole.py:84-88
        return drift_map(length, energy, mass_eV)
    if kind == "Quadrupole":
        # cheetah/accelerator/quadrupole.py:93-110
        R = base_rmatrix(length, _get(el, "k1", like), length.new_zeros(()), energy, mass_eV)
        misalignment = _get(el, "misalignment", like, default=(0.0, 0.0))
        entry, exit_ = tilt_misalignment_maps(_get(el, "tilt", like), misalignment)
        return exit_ @ R @ entry
    if kind in ("Dipole", "RBend"):
        # cheetah/accelerator/dipole.py:372-394, rbend.py:83-101
        angle = _get(el, "angle", like)
        hx = angle / length
        e1 = _get(el, "dipole_e1", like)
        e2 = _get(el, "dipole_e2", like)
        if kind == "RBend":
            e1 = _get(el, "rbend_e1", like) + angle / 2
            e2 = _get(el, "rbend_e2", like) + angle / 2
        fint = _get(el, "fringe_integral", like)
        fint_exit = el.get("fringe_integral_exit")
        fint_exit = fint if fint_exit is None else _t(fint_exit, like)
        gap = _get(el, "gap", like)
        R = base_rmatrix(length, _get(el, "k1", like), hx, energy, mass_eV)
        R = _dipole_edge(hx, e2, fint_exit, gap) @ R @ _dipole_edge(hx, e1, fint, gap)
        rot = rotation_map(_get(el, "tilt", like))
        return rot.mT @ R @ rot
    if kind in ("HorizontalCorrector", "VerticalCorrector", "CombinedCorrector"):
        # horizontal_corrector.py:60-78, vertical_corrector.py:60-78,
        # combined_corrector.py:76-98
        if kind == "HorizontalCorrector":
            ah, av = _get(el, "angle", like), None
        elif kind == "VerticalCorrector":
            ah, av = None, _get(el, "angle", like)
        else:
            ah, av = _get(el, "horizontal_angle", like), _get(el, "vertical_angle", like)
        R = drift_map(length, energy, mass_eV, extra=[a for a in (ah, av) if a is not None])
        if ah is not None:
            R[..., 1, 6] = ah
        if av is not None:
            R[..., 3, 6] = av
        return R
    if kind == "Solenoid":
        # cheetah/accelerator/solenoid.py:74-116
        k = _get(el, "k", like)
        gamma, _, _ = relativistic_factors(energy, mass_eV)
        c = (length * k).cos()
        s = (length * k).sin()
        s_k = (length * k / torch.pi).sinc() * length
        shape = torch.broadcast_shapes(length.shape, k.shape, energy.shape)
        R = _eye(shape, length)
        R[..., 0, 0] = c.square()
        R[..., 0, 1] = c * s_k
        R[..., 0, 2] = s * c
        R[..., 0, 3] = s * s_k
        R[..., 1, 0] = -k * s * c
        R[..., 1, 1] = c.square()
        R[..., 1, 2] = -k * s.square()
        R[..., 1, 3] = s * c
        R[..., 2, 0] = -s * c
        R[..., 2, 1] = -s * s_k
        R[..., 2, 2] = c.square()
        R[..., 2, 3] = c * s_k
        R[..., 3, 0] = k * s.square()
        R[..., 3, 1] = -s * c
        R[..., 3, 2] = -k * s * c
        R[..., 3, 3] = c.square()
        R[..., 4, 5] = length / (1 - gamma.square())
        entry, exit_ = misalignment_maps(_get(el, "misalignment", like, default=(0.0, 0.0)))
        return exit_ @ R @ entry
    if kind == "Undulator":
        # cheetah/accelerator/undulator.py:78-125
        kx = _get(el, "kx", like)
        ky = _get(el, "ky", like)
        period = _get(el, "period", like, default=0.0)
        gamma, igamma2, beta = relativistic_factors(energy, mass_eV)
        shape = torch.broadcast_shapes(
            length.shape, igamma2.shape, kx.shape, ky.shape, period.shape
        )
        R = _eye(shape, length)
        R[..., 4, 5] = (
            -length * igamma2 * (beta.square().reciprocal() + 0.5 * (kx.square() + ky.square()))
        )
        spatial_frequency = torch.where(
            period > 0.0,
            math.sqrt(2) * torch.pi / (period * gamma * beta),
            torch.zeros((), dtype=like.dtype),
        )
        omega_x = spatial_frequency * kx
        R[..., 2, 2] = (omega_x * length).cos()
        R[..., 2, 3] = (omega_x * length / torch.pi).sinc() * length
        R[..., 3, 2] = -(omega_x * length).sin() * omega_x
        R[..., 3, 3] = (omega_x * length).cos()
        omega_y = spatial_frequency * ky
        R[..., 0, 0] = (omega_y * length).cos()
        R[..., 0, 1] = (omega_y * length / torch.pi).sinc() * length
        R[..., 1, 0] = -(omega_y * length).sin() * omega_y
        R[..., 1, 1] = (omega_y * length).cos()
        return R
    if kind == "Cavity":
        return _cavity_map(el, energy, mass_eV, num_elementary_charges)
    if kind in ("Marker", "BPM", "Screen", "Aperture"):
        # marker.py:44-50, bpm.py:69-75, screen.py:176-185, aperture.py:82-88
        return _eye(energy.shape, energy)
    if kind == "CustomTransferMap":
        # custom_transfer_map.py:111-114
        return torch.as_tensor(el["predefined_transfer_map"], dtype=like.dtype)
    if kind == "Segment":
        # cheetah/accelerator/segment.py:534-541
        tm = torch.eye(7, dtype=like.dtype)
        for sub in el["elements"]:
            tm = first_order_map(sub, energy, mass_eV, num_elementary_charges) @ tm
        return tm
    raise NotImplementedError(f"oracle has no linear map for element type {kind}")


def is_skippable(el: dict) -> bool:
    """Whether the element merges into a linear run."""
    kind = el["type"]
    if kind == "Segment":
        return all(is_skippable(sub) for sub in el["elements"])
    if kind in ("Drift", "Quadrupole", "Dipole", "RBend", "Sextupole"):
        return el.get("tracking_method", "linear") == "linear"
    if kind in ("BPM", "Screen"):
        return not el.get("is_active", False)
    if kind == "Aperture":
        return not el.get("is_active", True)
    if kind == "Cavity":
        voltage = el.get("voltage")
        return voltage is None or not bool((torch.as_tensor(voltage) != 0).any())
    if kind == "SpaceChargeKick":
        return False
    return kind in LINEAR_TYPES


def flatten(elements: list) -> list:
    """Flatten nested Segments (cheetah/accelerator/segment.py:143-157)."""
    flat = []
    for el in elements:
        if el["type"] == "Segment":
            flat.extend(flatten(el["elements"]))
        else:
            flat.append(el)
    return flat


def _with(beam: dict, **updates) -> dict:
    out = dict(beam)
    out.update(updates)
    return out


def track_linear_run(run: list, beam: dict) -> dict:
    """Merged map of a skippable run applied once (segment.py:534-541, element.py:181-191)."""
    energy = beam["energy"]
    tm = torch.eye(7, dtype=energy.dtype)
    length = _t(0.0, energy)
    for el in run:
        tm = first_order_map(el, energy, beam["mass_eV"], beam["num_elementary_charges"]) @ tm
        length = length + element_length(el, energy)
    return _with(beam, particles=beam["particles"] @ tm.mT, s=beam["s"] + length)


def track_aperture(el: dict, beam: dict) -> dict:
    """Aperture survival mask (cheetah/accelerator/aperture.py:90-132)."""
    particles = beam["particles"]
    x, y = particles[..., 0], particles[..., 2]
    x_max = _get(el, "x_max", particles, default=float("inf"))
    y_max = _get(el, "y_max", particles, default=float("inf"))
    assert (x_max >= 0).all() and (y_max >= 0).all()
    shape = el.get("shape", "rectangular")
    if shape == "rectangular":
        mask = torch.logical_and(
            torch.logical_and(x > -x_max.unsqueeze(-1), x < x_max.unsqueeze(-1)),
            torch.logical_and(y > -y_max.unsqueeze(-1), y < y_max.unsqueeze(-1)),
        )
    elif shape == "elliptical":
        mask = (
            x.square() / x_max.square().unsqueeze(-1)
            + y.square() / y_max.square().unsqueeze(-1)
        ) <= 1.0
    else:
        raise AssertionError(f"Unknown aperture shape {shape}")
    return _with(beam, survival_probabilities=beam["survival_probabilities"] * mask)


def track_cavity(el: dict, beam: dict) -> dict:
    """Active cavity, ParticleBeam branch (cheetah/accelerator/cavity.py:100-251): linear R,
    then the exact energy-deviation update and second-order longitudinal terms."""
    particles = beam["particles"]
    energy, mass_eV = beam["energy"], beam["mass_eV"]
    charges = beam["num_elementary_charges"]
    length = _get(el, "length", particles)
    voltage = _get(el, "voltage", particles)
    phase = _get(el, "phase", particles)
    frequency = _get(el, "frequency", particles)

    gamma0, igamma2, beta0 = relativistic_factors(energy, mass_eV)
    phi = phase.deg2rad()
    tm = _cavity_map(el, energy, mass_eV, charges)
    out = particles @ tm.mT
    delta_energy = voltage * phi.cos() * charges * -1
    T566 = 1.5 * length * igamma2 / beta0.pow(3)
    T556 = length.new_zeros(())
    T555 = length.new_zeros(())
    k = 2.0 * torch.pi * frequency / SPEED_OF_LIGHT
    outgoing_energy = energy + delta_energy
    gamma1, _, beta1 = relativistic_factors(outgoing_energy, mass_eV)

    out[..., 5] = particles[..., 5] * energy.unsqueeze(-1) * beta0.unsqueeze(-1) / (
        outgoing_energy.unsqueeze(-1) * beta1.unsqueeze(-1)
    ) + voltage.unsqueeze(-1) * beta0.unsqueeze(-1) / (
        outgoing_energy.unsqueeze(-1) * beta1.unsqueeze(-1)
    ) * (
        (-particles[..., 4] * beta0.unsqueeze(-1) * k.unsqueeze(-1) + phi.unsqueeze(-1)).cos()
        - phi.cos().unsqueeze(-1)
    )
    dgamma = voltage / mass_eV
    if (delta_energy > 0).any():
        T566 = (
            length
            * (beta0.pow(3) * gamma0.pow(3) - beta1.pow(3) * gamma1.pow(3))
            / (2.0 * beta0 * beta1.pow(3) * gamma0 * (gamma0 - gamma1) * gamma1.pow(3))
        )
        T556 = (
            beta0 * k * length * dgamma * gamma0
            * (beta1.pow(3) * gamma1.pow(3) + beta0 * (gamma0 - gamma1.pow(3)))
            * phi.sin()
            / (beta1.pow(3) * gamma1.pow(3) * (gamma0 - gamma1).square())
        )
        T555 = (
            beta0.square() * k.square() * length * dgamma / 2.0
            * (
                dgamma
                * (
                    2.0 * gamma0 * gamma1.pow(3) * (beta0 * beta1.pow(3) - 1.0)
                    + gamma0.square() + 3.0 * gamma1.square() - 2.0
                )
                / (beta1.pow(3) * gamma1.pow(3) * (gamma0 - gamma1).pow(3))
                * phi.sin().square()
                - (gamma1 * gamma0 * (beta1 * beta0 - 1.0) + 1.0)
                / (beta1 * gamma1 * (gamma0 - gamma1).square())
                * phi.cos()
            )
        )
    out[..., 4] = out[..., 4] + (
        T566.unsqueeze(-1) * particles[..., 5].square()
        + T556.unsqueeze(-1) * particles[..., 4] * particles[..., 5]
        + T555.unsqueeze(-1) * particles[..., 4].square()
    )
    return _with(beam, particles=out, energy=outgoing_energy, s=beam["s"] + length)


def track_parameter_beam(elements: list, mu, cov, energy, mass_eV, num_elementary_charges=-1.0):
    """``Segment.track`` for a ParameterBeam (element.py:166-179, cavity.py:100-251): linear maps
    on (mu, cov); an active cavity adds its special updates of the longitudinal entries.
    Returns (mu, cov, energy)."""
    run: list = []

    def flush():
        nonlocal mu, cov, run
        for el in run:
            tm = first_order_map(el, energy, mass_eV, num_elementary_charges)
            mu = (tm @ mu.unsqueeze(-1)).squeeze(-1)
            cov = tm @ cov @ tm.mT
        run = []

    for el in flatten(elements):
        if is_skippable(el) or el["type"] == "Aperture":  # apertures ignore ParameterBeams
            if el["type"] != "Aperture":
                run.append(el)
            continue
        flush()
        assert el["type"] == "Cavity", el["type"]
        length = _get(el, "length", mu)
        voltage = _get(el, "voltage", mu)
        phi = _get(el, "phase", mu).deg2rad()
        frequency = _get(el, "frequency", mu)
        gamma0, igamma2, beta0 = relativistic_factors(energy, mass_eV)
        tm = _cavity_map(el, energy, mass_eV, num_elementary_charges)
        out_mu = (tm @ mu.unsqueeze(-1)).squeeze(-1)
        out_cov = tm @ cov @ tm.mT
        delta_energy = voltage * phi.cos() * num_elementary_charges * -1
        T566 = 1.5 * length * igamma2 / beta0.pow(3)
        T556 = length.new_zeros(())
        T555 = length.new_zeros(())
        k = 2.0 * torch.pi * frequency / SPEED_OF_LIGHT
        outgoing_energy = energy + delta_energy
        gamma1, _, beta1 = relativistic_factors(outgoing_energy, mass_eV)
        out_mu[..., 5] = mu[..., 5] * energy * beta0 / (outgoing_energy * beta1) + voltage * beta0 / (
            outgoing_energy * beta1
        ) * ((-mu[..., 4] * beta0 * k + phi).cos() - phi.cos())
        out_cov[..., 5, 5] = cov[..., 5, 5]
        dgamma = voltage / mass_eV
        if (delta_energy > 0).any():
            T566 = (
                length * (beta0.pow(3) * gamma0.pow(3) - beta1.pow(3) * gamma1.pow(3))
                / (2.0 * beta0 * beta1.pow(3) * gamma0 * (gamma0 - gamma1) * gamma1.pow(3))
            )
            T556 = (
                beta0 * k * length * dgamma * gamma0
                * (beta1.pow(3) * gamma1.pow(3) + beta0 * (gamma0 - gamma1.pow(3))) * phi.sin()
                / (beta1.pow(3) * gamma1.pow(3) * (gamma0 - gamma1).square())
            )
            T555 = (
                beta0.square() * k.square() * length * dgamma / 2.0
                * (
                    dgamma
                    * (2.0 * gamma0 * gamma1.pow(3) * (beta0 * beta1.pow(3) - 1.0)
                       + gamma0.square() + 3.0 * gamma1.square() - 2.0)
                    / (beta1.pow(3) * gamma1.pow(3) * (gamma0 - gamma1).pow(3)) * phi.sin().square()
                    - (gamma1 * gamma0 * (beta1 * beta0 - 1.0) + 1.0)
                    / (beta1 * gamma1 * (gamma0 - gamma1).square()) * phi.cos()
                )
            )
        out_mu[..., 4] = out_mu[..., 4] + (
            T566 * mu[..., 5].square() + T556 * mu[..., 4] * mu[..., 5] + T555 * mu[..., 4].square()
        )
        longitudinal = (
            T566 * cov[..., 5, 5].square() + T556 * cov[..., 4, 5] * cov[..., 5, 5]
            + T555 * cov[..., 4, 4].square()
        )
        out_cov[..., 4, 4] = longitudinal
        out_cov[..., 4, 5] = longitudinal
        out_cov[..., 5, 4] = longitudinal
        mu, cov, energy = out_mu, out_cov, outgoing_energy
    flush()
    return mu, cov, energy


# --------------------------------------------------------------------------------------
# Space charge
# --------------------------------------------------------------------------------------


def weighted_std(values: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """Unbiased weighted standard deviation (cheetah/utils/statistics.py:30-62)."""
    sum_w = weights.sum(dim=-1)
    mean = (values * weights).sum(dim=-1) / sum_w
    correction = sum_w - weights.square().sum(dim=-1) / sum_w
    var = (weights * (values - mean.unsqueeze(-1)).square()).sum(dim=-1) / correction
    return var.sqrt()


def reference_beta(gamma: torch.Tensor) -> torch.Tensor:
    """Beam.relativistic_beta (cheetah/particles/beam.py:328-336)."""
    beta = torch.ones_like(gamma)
    nonzero = gamma.abs() > 0
    beta[nonzero] = (1 - gamma[gamma > 0].square().reciprocal()).sqrt()
    return beta


def to_xyz_pxpypz(particles, energy, mass_eV) -> torch.Tensor:
    """Cheetah -> SI coordinates (cheetah/particles/particle_beam.py:1316-1346)."""
    mass_kg = mass_eV * EV_TO_KG
    gamma0 = energy / mass_eV
    beta0 = reference_beta(gamma0)
    p0 = gamma0 * beta0 * mass_kg * SPEED_OF_LIGHT
    gamma = gamma0.unsqueeze(-1) * (1.0 + particles[..., 5] * beta0.unsqueeze(-1))
    beta = (1 - gamma.square().reciprocal()).sqrt()
    momentum = gamma * mass_kg * beta * SPEED_OF_LIGHT
    px = particles[..., 1] * p0.unsqueeze(-1)
    py = particles[..., 3] * p0.unsqueeze(-1)
    zs = particles[..., 4] * -beta0.unsqueeze(-1)
    pz = (momentum.square() - px.square() - py.square()).sqrt()
    out = particles.clone()
    out[..., 1] = px
    out[..., 3] = py
    out[..., 4] = zs
    out[..., 5] = pz
    return out


def from_xyz_pxpypz(xp, energy, mass_eV) -> torch.Tensor:
    """SI -> Cheetah coordinates (cheetah/particles/particle_beam.py:1262-1314)."""
    mass_kg = mass_eV * EV_TO_KG
    gamma0 = energy / mass_eV
    beta0 = reference_beta(gamma0)
    p0 = gamma0 * beta0 * mass_kg * SPEED_OF_LIGHT
    p = (xp[..., 1].square() + xp[..., 3].square() + xp[..., 5].square()).sqrt()
    gamma = (1 + (p / (mass_kg * SPEED_OF_LIGHT)).square()).sqrt()
    out = xp.clone()
    out[..., 1] = xp[..., 1] / p0.unsqueeze(-1)
    out[..., 3] = xp[..., 3] / p0.unsqueeze(-1)
    out[..., 4] = -xp[..., 4] / beta0.unsqueeze(-1)
    out[..., 5] = (gamma - gamma0.unsqueeze(-1)) / (beta0 * gamma0).unsqueeze(-1)
    return out


def cic_deposit_nd(positions, bins, extent, charges) -> torch.Tensor:
    """1-, 2- or 3-D cloud-in-cell deposit (cheetah/utils/cloud_in_cell.py:67-384; the three
    specialisations share the per-axis rule): positions (..., N, d), extent (..., d, 2),
    charges (..., N) -> (..., *bins)."""
    dims = positions.shape[-1]
    nbins = [int(n) for n in bins]
    vector_shape = positions.shape[:-2]
    total = 1
    for n in nbins:
        total *= n
    grid = positions.new_zeros(*vector_shape, total)
    inside = torch.ones_like(charges, dtype=torch.bool)
    corners = []
    for d in range(dims):
        p = positions[..., d]
        left, right = extent[..., d, 0].unsqueeze(-1), extent[..., d, 1].unsqueeze(-1)
        inside = inside & (p >= left) & (p <= right)
        q = (p - left) / ((right - left) / nbins[d]) - 0.5
        base = q.floor().long()
        frac = q - base
        corners.append([
            (base.clamp(0, nbins[d] - 1), (1.0 - frac) * ((base >= 0) & (base < nbins[d]))),
            ((base + 1).clamp(0, nbins[d] - 1), frac * ((base + 1 >= 0) & (base + 1 < nbins[d]))),
        ])
    masked = charges * inside
    import itertools

    for choice in itertools.product((0, 1), repeat=dims):
        index = torch.zeros_like(corners[0][0][0])
        weight = masked
        for d, c in enumerate(choice):
            index = index * nbins[d] + corners[d][c][0]
            weight = weight * corners[d][c][1]
        grid.scatter_add_(dim=-1, index=index, src=weight)
    return grid.reshape(*vector_shape, *nbins)


def cic_deposit_3d(positions, bins, extent, charges) -> torch.Tensor:
    """3-D cloud-in-cell deposit (cheetah/utils/cloud_in_cell.py:244-384).

    positions (B, N, 3); extent (B, 3, 2); charges (B, N) -> (B, nx, ny, nz).
    Cell-centred convention: bin-space position = (p - left) / width - 0.5.
    """
    nbins = list(bins)
    grid = positions.new_zeros(*positions.shape[:-2], nbins[0] * nbins[1] * nbins[2])
    inside = torch.ones_like(charges, dtype=torch.bool)
    lo_idx, frac, lo_ok, hi_ok = [], [], [], []
    for d in range(3):
        p = positions[..., d].contiguous()
        left = extent[..., d, 0].unsqueeze(-1)
        right = extent[..., d, 1].unsqueeze(-1)
        inside = inside & (p >= left) & (p <= right)
        width = (right - left) / nbins[d]
        q = (p - left) / width - 0.5
        qi = q.floor().long()
        lo_idx.append(qi)
        frac.append(q - qi)
        lo_ok.append((qi >= 0) & (qi < nbins[d]))
        hi_ok.append((qi + 1 >= 0) & (qi + 1 < nbins[d]))
    masked = charges * inside
    strides = (nbins[1] * nbins[2], nbins[2], 1)
    for ox in (0, 1):
        for oy in (0, 1):
            for oz in (0, 1):
                idx = 0
                weight = 1.0
                for d, o in enumerate((ox, oy, oz)):
                    corner = (lo_idx[d] + o).clamp(0, nbins[d] - 1)
                    idx = idx + corner * strides[d]
                    w = (1.0 - frac[d]) * lo_ok[d] if o == 0 else frac[d] * hi_ok[d]
                    weight = weight * w if not isinstance(weight, float) else w
                grid.scatter_add_(dim=-1, index=idx, src=masked * weight)
    return grid.reshape(*positions.shape[:-2], *nbins)


def _igf_antiderivative(x, y, tau) -> torch.Tensor:
    """Antiderivative of 1/r (cheetah/accelerator/space_charge_kick.py:103-123)."""
    r = (x.square() + y.square() + tau.square()).sqrt()
    return (
        -0.5 * tau.square() * (x * y / (tau * r)).atan()
        - 0.5 * y.square() * (x * tau / (y * r)).atan()
        - 0.5 * x.square() * (y * tau / (x * r)).atan()
        + y * tau * (x / (y.square() + tau.square()).sqrt()).asinh()
        + x * tau * (y / (x.square() + tau.square()).sqrt()).asinh()
        + x * y * (tau / (x.square() + y.square()).sqrt()).asinh()
    )


def integrated_green_function(cell_size, gamma, grid_shape) -> torch.Tensor:
    """IGF on the doubled grid (cheetah/accelerator/space_charge_kick.py:163-291).

    cell_size (B, 3); gamma (B,) -> (B, 2nx, 2ny, 2nz).  Only d_tau is scaled by gamma and
    plane index n of each doubled axis stays zero, as in the reference.
    """
    nx, ny, nz = grid_shape
    dx = cell_size[..., 0][..., None, None, None]
    dy = cell_size[..., 1][..., None, None, None]
    dt = (cell_size[..., 2] * gamma)[..., None, None, None]
    ix, iy, it = torch.meshgrid(
        torch.arange(nx), torch.arange(ny), torch.arange(nz), indexing="ij"
    )
    xg, yg, tg = ix[None] * dx, iy[None] * dy, it[None] * dt
    G = None
    # Same evaluation order as the reference's 8-term signed sum (:195-236)
    for sx, sy, st, sign in (
        (+1, +1, +1, +1),
        (-1, +1, +1, -1),
        (+1, -1, +1, -1),
        (+1, +1, -1, -1),
        (+1, -1, -1, +1),
        (-1, +1, -1, +1),
        (-1, -1, +1, +1),
        (-1, -1, -1, -1),
    ):
        term = _igf_antiderivative(
            xg + sx * 0.5 * dx, yg + sy * 0.5 * dy, tg + st * 0.5 * dt
        )
        G = term if G is None else (G + term if sign > 0 else G - term)

    out = cell_size.new_zeros(*cell_size.shape[:-1], 2 * nx, 2 * ny, 2 * nz)
    out[..., :nx, :ny, :nz] = G
    out[..., nx + 1 :, :ny, :nz] = G[..., 1:, :, :].flip(dims=[-3])
    out[..., :nx, ny + 1 :, :nz] = G[..., :, 1:, :].flip(dims=[-2])
    out[..., :nx, :ny, nz + 1 :] = G[..., :, :, 1:].flip(dims=[-1])
    out[..., nx + 1 :, ny + 1 :, :nz] = G[..., 1:, 1:, :].flip(dims=[-3, -2])
    out[..., :nx, ny + 1 :, nz + 1 :] = G[..., :, 1:, 1:].flip(dims=[-2, -1])
    out[..., nx + 1 :, :ny, nz + 1 :] = G[..., 1:, :, 1:].flip(dims=[-3, -1])
    out[..., nx + 1 :, ny + 1 :, nz + 1 :] = G[..., 1:, 1:, 1:].flip(dims=[-3, -2, -1])
    return out


def space_charge_potential(xp, charges, survival, cell_size, grid_dimensions, gamma, grid_shape):
    """Hockney/IGF Poisson solve (space_charge_kick.py:125-161, :293-322)."""
    nx, ny, nz = grid_shape
    rho = cic_deposit_3d(
        xp[..., [0, 2, 4]],
        grid_shape,
        torch.stack([-grid_dimensions, grid_dimensions], dim=-1),
        charges * survival,
    )
    rho = rho * cell_size.prod(dim=-1).reciprocal()[..., None, None, None]
    padded = xp.new_zeros(*xp.shape[:-2], 2 * nx, 2 * ny, 2 * nz)
    padded[..., :nx, :ny, :nz] = rho
    rho_ft = torch.fft.rfftn(padded, dim=[1, 2, 3])
    green_ft = torch.fft.rfftn(
        integrated_green_function(cell_size, gamma, grid_shape), dim=[1, 2, 3]
    )
    potential = (1.0 / (4 * torch.pi * EPSILON_0)) * torch.fft.irfftn(
        rho_ft * green_ft, dim=[1, 2, 3]
    ).real
    return potential[..., :nx, :ny, :nz]


def space_charge_field(potential, cell_size, gamma) -> tuple:
    """-(1/gamma^2) grad(phi), central differences, zero boundary (:324-365)."""
    inv = cell_size.reciprocal()
    igamma2 = torch.zeros_like(gamma)
    igamma2[gamma != 0] = gamma[gamma != 0].square().reciprocal()
    gx = torch.zeros_like(potential)
    gy = torch.zeros_like(potential)
    gt = torch.zeros_like(potential)
    gx[..., 1:-1, :, :] = (potential[..., 2:, :, :] - potential[..., :-2, :, :]) * (
        0.5 * inv[..., 0, None, None, None]
    )
    gy[..., :, 1:-1, :] = (potential[..., :, 2:, :] - potential[..., :, :-2, :]) * (
        0.5 * inv[..., 1, None, None, None]
    )
    gt[..., :, :, 1:-1] = (potential[..., :, :, 2:] - potential[..., :, :, :-2]) * (
        0.5 * inv[..., 2, None, None, None]
    )
    scale = -igamma2[..., None, None, None]
    return scale * gx, scale * gy, scale * gt


def gather_forces(xp, fields, cell_size, grid_dimensions, grid_shape) -> torch.Tensor:
    """Node-centred trilinear gather x elementary charge (space_charge_kick.py:367-475)."""
    fx, fy, fz = fields
    B, N = xp.shape[0], xp.shape[1]
    pos = xp[..., [0, 2, 4]]
    norm = (pos + grid_dimensions.unsqueeze(-2)) / cell_size.unsqueeze(-2)
    base = norm.floor().to(torch.int)
    forces = xp.new_zeros(B, N, 3)
    batch = torch.arange(B).unsqueeze(-1).expand(B, N)
    for ox in (0, 1):
        for oy in (0, 1):
            for oz in (0, 1):
                corner = base + torch.tensor([ox, oy, oz], dtype=torch.int)
                weight = (1 - (norm - corner).abs()).prod(dim=-1)
                valid = (
                    (corner[..., 0] >= 0)
                    & (corner[..., 0] < grid_shape[0])
                    & (corner[..., 1] >= 0)
                    & (corner[..., 1] < grid_shape[1])
                    & (corner[..., 2] >= 0)
                    & (corner[..., 2] < grid_shape[2])
                )
                cx = corner[..., 0].clamp(0, grid_shape[0] - 1).long()
                cy = corner[..., 1].clamp(0, grid_shape[1] - 1).long()
                cz = corner[..., 2].clamp(0, grid_shape[2] - 1).long()
                w = weight * ELEMENTARY_CHARGE
                for comp, grid in enumerate((fx, fy, fz)):
                    forces[..., comp] += w * grid[batch, cx, cy, cz].where(valid, 0)
    return forces


def track_space_charge(el: dict, beam: dict) -> dict:
    """One space-charge kick (cheetah/accelerator/space_charge_kick.py:477-586)."""
    particles = beam["particles"]
    N = particles.shape[-2]
    dtype = particles.dtype
    effect_length = torch.as_tensor(el["effect_length"], dtype=dtype)
    grid_shape = tuple(el.get("grid_shape", (32, 32, 32)))
    extents = [
        _get(el, key, particles, default=3.0)
        for key in ("grid_extent_x", "grid_extent_y", "grid_extent_tau")
    ]
    vector_shape = torch.broadcast_shapes(
        particles.shape[:-2],
        beam["energy"].shape,
        beam["particle_charges"].shape[:-1],
        beam["survival_probabilities"].shape[:-1],
        (1,),
    )
    p = particles.broadcast_to(*vector_shape, N, 7).flatten(end_dim=-3)
    energy = beam["energy"].broadcast_to(vector_shape).flatten()
    charges = beam["particle_charges"].broadcast_to(*vector_shape, N).flatten(end_dim=-2)
    survival = (
        beam["survival_probabilities"].broadcast_to(*vector_shape, N).flatten(end_dim=-2)
    )
    mass_eV = beam["mass_eV"]
    gamma = energy / mass_eV
    beta = reference_beta(gamma)

    grid_dimensions = torch.stack(
        [
            extents[0] * weighted_std(p[..., 0], survival),
            extents[1] * weighted_std(p[..., 2], survival),
            extents[2] * weighted_std(p[..., 4], survival),
        ],
        dim=-1,
    )
    cell_size = 2 * grid_dimensions / torch.tensor(grid_shape, dtype=dtype)
    dt = effect_length.flatten() / (SPEED_OF_LIGHT * beta)

    xp = to_xyz_pxpypz(p, energy, mass_eV)
    potential = space_charge_potential(
        xp, charges, survival, cell_size, grid_dimensions, gamma, grid_shape
    )
    fields = space_charge_field(potential, cell_size, gamma)
    forces = gather_forces(xp, fields, cell_size, grid_dimensions, grid_shape)
    xp[..., 1] = xp[..., 1] + forces[..., 0] * dt.unsqueeze(-1)
    xp[..., 3] = xp[..., 3] + forces[..., 1] * dt.unsqueeze(-1)
    xp[..., 5] = xp[..., 5] + forces[..., 2] * dt.unsqueeze(-1)

    out_shape = torch.broadcast_shapes(
        particles.shape[:-2],
        beam["energy"].shape,
        beam["particle_charges"].shape[:-1],
        beam["survival_probabilities"].shape[:-1],
        effect_length.shape,
    )
    xp = xp.reshape(*out_shape, N, 7)
    return _with(beam, particles=from_xyz_pxpypz(xp, beam["energy"], mass_eV))


# --------------------------------------------------------------------------------------
# Segment.track
# --------------------------------------------------------------------------------------


def track(elements: list, beam: dict) -> dict:
    """``Segment.track`` for a ParticleBeam (cheetah/accelerator/segment.py:545-574)."""
    run: list = []
    for el in flatten(elements):
        if is_skippable(el):
            run.append(el)
            continue
        if run:
            beam = track_linear_run(run, beam)
            run = []
        if el["type"] == "Aperture":
            beam = track_aperture(el, beam)
        elif el["type"] == "SpaceChargeKick":
            beam = track_space_charge(el, beam)
        elif el["type"] == "Cavity":
            beam = track_cavity(el, beam)
        elif el["type"] in (
            "Drift", "Quadrupole", "Dipole", "RBend", "Sextupole", "TransverseDeflectingCavity"
        ):
            from . import nonlinear_oracle  # drift_kick_drift / second_order

            beam = nonlinear_oracle.track_nonlinear(el, beam)
        else:
            raise NotImplementedError(
                f"oracle: non-skippable element type {el['type']} is outside the hot path"
            )
    if run:
        beam = track_linear_run(run, beam)
    return beam


def make_beam(
    particles,
    energy,
    particle_charges=None,
    survival_probabilities=None,
    s=None,
    mass_eV=ELECTRON_MASS_EV,
    num_elementary_charges=-1.0,
) -> dict:
    """ParticleBeam constructor defaults (cheetah/particles/particle_beam.py:60-106)."""
    n = particles.shape[-2]
    dtype = particles.dtype
    if particle_charges is None:
        particle_charges = torch.full(
            (n,), num_elementary_charges * ELEMENTARY_CHARGE, dtype=dtype
        )
    if survival_probabilities is None:
        survival_probabilities = torch.ones(n, dtype=dtype)
    if s is None:
        s = torch.tensor(0.0, dtype=dtype)
    return {
        "particles": particles,
        "energy": torch.as_tensor(energy, dtype=dtype),
        "particle_charges": particle_charges,
        "survival_probabilities": survival_probabilities,
        "s": s,
        "mass_eV": torch.as_tensor(mass_eV, dtype=dtype),
        "num_elementary_charges": torch.as_tensor(num_elementary_charges, dtype=dtype),
    }
