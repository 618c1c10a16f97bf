"""CPU oracle for the per-particle NON-LINEAR tracking methods.

TEST INFRASTRUCTURE ONLY -- the checker, never the product (see ``oracle/track_oracle.py``).

Functional restatement, in plain PyTorch CPU ops and in the reference's operation order, of
(desy-ml/cheetah @ 60d1053, paths relative to /root/reference):

* ``"drift_kick_drift"`` (Bmad-X) tracking
    cheetah/utils/bmadx.py:7-318 (coordinate conversions, offsets, exact drift, quadrupole
    coefficients, low-energy z correction), cheetah/accelerator/drift.py:106-154,
    quadrupole.py:168-251, dipole.py:183-370, transverse_deflecting_cavity.py:122-209
* ``"second_order"`` tracking
    cheetah/track_methods.py:80-281 (``base_ttensor``), cheetah/utils/autograd.py (the
    singularity-free compound functions), cheetah/accelerator/element.py:195-225 (apply),
    drift.py:67-83, quadrupole.py:112-143, dipole.py:396-466, sextupole.py:90-116

Parity pin: ``tests/test_oracle_golden.py`` checks these functions against the reference's own
golden pickles (``Drift/Quadrupole/Dipole/RBend/Sextupole_ParticleBeam_{drift_kick_drift,
second_order}.pkl``, ``TransverseDeflectingCavity_ParticleBeam_{active,inactive}.pkl``) and
against the Bmad-X fixtures ``tests/resources/bmadx/*.pt`` at the reference's tolerances
(1e-14 in float64), both converted to ``tests/golden/nonlinear.npz`` by ``oracle/make_golden.py``.
"""

from __future__ import annotations

import torch

from .track_oracle import (
    SPEED_OF_LIGHT,
    _get,
    _with,
    base_rmatrix,
    drift_map,
    relativistic_factors,
    rotation_map,
    si1mdiv,
    tilt_misalignment_maps,
)

# --------------------------------------------------------------------------------------
# Bmad-X helpers (cheetah/utils/bmadx.py)
# --------------------------------------------------------------------------------------


def cheetah_to_bmad_z_pz(tau, delta, ref_energy, mc2):
    """cheetah/utils/bmadx.py:7-30."""
    p0c = (ref_energy.square() - mc2.square()).sqrt()
    energy = ref_energy.unsqueeze(-1) + delta * p0c.unsqueeze(-1)
    p = (energy.square() - mc2.square()).sqrt()
    beta = p / energy
    z = -beta * tau
    pz = (p - p0c.unsqueeze(-1)) / p0c.unsqueeze(-1)
    return z, pz, p0c


def bmad_to_cheetah_z_pz(z, pz, p0c, mc2):
    """cheetah/utils/bmadx.py:33-55."""
    ref_energy = (p0c.square() + mc2.square()).sqrt()
    p = (1 + pz) * p0c.unsqueeze(-1)
    energy = (p.square() + mc2.square()).sqrt()
    beta = p / energy
    tau = -z / beta
    delta = (energy - ref_energy.unsqueeze(-1)) / p0c.unsqueeze(-1)
    return tau, delta, ref_energy


def offset_particle_set(x_offset, y_offset, tilt, x_lab, px_lab, y_lab, py_lab):
    """Lab -> element frame (cheetah/utils/bmadx.py:115-146)."""
    s, c = tilt.sin(), tilt.cos()
    x_int = x_lab - x_offset.unsqueeze(-1)
    y_int = y_lab - y_offset.unsqueeze(-1)
    x = x_int * c.unsqueeze(-1) + y_int * s.unsqueeze(-1)
    y = -x_int * s.unsqueeze(-1) + y_int * c.unsqueeze(-1)
    px = px_lab * c.unsqueeze(-1) + py_lab * s.unsqueeze(-1)
    py = -px_lab * s.unsqueeze(-1) + py_lab * c.unsqueeze(-1)
    return x, px, y, py


def offset_particle_unset(x_offset, y_offset, tilt, x_ele, px_ele, y_ele, py_ele):
    """Element -> lab frame (cheetah/utils/bmadx.py:149-180)."""
    s, c = tilt.sin(), tilt.cos()
    x_int = x_ele * c.unsqueeze(-1) - y_ele * s.unsqueeze(-1)
    y_int = x_ele * s.unsqueeze(-1) + y_ele * c.unsqueeze(-1)
    x = x_int + x_offset.unsqueeze(-1)
    y = y_int + y_offset.unsqueeze(-1)
    px = px_ele * c.unsqueeze(-1) - py_ele * s.unsqueeze(-1)
    py = px_ele * s.unsqueeze(-1) + py_ele * c.unsqueeze(-1)
    return x, px, y, py


def low_energy_z_correction(pz, p0c, mc2, ds):
    """cheetah/utils/bmadx.py:183-220."""
    beta = (
        (1 + pz) * p0c.unsqueeze(-1)
        / (((1 + pz) * p0c.unsqueeze(-1)).square() + mc2.square()).sqrt()
    )
    beta0 = p0c / (p0c.square() + mc2.square()).sqrt()
    e_tot = (p0c.square() + mc2.square()).sqrt()
    evaluation = mc2 * (beta0.unsqueeze(-1) * pz).square()
    dz = ds.unsqueeze(-1) * pz * (
        1
        - 3 * (pz * beta0.square().unsqueeze(-1)) / 2
        + pz.square()
        * beta0.square().unsqueeze(-1)
        * (2 * beta0.square().unsqueeze(-1) - (mc2 / e_tot.unsqueeze(-1)).square() / 2)
    ) * (mc2 / e_tot.unsqueeze(-1)).square() * (evaluation < 3e-7 * e_tot.unsqueeze(-1)) + (
        ds.unsqueeze(-1) * (beta - beta0.unsqueeze(-1)) / beta0.unsqueeze(-1)
    ) * (evaluation >= 3e-7 * e_tot.unsqueeze(-1))
    return dz


def calculate_quadrupole_coefficients(k1, length, rel_p):
    """cheetah/utils/bmadx.py:223-260 (sign of k1 flipped w.r.t. track_methods)."""
    kx = torch.complex(-k1, torch.zeros_like(k1)).sqrt()
    cx = (kx * length.unsqueeze(-1)).cos().real
    sx = ((kx * length.unsqueeze(-1) / torch.pi).sinc() * length.unsqueeze(-1)).real
    a11 = cx
    a12 = sx / rel_p
    a21 = k1 * sx * rel_p
    a22 = cx
    c1 = k1 * (-cx * sx + length.unsqueeze(-1)) / 4
    c2 = -k1 * sx.square() / (2 * rel_p)
    c3 = -(cx * sx + length.unsqueeze(-1)) / (4 * rel_p.square())
    return [[a11, a12], [a21, a22]], [c1, c2, c3]


def sqrt_one(x):
    """sqrt(1 + x) - 1 to machine precision (cheetah/utils/bmadx.py:263-268)."""
    return x / ((1 + x).sqrt() + 1)


def track_a_drift(length, x_in, px_in, y_in, py_in, z_in, pz_in, p0c, mc2):
    """Exact drift (cheetah/utils/bmadx.py:271-302)."""
    P = 1.0 + pz_in
    Px = px_in / P
    Py = py_in / P
    Pxy2 = Px.square() + Py.square()
    Pl = (1.0 - Pxy2).sqrt()
    dz = length.unsqueeze(-1) * (
        sqrt_one(
            (mc2.square() * (2 * pz_in + pz_in.square()))
            / ((p0c.unsqueeze(-1) * P).square() + mc2.square())
        )
        + sqrt_one(-Pxy2) / Pl
    )
    x_out = x_in + length.unsqueeze(-1) * Px / Pl
    y_out = y_in + length.unsqueeze(-1) * Py / Pl
    z_out = z_in + dz
    return x_out, y_out, z_out


def particle_rf_time(z, pz, p0c, mc2):
    """cheetah/utils/bmadx.py:305-313."""
    beta = (
        (1 + pz) * p0c.unsqueeze(-1)
        / (((1 + pz) * p0c.unsqueeze(-1)).square() + mc2.square()).sqrt()
    )
    return -z / (beta * SPEED_OF_LIGHT)


def _sinc(x):
    return (x / torch.pi).sinc()


def _cosc(x):
    return -0.5 * _sinc(x / 2).square()


def sqrta2minusbdiva(a, b):
    """(sqrt(a^2 + b) - a) / b, limit 1 / (2 a) (cheetah/utils/autograd.py:652-670)."""
    safe = torch.where(b != 0, b, torch.ones_like(b))
    return torch.where(b != 0, ((a.square() + b).sqrt() - a) / safe, (2.0 * a).reciprocal())


def _stack_outgoing(beam, x, px, y, py, tau, delta, ref_energy, length):
    x, px, y, py, tau, delta = torch.broadcast_tensors(x, px, y, py, tau, delta)
    particles = torch.stack([x, px, y, py, tau, delta, torch.ones_like(x)], dim=-1)
    return _with(beam, particles=particles, energy=ref_energy, s=beam["s"] + length)


def _columns(particles):
    return tuple(particles[..., i] for i in range(6))


def track_drift_dkd(el: dict, beam: dict) -> dict:
    """Drift, ``drift_kick_drift`` (cheetah/accelerator/drift.py:106-154)."""
    particles, mc2 = beam["particles"], beam["mass_eV"]
    length = _get(el, "length", particles)
    x, px, y, py, tau, delta = _columns(particles)
    z, pz, p0c = cheetah_to_bmad_z_pz(tau, delta, beam["energy"], mc2)
    x, y, z = track_a_drift(length, x, px, y, py, z, pz, p0c, mc2)
    tau, delta, ref_energy = bmad_to_cheetah_z_pz(z, pz, p0c, mc2)
    return _stack_outgoing(beam, x, px, y, py, tau, delta, ref_energy, length)


def track_quadrupole_dkd(el: dict, beam: dict) -> dict:
    """Quadrupole, ``drift_kick_drift`` (cheetah/accelerator/quadrupole.py:168-251)."""
    particles, mc2 = beam["particles"], beam["mass_eV"]
    length = _get(el, "length", particles)
    k1_el = _get(el, "k1", particles)
    tilt = _get(el, "tilt", particles)
    misalignment = el.get("misalignment")
    if misalignment is None:
        misalignment = particles.new_zeros(2)
    misalignment = misalignment.to(particles.dtype)
    num_steps = int(el.get("num_steps", 1))

    x, px, y, py, tau, delta = _columns(particles)
    z, pz, p0c = cheetah_to_bmad_z_pz(tau, delta, beam["energy"], mc2)
    x_offset, y_offset = misalignment[..., 0], misalignment[..., 1]
    step_length = length / num_steps
    b1 = k1_el * length

    x, px, y, py = offset_particle_set(x_offset, y_offset, tilt, x, px, y, py)
    for _ in range(num_steps):
        rel_p = 1 + pz
        k1 = b1.unsqueeze(-1) / (length.unsqueeze(-1) * rel_p)
        tx, dzx = calculate_quadrupole_coefficients(-k1, step_length, rel_p)
        ty, dzy = calculate_quadrupole_coefficients(k1, step_length, rel_p)
        z = (
            z
            + dzx[0] * x.square() + dzx[1] * x * px + dzx[2] * px.square()
            + dzy[0] * y.square() + dzy[1] * y * py + dzy[2] * py.square()
        )
        x_next = tx[0][0] * x + tx[0][1] * px
        px_next = tx[1][0] * x + tx[1][1] * px
        y_next = ty[0][0] * y + ty[0][1] * py
        py_next = ty[1][0] * y + ty[1][1] * py
        x, px, y, py = x_next, px_next, y_next, py_next
        z = z + low_energy_z_correction(pz, p0c, mc2, step_length)
    x, px, y, py = offset_particle_unset(x_offset, y_offset, tilt, x, px, y, py)
    pz, _ = torch.broadcast_tensors(pz, x)
    tau, delta, ref_energy = bmad_to_cheetah_z_pz(z, pz, p0c, mc2)
    return _stack_outgoing(beam, x, px, y, py, tau, delta, ref_energy, length)


def _dipole_parameters(el: dict, like: torch.Tensor) -> dict:
    """Dipole / RBend attributes with the constructor defaults (dipole.py:58-129,
    rbend.py:48-101: ``dipole_e = rbend_e + angle / 2``; ``*_exit`` default to the entrance)."""
    angle = _get(el, "angle", like)
    if el["type"] == "RBend":
        e1 = _get(el, "rbend_e1", like) + angle / 2
        e2 = _get(el, "rbend_e2", like) + angle / 2
    else:
        e1 = _get(el, "dipole_e1", like)
        e2 = _get(el, "dipole_e2", like)
    fint = _get(el, "fringe_integral", like)
    gap = _get(el, "gap", like)
    return {
        "length": _get(el, "length", like), "angle": angle, "k1": _get(el, "k1", like),
        "e1": e1, "e2": e2, "tilt": _get(el, "tilt", like),
        "fint": fint,
        "fint_exit": _get(el, "fringe_integral_exit", like) if "fringe_integral_exit" in el else fint,
        "gap": gap,
        "gap_exit": _get(el, "gap_exit", like) if "gap_exit" in el else gap,
    }


def _bmadx_fringe_linear(d: dict, location: str, x, px, y, py):
    """cheetah/accelerator/dipole.py:338-370."""
    g = d["angle"] / d["length"]
    e = d["e1"] if location == "entrance" else d["e2"]
    f_int = d["fint"] if location == "entrance" else d["fint_exit"]
    h_gap = 0.5 * (d["gap"] if location == "entrance" else d["gap_exit"])
    hx = g * e.tan()
    hy = -g * (e - 2 * f_int * h_gap * g * (1 + e.sin().square()) / e.cos()).tan()
    return px + x * hx.unsqueeze(-1), py + y * hy.unsqueeze(-1)


def _bmadx_body(d: dict, x, px, y, py, z, pz, p0c, mc2):
    """Sector-bend body (cheetah/accelerator/dipole.py:244-336)."""
    angle, length = d["angle"], d["length"]
    px_norm = ((1 + pz).square() - py.square()).sqrt()
    phi1 = (px / px_norm).arcsin()
    g = angle / length
    gp = g.unsqueeze(-1) / px_norm

    alpha = (
        2 * (1 + g.unsqueeze(-1) * x) * (angle.unsqueeze(-1) + phi1).sin()
        * length.unsqueeze(-1) * _sinc(angle).unsqueeze(-1)
        - gp * ((1 + g.unsqueeze(-1) * x) * length.unsqueeze(-1) * _sinc(angle).unsqueeze(-1)).square()
    )
    x2_t1 = x * angle.cos().unsqueeze(-1) + length.unsqueeze(-1).square() * g.unsqueeze(
        -1
    ) * _cosc(angle.unsqueeze(-1))
    x2_t2 = ((angle.unsqueeze(-1) + phi1).cos().square() + gp * alpha).sqrt()
    x2_t3 = (angle.unsqueeze(-1) + phi1).cos()
    c1 = x2_t1 + alpha / (x2_t2 + x2_t3)
    c2 = x2_t1 + alpha * sqrta2minusbdiva(x2_t3, gp * alpha)
    temp = (angle.unsqueeze(-1) + phi1).abs()
    x2 = c1.where(temp < torch.pi / 2, c2)

    Lcu = x2 - length.square().unsqueeze(-1) * g.unsqueeze(-1) * _cosc(angle.unsqueeze(-1)) - x * (
        angle.cos().unsqueeze(-1)
    )
    Lcv = -length.unsqueeze(-1) * _sinc(angle.unsqueeze(-1)) - x * angle.sin().unsqueeze(-1)
    theta_p = 2 * (angle.unsqueeze(-1) + phi1 - torch.pi / 2 - torch.arctan2(Lcv, Lcu))
    Lc = (Lcu.square() + Lcv.square()).sqrt()
    Lp = Lc / _sinc(theta_p / 2)

    P = p0c.unsqueeze(-1) * (1 + pz)
    E = (P.square() + mc2.square()).sqrt()
    E0 = (p0c.square() + mc2.square()).sqrt()
    beta = P / E
    beta0 = p0c / E0

    x_f = x2
    px_f = px_norm * (angle.unsqueeze(-1) + phi1 - theta_p).sin()
    y_f = y + py * Lp / px_norm
    z_f = z + (beta * length.unsqueeze(-1) / beta0.unsqueeze(-1)) - ((1 + pz) * Lp / px_norm)
    return x_f, px_f, y_f, py, z_f, pz


def track_dipole_dkd(el: dict, beam: dict) -> dict:
    """Dipole / RBend, ``drift_kick_drift`` (cheetah/accelerator/dipole.py:183-242)."""
    particles, mc2 = beam["particles"], beam["mass_eV"]
    d = _dipole_parameters(el, particles)
    fringe_at = el.get("fringe_at", "both")
    zero = particles.new_zeros(())
    x, px, y, py, tau, delta = _columns(particles)
    z, pz, p0c = cheetah_to_bmad_z_pz(tau, delta, beam["energy"], mc2)
    x, px, y, py = offset_particle_set(zero, zero, d["tilt"], x, px, y, py)
    if fringe_at in ("entrance", "both"):
        px, py = _bmadx_fringe_linear(d, "entrance", x, px, y, py)
    x, px, y, py, z, pz = _bmadx_body(d, x, px, y, py, z, pz, p0c, mc2)
    if fringe_at in ("exit", "both"):
        px, py = _bmadx_fringe_linear(d, "exit", x, px, y, py)
    x, px, y, py = offset_particle_unset(zero, zero, d["tilt"], x, px, y, py)
    tau, delta, ref_energy = bmad_to_cheetah_z_pz(z, pz, p0c, mc2)
    return _stack_outgoing(beam, x, px, y, py, tau, delta, ref_energy, d["length"])


def track_tdc(el: dict, beam: dict) -> dict:
    """TransverseDeflectingCavity (cheetah/accelerator/transverse_deflecting_cavity.py:122-209)."""
    particles, mc2 = beam["particles"], beam["mass_eV"]
    length = _get(el, "length", particles)
    voltage_el = _get(el, "voltage", particles)
    phase_el = _get(el, "phase", particles)
    frequency = _get(el, "frequency", particles)
    tilt = _get(el, "tilt", particles)
    misalignment = el.get("misalignment")
    if misalignment is None:
        misalignment = particles.new_zeros(2)
    misalignment = misalignment.to(particles.dtype)

    x, px, y, py, tau, delta = _columns(particles)
    z, pz, p0c = cheetah_to_bmad_z_pz(tau, delta, beam["energy"], mc2)
    x_offset, y_offset = misalignment[..., 0], misalignment[..., 1]
    x, px, y, py = offset_particle_set(x_offset, y_offset, tilt, x, px, y, py)
    x, y, z = track_a_drift(length / 2, x, px, y, py, z, pz, p0c, mc2)

    voltage = voltage_el * -1 * beam["num_elementary_charges"] / p0c
    k_rf = 2 * torch.pi * frequency / SPEED_OF_LIGHT
    phase = 2 * torch.pi * (
        phase_el.unsqueeze(-1) - particle_rf_time(z, pz, p0c, mc2) * frequency.unsqueeze(-1)
    )
    px = px + voltage.unsqueeze(-1) * phase.sin()
    beta_old = (
        (1 + pz) * p0c.unsqueeze(-1)
        / (((1 + pz) * p0c.unsqueeze(-1)).square() + mc2.square()).sqrt()
    )
    E_old = (1 + pz) * p0c.unsqueeze(-1) / beta_old
    E_new = E_old + voltage.unsqueeze(-1) * phase.cos() * k_rf.unsqueeze(-1) * x * p0c.unsqueeze(-1)
    pc = (E_new.square() - mc2.square()).sqrt()
    beta = pc / E_new
    pz = (pc - p0c.unsqueeze(-1)) / p0c.unsqueeze(-1)
    z = z * beta / beta_old

    x, y, z = track_a_drift(length / 2, x, px, y, py, z, pz, p0c, mc2)
    x, px, y, py = offset_particle_unset(x_offset, y_offset, tilt, x, px, y, py)
    tau, delta, ref_energy = bmad_to_cheetah_z_pz(z, pz, p0c, mc2)
    return _stack_outgoing(beam, x, px, y, py, tau, delta, ref_energy, length)


# --------------------------------------------------------------------------------------
# Second-order maps (cheetah/track_methods.py:80-281, cheetah/utils/autograd.py)
# --------------------------------------------------------------------------------------


def _csqrt(x):
    return torch.complex(x, torch.zeros_like(x)).sqrt()


def _safe(x):
    return torch.where(x != 0, x, torch.ones_like(x))


def sicos1mdiv(x):
    """(1 - si(sqrt x) cos(sqrt x)) / x, limit 1/6 as coded (autograd.py:149-174)."""
    r = _csqrt(x)
    cx, sx = r.cos().real, (r / torch.pi).sinc().real
    return torch.where(x != 0, (1.0 - sx * cx) / _safe(x), torch.full_like(x, 1.0 / 6.0))


def sipsicos3mdiv(x):
    """(3 - 4 si + si cos) / (2 x), limit 0 (autograd.py:209-235)."""
    r = _csqrt(x)
    cx, sx = r.cos().real, (r / torch.pi).sinc().real
    return torch.where(x != 0, (3.0 - 4.0 * sx + sx * cx) / (2.0 * _safe(x)), torch.zeros_like(x))


def cossqrtmcosdivdiff(a, b):
    """(cos sqrt b - cos sqrt a) / (a - b), limit si(sqrt a) / 2 (autograd.py:361-388)."""
    ra, rb = _csqrt(a), _csqrt(b)
    sa, ca, cb = (ra / torch.pi).sinc().real, ra.cos().real, rb.cos().real
    return torch.where(a != b, (cb - ca) / _safe(a - b), 0.5 * sa)


def simsidivdiff(a, b):
    """(si sqrt a - si sqrt b) / (b - a) (autograd.py:433-461)."""
    ra, rb = _csqrt(a), _csqrt(b)
    sa, sb, cb = (ra / torch.pi).sinc().real, (rb / torch.pi).sinc().real, rb.cos().real
    limit = torch.where(b != 0, 0.5 * (sb - cb) / _safe(b), torch.full_like(b, 1.0 / 6.0))
    return torch.where(a != b, (sa - sb) / _safe(b - a), limit)


def si2msi2divdiff(a, b):
    """(si^2 sqrt b - si^2 sqrt a) / (a - b) (autograd.py:546-579)."""
    ra, rb = _csqrt(a), _csqrt(b)
    sa, sb, cb = (ra / torch.pi).sinc().real, (rb / torch.pi).sinc().real, rb.cos().real
    limit = torch.where(
        b != 0, (1.0 - cb.square() - b * sb * cb) / _safe(b).square(), torch.full_like(b, 1.0 / 3.0)
    )
    return torch.where(a != b, (sb.square() - sa.square()) / _safe(a - b), limit)


def base_ttensor(length, k1, k2, hx, energy, mass_eV) -> torch.Tensor:
    """Second-order map of a combined-function body (cheetah/track_methods.py:80-281)."""
    _, igamma2, beta = relativistic_factors(energy, mass_eV)
    kx2 = k1 + hx.square()
    ky2 = -k1
    kx, ky = _csqrt(kx2), _csqrt(ky2)
    cx = (kx * length).cos().real
    cy = (ky * length).cos().real
    sx = ((kx * length / torch.pi).sinc() * length).real
    sy = ((ky * length / torch.pi).sinc() * length).real
    r = (0.5 * kx * length / torch.pi).sinc()
    dx = 0.5 * length.square() * r.square().real

    fx = length.pow(3) * si1mdiv(kx2 * length.square())
    f2y = length.pow(3) * sicos1mdiv(ky2 * length.square())
    j1 = fx
    j2 = length.pow(3) * sipsicos3mdiv(kx2 * length.square())
    j3 = torch.where(
        kx2 != 0,
        (15.0 * length - 22.5 * sx + 9.0 * sx * cx - 1.5 * sx * cx.square() + kx2 * sx.pow(3))
        / (6.0 * _safe(kx2).pow(3)),
        length.pow(7) / 56.0,
    )
    j_denominator = kx2 - 4.0 * ky2
    jc = length.square() * cossqrtmcosdivdiff(kx2 * length.square(), ky2 * length.square())
    js = length.pow(3) * simsidivdiff(kx2 * length.square(), ky2 * length.square())
    jd = length.pow(4) * si2msi2divdiff(kx2 * length.square(), ky2 * length.square())
    jf = torch.where(j_denominator != 0, (f2y - fx) / _safe(j_denominator), length.pow(5) / 120.0)
    khk = k2 + 2.0 * hx * k1

    vector_shape = torch.broadcast_shapes(length.shape, k1.shape, k2.shape, hx.shape, energy.shape)
    T = length.new_zeros((7, 7, 7)).repeat(*vector_shape, 1, 1, 1)
    T[..., 0, 0, 0] = -khk * (sx.square() + dx) / 6.0 - 0.5 * hx * kx2 * sx.square()
    T[..., 0, 0, 1] = 2.0 * (-khk * sx * dx / 6.0 + 0.5 * hx * sx * cx)
    T[..., 0, 1, 1] = -khk * dx.square() / 6.0 + 0.5 * hx * dx * cx
    T[..., 0, 0, 5] = 2.0 * (
        -hx / 12.0 / beta * khk * (3.0 * sx * j1 - dx.square())
        + 0.5 * hx.square() / beta * sx.square()
        + 0.25 / beta * k1 * length * sx
    )
    T[..., 0, 1, 5] = 2.0 * (
        -hx / 12.0 / beta * khk * (sx * dx.square() - 2.0 * cx * j2)
        + 0.25 * hx.square() / beta * (sx * dx + cx * j1)
        - 0.25 / beta * (sx + length * cx)
    )
    T[..., 0, 5, 5] = (
        -(hx.square()) / 6.0 / beta.square() * khk * (dx.square() * dx - 2.0 * sx * j2)
        + 0.5 * hx.pow(3) / beta.square() * sx * j1
        - 0.5 * hx / beta.square() * length * sx
        - 0.5 * hx / (beta.square()) * igamma2 * dx
    )
    T[..., 0, 2, 2] = k1 * k2 * jd + 0.5 * (k2 + hx * k1) * dx
    T[..., 0, 2, 3] = 2.0 * (0.5 * k2 * js)
    T[..., 0, 3, 3] = k2 * jd - 0.5 * hx * dx
    T[..., 1, 0, 0] = -khk * sx * (1.0 + 2.0 * cx) / 6.0
    T[..., 1, 0, 1] = -2.0 * khk * dx * (1.0 + 2.0 * cx) / 6.0
    T[..., 1, 1, 1] = -khk * sx * dx / 3.0 - 0.5 * hx * sx
    T[..., 1, 0, 5] = 2.0 * (
        -hx / 12.0 / beta * khk * (3.0 * cx * j1 + sx * dx) - 0.25 / beta * k1 * (sx - length * cx)
    )
    T[..., 1, 1, 5] = 2.0 * (
        -hx / 12.0 / beta * khk * (3.0 * sx * j1 + dx.square()) + 0.25 / beta * k1 * length * sx
    )
    T[..., 1, 5, 5] = (
        -(hx.square()) / 6.0 / beta.square() * khk * (sx * dx.square() - 2.0 * cx * j2)
        - 0.5 * hx / beta.square() * k1 * (cx * j1 - sx * dx)
        - 0.5 * hx / beta.square() * igamma2 * sx
    )
    T[..., 1, 2, 2] = k1 * k2 * js + 0.5 * (k2 + hx * k1) * sx
    T[..., 1, 2, 3] = 2.0 * (0.5 * k2 * jc)
    T[..., 1, 3, 3] = k2 * js - 0.5 * hx * sx
    T[..., 2, 0, 2] = 2.0 * (0.5 * k2 * (cy * jc - 2.0 * k1 * sy * js) + 0.5 * hx * k1 * sx * sy)
    T[..., 2, 0, 3] = 2.0 * (0.5 * k2 * (sy * jc - 2.0 * cy * js) + 0.5 * hx * sx * cy)
    T[..., 2, 1, 2] = 2.0 * (0.5 * k2 * (cy * js - 2.0 * k1 * sy * jd) + 0.5 * hx * k1 * dx * sy)
    T[..., 2, 1, 3] = 2.0 * (0.5 * k2 * (sy * js - 2.0 * cy * jd) + 0.5 * hx * dx * cy)
    T[..., 2, 2, 5] = 2.0 * (
        0.5 * hx / beta * k2 * (cy * jd - 2.0 * k1 * sy * jf)
        + 0.5 * hx.square() / beta * k1 * j1 * sy
        - 0.25 / beta * k1 * length * sy
    )
    T[..., 2, 3, 5] = 2.0 * (
        0.5 * hx / beta * k2 * (sy * jd - 2.0 * cy * jf)
        + 0.5 * hx.square() / beta * j1 * cy
        - 0.25 / beta * (sy + length * cy)
    )
    T[..., 3, 0, 2] = 2.0 * (
        0.5 * k1 * k2 * (2.0 * cy * js - sy * jc) + 0.5 * (k2 + hx * k1) * sx * cy
    )
    T[..., 3, 0, 3] = 2.0 * (
        0.5 * k2 * (2.0 * k1 * sy * js - cy * jc) + 0.5 * (k2 + hx * k1) * sx * sy
    )
    T[..., 3, 1, 2] = 2.0 * (
        0.5 * k1 * k2 * (2.0 * cy * jd - sy * js) + 0.5 * (k2 + hx * k1) * dx * cy
    )
    T[..., 3, 1, 3] = 2.0 * (
        0.5 * k2 * (2.0 * k1 * sy * jd - cy * js) + 0.5 * (k2 + hx * k1) * dx * sy
    )
    T[..., 3, 2, 5] = 2.0 * (
        0.5 * hx / beta * k1 * k2 * (2.0 * cy * jf - sy * jd)
        + 0.5 * hx / beta * (k2 + hx * k1) * j1 * cy
        + 0.25 / beta * k1 * (sy - length * cy)
    )
    T[..., 3, 3, 5] = 2.0 * (
        0.5 * hx / beta * k2 * (2.0 * k1 * sy * jf - cy * jd)
        + 0.5 * hx / beta * (k2 + hx * k1) * j1 * sy
        - 0.25 / beta * k1 * length * sy
    )
    T[..., 4, 0, 0] = -(
        hx / 12.0 / beta * khk * (sx * dx + 3.0 * j1) - 0.25 / beta * k1 * (length - sx * cx)
    )
    T[..., 4, 0, 1] = -2.0 * (hx / 12.0 / beta * khk * dx.square() + 0.25 / beta * k1 * sx.square())
    T[..., 4, 1, 1] = -(
        hx / 6.0 / beta * khk * j2 - 0.5 / beta * sx - 0.25 / beta * k1 * (j1 - sx * dx)
    )
    T[..., 4, 0, 5] = -2.0 * (
        hx.square() / 12.0 / beta.square() * khk * (3.0 * dx * j1 - 4.0 * j2)
        + 0.25 * hx / beta.square() * k1 * j1 * (1.0 + cx)
        + 0.5 * hx / beta.square() * igamma2 * sx
    )
    T[..., 4, 1, 5] = -2.0 * (
        hx.square() / 12.0 / beta.square() * khk * (dx * dx.square() - 2.0 * sx * j2)
        + 0.25 * hx / beta.square() * k1 * sx * j1
        + 0.5 * hx / beta.square() * igamma2 * dx
    )
    T[..., 4, 5, 5] = -(
        hx.pow(3) / 6.0 / beta.pow(3) * khk * (3.0 * j3 - 2.0 * dx * j2)
        + hx.square() / 6.0 / beta.pow(3) * k1 * (sx * dx.square() - j2 * (1.0 + 2.0 * cx))
        + 1.5 / beta.pow(3) * igamma2 * (hx.square() * j1 - length)
    )
    T[..., 4, 2, 2] = -(
        -hx / beta * k1 * k2 * jf
        - 0.5 * hx / beta * (k2 + hx * k1) * j1
        + 0.25 / beta * k1 * (length - cy * sy)
    )
    T[..., 4, 2, 3] = -2.0 * (-0.5 * hx / beta * k2 * jd - 0.25 / beta * k1 * sy.square())
    T[..., 4, 3, 3] = -(
        -hx / beta * k2 * jf + 0.5 * hx.square() / beta * j1 - 0.25 / beta * (length + cy * sy)
    )
    return T


def _dipole_edge_map(hx, e, fint, gap) -> torch.Tensor:
    """Entrance / exit face map (dipole.py:430-466; the exit face uses ``gap``, not
    ``gap_exit``, like the reference)."""
    sec_e = e.cos().reciprocal()
    phi = fint * hx * gap * sec_e * (1 + e.sin().square())
    tm = torch.eye(7, dtype=hx.dtype).repeat(*phi.shape, 1, 1)
    tm[..., 1, 0] = hx * e.tan()
    tm[..., 3, 2] = -hx * (e - phi).tan()
    return tm


def second_order_map(el: dict, energy: torch.Tensor, mass_eV) -> torch.Tensor:
    """``second_order_transfer_map`` of Drift / Quadrupole / Sextupole / Dipole / RBend
    (drift.py:67-83, quadrupole.py:112-143, sextupole.py:90-116, dipole.py:396-428)."""
    kind = el["type"]
    like = energy
    zero = like.new_zeros(())
    length = _get(el, "length", like)
    if kind == "Drift":
        T = base_ttensor(length, zero, zero, zero, energy, mass_eV)
        T[..., :, 6, :] = drift_map(length, energy, mass_eV)
        return T
    if kind in ("Quadrupole", "Sextupole"):
        k1 = _get(el, "k1", like) if kind == "Quadrupole" else zero
        k2 = _get(el, "k2", like) if kind == "Sextupole" else zero
        T = base_ttensor(length, k1, k2, zero, energy, mass_eV)
        if kind == "Quadrupole":
            T[..., :, 6, :] = base_rmatrix(length, k1, zero, energy, mass_eV)
        else:
            T[..., :, 6, :] = drift_map(length, energy, mass_eV)
        misalignment = el.get("misalignment")
        if misalignment is None:
            misalignment = like.new_zeros(2)
        R_entry, R_exit = tilt_misalignment_maps(_get(el, "tilt", like), misalignment.to(like.dtype))
        return torch.einsum("...ij,...jkl,...kn,...lm->...inm", R_exit, T, R_entry, R_entry)
    if kind in ("Dipole", "RBend"):
        d = _dipole_parameters(el, like)
        hx = d["angle"] / d["length"]
        R_enter = _dipole_edge_map(hx, d["e1"], d["fint"], d["gap"])
        R_exit = _dipole_edge_map(hx, d["e2"], d["fint_exit"], d["gap"])
        T = base_ttensor(d["length"], d["k1"], zero, hx, energy, mass_eV)
        T[..., :, 6, :] = base_rmatrix(d["length"], d["k1"], hx, energy, mass_eV)
        T = torch.einsum("...ij,...jkl,...kn,...lm->...inm", R_exit, T, R_enter, R_enter)
        rotation = rotation_map(d["tilt"])
        return torch.einsum("...ji,...jkl,...kn,...lm->...inm", rotation, T, rotation, rotation)
    raise NotImplementedError(f"oracle has no second-order map for element type {kind}")


def track_second_order(el: dict, beam: dict) -> dict:
    """``Element._track_second_order`` (cheetah/accelerator/element.py:195-225)."""
    particles = beam["particles"]
    T = second_order_map(el, beam["energy"], beam["mass_eV"])
    out = torch.einsum("...ijk,...j,...k->...i", T.unsqueeze(-4), particles, particles)
    return _with(beam, particles=out, s=beam["s"] + _get(el, "length", particles))


def track_nonlinear(el: dict, beam: dict) -> dict:
    """Dispatch one non-skippable Drift / Quadrupole / Dipole / RBend / Sextupole / TDC."""
    kind = el["type"]
    method = el.get("tracking_method", "drift_kick_drift" if kind == "TransverseDeflectingCavity" else "linear")
    if kind == "TransverseDeflectingCavity":
        return track_tdc(el, beam)
    if method == "second_order":
        return track_second_order(el, beam)
    if method == "drift_kick_drift":
        if kind == "Drift":
            return track_drift_dkd(el, beam)
        if kind == "Quadrupole":
            return track_quadrupole_dkd(el, beam)
        if kind in ("Dipole", "RBend"):
            return track_dipole_dkd(el, beam)
    raise NotImplementedError(f"oracle: {kind} with tracking_method={method!r}")
