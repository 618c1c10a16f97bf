"""Randomised differential test: seeded random lattices with randomly vectorised parameters and
beams, product (CUDA) against the float64 oracle -- output values AND the broadcast shapes of
particles, survival probabilities, ``s`` and energy (NumPy-style vector dims)."""

import pytest
import torch

from oracle import track_oracle as oracle

from . import golden_utils as gu

pytestmark = pytest.mark.gpu
DEVICE = "cuda"
SHAPES = [(), (), (), (3,), (2, 1), (2, 3)]


def _vector(generator, low, high, shape):
    return low + (high - low) * torch.rand(shape, generator=generator, dtype=torch.float64)


def random_lattice(seed: int) -> tuple[list, tuple]:
    g = torch.Generator().manual_seed(seed)
    pick = lambda: SHAPES[int(torch.randint(0, len(SHAPES), (1,), generator=g))]  # noqa: E731
    lattice = []
    n_elements = int(torch.randint(3, 12, (1,), generator=g))
    for i in range(n_elements):
        kind = ["Drift", "Quadrupole", "HorizontalCorrector", "VerticalCorrector", "Dipole",
                "Solenoid", "Marker", "Aperture", "Cavity", "Sextupole", "Undulator",
                "CombinedCorrector"][int(torch.randint(0, 12, (1,), generator=g))]
        el = {"type": kind, "name": f"e{i}"}
        if kind not in ("Marker", "Aperture"):
            el["length"] = _vector(g, 0.1, 1.0, pick())
        if kind == "Quadrupole":
            el["k1"] = _vector(g, -8.0, 8.0, pick())
            el["tilt"] = _vector(g, -0.3, 0.3, pick())
            el["misalignment"] = _vector(g, -1e-4, 1e-4, (*pick(), 2))
        elif kind in ("HorizontalCorrector", "VerticalCorrector"):
            el["angle"] = _vector(g, -1e-4, 1e-4, pick())
        elif kind == "CombinedCorrector":
            el["horizontal_angle"] = _vector(g, -1e-4, 1e-4, pick())
            el["vertical_angle"] = _vector(g, -1e-4, 1e-4, pick())
        elif kind == "Dipole":
            el["angle"] = _vector(g, -0.3, 0.3, pick())
            el["k1"] = _vector(g, -1.0, 1.0, pick())
            el["dipole_e1"] = _vector(g, -0.1, 0.1, pick())
            el["dipole_e2"] = _vector(g, -0.1, 0.1, pick())
            el["tilt"] = _vector(g, -0.2, 0.2, pick())
            el["fringe_integral"] = _vector(g, 0.0, 0.5, ())
            el["gap"] = _vector(g, 0.0, 0.05, ())
        elif kind == "Solenoid":
            el["k"] = _vector(g, -2.0, 2.0, pick())
            el["misalignment"] = _vector(g, -1e-4, 1e-4, (2,))
        elif kind == "Aperture":
            el["x_max"] = _vector(g, 1.5e-4, 6e-4, pick())
            el["y_max"] = _vector(g, 1.5e-4, 6e-4, pick())
            el["shape"] = ["rectangular", "elliptical"][int(torch.randint(0, 2, (1,), generator=g))]
            el["is_active"] = True
        elif kind == "Sextupole":
            el["k2"] = _vector(g, -5.0, 5.0, pick())
            el["tracking_method"] = "linear"
        elif kind == "Undulator":
            el["period"] = _vector(g, 0.02, 0.1, ())
            el["kx"] = _vector(g, 0.0, 2.0, pick())
            el["ky"] = _vector(g, 0.0, 2.0, ())
        lattice.append(el)
    beam_shape = [(), (), (3,), (2, 3)][int(torch.randint(0, 4, (1,), generator=g))]
    return lattice, beam_shape


@pytest.mark.parametrize("seed", range(40))
def test_random_vectorised_lattices(seed):
    lattice, beam_shape = random_lattice(seed)
    g = torch.Generator().manual_seed(1000 + seed)
    n = 257
    sigma = torch.tensor([2e-4, 3e-5, 2e-4, 3e-5, 1e-4, 1e-3], dtype=torch.float64)
    particles = torch.randn((*beam_shape, n, 7), generator=g, dtype=torch.float64)
    particles[..., :6] *= sigma
    particles[..., 6] = 1.0
    energy_shape = [(), (), (3,)][seed % 3] if beam_shape in ((), (3,), (2, 3)) else ()
    energy = _vector(g, 5e7, 2e8, energy_shape)
    beam = oracle.make_beam(particles, energy)
    beam["survival_probabilities"] = (torch.rand(n, generator=g, dtype=torch.float64) > 0.1).double()
    expected = oracle.track(lattice, beam)

    segment = gu.product_segment(lattice, DEVICE, torch.float64)
    out = segment.track(gu.product_beam(beam, DEVICE, torch.float64))
    assert tuple(out.particles.shape) == tuple(expected["particles"].shape), (
        out.particles.shape, expected["particles"].shape)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < 1e-9
    assert tuple(out.survival_probabilities.shape) == tuple(expected["survival_probabilities"].shape)
    assert torch.equal(out.survival_probabilities.cpu(), expected["survival_probabilities"])
    assert tuple(out.s.shape) == tuple(expected["s"].shape)
    assert torch.allclose(out.s.cpu(), expected["s"], rtol=1e-12)
    assert tuple(out.energy.shape) == tuple(expected["energy"].shape)


def random_mixed_lattice(seed: int) -> tuple[list, tuple]:
    """Linear, drift_kick_drift and second_order elements, markers and apertures."""
    g = torch.Generator().manual_seed(7000 + seed)
    pick = lambda: SHAPES[int(torch.randint(0, len(SHAPES), (1,), generator=g))]  # noqa: E731
    choice = lambda options: options[int(torch.randint(0, len(options), (1,), generator=g))]  # noqa: E731
    lattice = []
    for i in range(int(torch.randint(3, 10, (1,), generator=g))):
        kind = choice(["Drift", "Drift", "Quadrupole", "Quadrupole", "Dipole", "Sextupole",
                       "TransverseDeflectingCavity", "Marker", "Aperture", "HorizontalCorrector"])
        el = {"type": kind, "name": f"e{i}"}
        if kind not in ("Marker", "Aperture"):
            el["length"] = _vector(g, 0.1, 0.8, pick())
        if kind == "Drift":
            el["tracking_method"] = choice(["linear", "drift_kick_drift", "second_order"])
        elif kind == "Quadrupole":
            el["k1"] = _vector(g, -6.0, 6.0, pick())
            el["tilt"] = _vector(g, -0.3, 0.3, pick())
            el["misalignment"] = _vector(g, -1e-4, 1e-4, (*pick(), 2))
            el["tracking_method"] = choice(["linear", "drift_kick_drift", "second_order"])
            el["num_steps"] = int(torch.randint(1, 5, (1,), generator=g))
        elif kind == "Dipole":
            el["angle"] = _vector(g, 0.05, 0.3, pick())
            el["dipole_e1"] = _vector(g, -0.1, 0.1, pick())
            el["dipole_e2"] = _vector(g, -0.1, 0.1, ())
            el["tilt"] = _vector(g, -0.2, 0.2, pick())
            el["fringe_integral"] = _vector(g, 0.0, 0.5, ())
            el["gap"] = _vector(g, 0.0, 0.05, ())
            el["tracking_method"] = choice(["linear", "drift_kick_drift", "second_order"])
            el["fringe_at"] = choice(["both", "entrance", "exit", "neither"])
        elif kind == "Sextupole":
            el["k2"] = _vector(g, -30.0, 30.0, pick())
            el["tilt"] = _vector(g, -0.3, 0.3, ())
            el["misalignment"] = _vector(g, -1e-4, 1e-4, (2,))
            el["tracking_method"] = choice(["linear", "second_order"])
        elif kind == "TransverseDeflectingCavity":
            el["voltage"] = _vector(g, -5e6, 5e6, pick())
            el["phase"] = _vector(g, 0.0, 1.0, pick())
            el["frequency"] = _vector(g, 1e9, 3e9, ())
            el["tilt"] = _vector(g, -0.2, 0.2, ())
            el["misalignment"] = _vector(g, -1e-4, 1e-4, (2,))
        elif kind == "HorizontalCorrector":
            el["angle"] = _vector(g, -1e-4, 1e-4, pick())
        elif kind == "Aperture":
            el["x_max"] = _vector(g, 2e-4, 8e-4, pick())
            el["y_max"] = _vector(g, 2e-4, 8e-4, ())
            el["shape"] = choice(["rectangular", "elliptical"])
            el["is_active"] = True
        lattice.append(el)
    beam_shape = choice([(), (), (3,), (2, 3)])
    return lattice, beam_shape


@pytest.mark.parametrize("seed", range(40))
def test_random_lattices_with_nonlinear_elements(seed):
    lattice, beam_shape = random_mixed_lattice(seed)
    g = torch.Generator().manual_seed(9000 + seed)
    n = 193
    sigma = torch.tensor([2e-4, 3e-5, 2e-4, 3e-5, 1e-4, 1e-3], dtype=torch.float64)
    particles = torch.randn((*beam_shape, n, 7), generator=g, dtype=torch.float64)
    particles[..., :6] *= sigma
    particles[..., 6] = 1.0
    energy = _vector(g, 5e7, 2e8, ())
    beam = oracle.make_beam(particles, energy)
    expected = oracle.track(lattice, beam)
    if not torch.isfinite(expected["particles"]).all():
        pytest.skip("random optics blew the beam up (transverse momentum > total momentum)")
    out = gu.product_segment(lattice, DEVICE, torch.float64).track(
        gu.product_beam(beam, DEVICE, torch.float64))
    assert tuple(out.particles.shape) == tuple(expected["particles"].shape), (
        out.particles.shape, expected["particles"].shape)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < 1e-9
    assert tuple(out.survival_probabilities.shape) == tuple(expected["survival_probabilities"].shape)
    assert torch.equal(out.survival_probabilities.cpu(), expected["survival_probabilities"])
    assert tuple(out.s.shape) == tuple(expected["s"].shape)
    assert torch.allclose(out.s.cpu(), expected["s"], rtol=1e-12)
    assert torch.allclose(out.energy.cpu(), expected["energy"], rtol=1e-13)


@pytest.mark.parametrize("seed", range(12))
def test_random_space_charge_lattices(seed):
    """FODO-like lines with SpaceChargeKicks and randomly vectorised beams / charges / lengths /
    quadrupole strengths (the fused kick + linear section + next-moments path included)."""
    g = torch.Generator().manual_seed(300 + seed)
    choice = lambda options: options[int(torch.randint(0, len(options), (1,), generator=g))]  # noqa: E731
    batch = choice([(), (2,), (3,)])
    vec = lambda low, high: _vector(g, low, high, choice([(), batch]))  # noqa: E731
    grid = choice([(8, 8, 8), (16, 16, 16), (8, 16, 8)])
    lattice = []
    for i in range(int(torch.randint(1, 4, (1,), generator=g))):
        lattice += [
            {"type": "Quadrupole", "name": f"q{i}", "length": _vector(g, 0.1, 0.3, ()),
             "k1": vec(-4.0, 4.0)},
            {"type": "Drift", "name": f"d{i}a", "length": vec(0.2, 0.6)},
            {"type": "SpaceChargeKick", "name": f"sc{i}", "effect_length": vec(0.5, 1.5),
             "grid_shape": list(grid)},
        ]
        if choice([True, False]):
            lattice.append({"type": "Drift", "name": f"d{i}b", "length": _vector(g, 0.2, 0.6, ())})
        if choice([True, False, False]):
            lattice.append({"type": "Aperture", "name": f"a{i}", "x_max": _vector(g, 3e-4, 6e-4, ()),
                            "y_max": _vector(g, 3e-4, 6e-4, ()), "shape": "rectangular",
                            "is_active": True})
    n = 3000
    beam_shape = choice([(), batch])
    sigma = torch.tensor([2e-4, 3e-5, 2e-4, 3e-5, 1e-4, 1e-3], dtype=torch.float64)
    particles = torch.randn((*beam_shape, n, 7), generator=g, dtype=torch.float64)
    particles[..., :6] *= sigma
    particles[..., 6] = 1.0
    charge_shape = choice([(), batch])
    charges = -_vector(g, 1e-10, 1e-9, (*charge_shape, 1)).expand(*charge_shape, n) / n
    beam = oracle.make_beam(particles, torch.tensor(5e7, dtype=torch.float64),
                            particle_charges=charges.contiguous())
    expected = oracle.track(lattice, beam)
    out = gu.product_segment(lattice, DEVICE, torch.float64).track(
        gu.product_beam(beam, DEVICE, torch.float64))
    assert tuple(out.particles.shape) == tuple(expected["particles"].shape), (
        out.particles.shape, expected["particles"].shape)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < 1e-8
    assert tuple(out.survival_probabilities.shape) == tuple(expected["survival_probabilities"].shape)
    mismatches = int((out.survival_probabilities.cpu() != expected["survival_probabilities"]).sum())
    assert mismatches == 0
    assert torch.allclose(out.s.cpu(), expected["s"], rtol=1e-12)


@pytest.mark.parametrize("seed", range(20))
def test_random_vectorised_lattices_float32(seed):
    """The same random lattices in float32 against the float64 oracle on float32-rounded inputs:
    3e-6 x column maximum; survival may differ for a particle within float32 rounding of an
    aperture edge (at most 2 of 257 x settings here, none in practice)."""
    from oracle import lattice_io

    lattice, beam_shape = random_lattice(seed)
    g = torch.Generator().manual_seed(1000 + seed)
    n = 257
    sigma = torch.tensor([2e-4, 3e-5, 2e-4, 3e-5, 1e-4, 1e-3], dtype=torch.float64)
    particles = torch.randn((*beam_shape, n, 7), generator=g, dtype=torch.float64)
    particles[..., :6] *= sigma
    particles[..., 6] = 1.0
    energy = _vector(g, 5e7, 2e8, ())
    beam = oracle.make_beam(particles, energy)
    lattice32 = lattice_io.cast(lattice_io.cast(lattice, torch.float32), torch.float64)
    beam32 = {k: v.to(torch.float32).to(torch.float64) for k, v in beam.items()}
    expected = oracle.track(lattice32, beam32)
    out = gu.product_segment(lattice, DEVICE, torch.float32).track(
        gu.product_beam(beam, DEVICE, torch.float32))
    assert out.particles.dtype == torch.float32
    assert tuple(out.particles.shape) == tuple(expected["particles"].shape)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < 3e-6
    flips = int((out.survival_probabilities.cpu().double() != expected["survival_probabilities"]).sum())
    assert flips <= 2, flips


@pytest.mark.parametrize("seed", range(20))
def test_random_lattices_parameter_beam(seed):
    """ParameterBeam through random vectorised linear lattices (ch_apply_maps_parameter)."""
    import warnings

    import cheetah_b200 as cb

    lattice, _ = random_lattice(seed)
    g = torch.Generator().manual_seed(2000 + seed)
    sigma = torch.tensor([2e-4, 3e-5, 2e-4, 3e-5, 1e-4, 1e-3, 0.0], dtype=torch.float64)
    mixing = torch.eye(7, dtype=torch.float64) + 0.1 * torch.randn(7, 7, generator=g, dtype=torch.float64)
    mixing[6], mixing[:, 6] = 0.0, 0.0
    cov = mixing @ torch.diag(sigma.square()) @ mixing.T
    mu = torch.cat([torch.randn(6, generator=g, dtype=torch.float64) * sigma[:6],
                    torch.ones(1, dtype=torch.float64)])
    energy = _vector(g, 5e7, 2e8, [(), (3,)][seed % 2])
    mass = torch.tensor(oracle.ELECTRON_MASS_EV, dtype=torch.float64)
    exp_mu, exp_cov, _ = oracle.track_parameter_beam(lattice, mu, cov, energy, mass)
    beam = cb.ParameterBeam(mu.to(DEVICE), cov.to(DEVICE), energy.to(DEVICE),
                            species=cb.Species("electron", device=DEVICE, dtype=torch.float64))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # apertures are ignored for ParameterBeams
        out = gu.product_segment(lattice, DEVICE, torch.float64).track(beam)
    assert tuple(out.mu.shape) == tuple(exp_mu.shape), (out.mu.shape, exp_mu.shape)
    assert tuple(out.cov.shape) == tuple(exp_cov.shape)
    scale = exp_mu.abs().amax(dim=-1, keepdim=True).clamp_min(1e-300)
    assert float(((out.mu.cpu() - exp_mu).abs() / scale).max()) < 1e-9
    diag = exp_cov.diagonal(dim1=-2, dim2=-1).abs().sqrt()
    cscale = (diag.unsqueeze(-1) * diag.unsqueeze(-2))[..., :6, :6].clamp_min(1e-300)
    assert float(((out.cov.cpu() - exp_cov).abs()[..., :6, :6] / cscale).max()) < 1e-8
