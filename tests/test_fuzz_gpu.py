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
