"""GPU parity of active Screen / BPM elements (ch_screen_image, ch_screen_kde, ch_screen_gaussian,
ch_sc_beam_moments) through the
public API against the reference's outputs and the CPU oracle.

Tolerances: images are sums of float atomics, so the order differs from the reference's
``scatter_add_``: float64 1e-10 relative to the brightest pixel, float32 2e-5 (cloud-in-cell
weights are smooth); histogram images are exactly equal unless a particle sits within one ulp of a
bin edge (none does in the fixtures); total charge on the screen to 1e-6.
"""

import pytest
import torch

from oracle import diagnostics_oracle as diag

from . import golden_utils as gu
from .test_diagnostics_oracle import ARRAYS, GAUSSIAN_SCREENS, PARTICLE_SCREENS, SCREENS, TAGS

pytestmark = pytest.mark.gpu
DEVICE = "cuda"


def screen_kwargs(spec: dict, dtype) -> dict:
    return {k: (tuple(v) if k == "resolution" else
                torch.tensor(v, device=DEVICE, dtype=dtype) if isinstance(v, (list, float)) else v)
            for k, v in spec.items()}


@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("name", PARTICLE_SCREENS)
def test_screen_images_match_the_reference(name, tag, dtype):
    import cheetah_b200 as cb

    beam = gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), DEVICE, dtype)
    screen = cb.Screen(is_active=True, name=name, **screen_kwargs(SCREENS[name], dtype))
    out = cb.Segment([screen]).track(beam)
    assert torch.equal(out.particles, beam.particles)
    image = screen.reading
    expected = gu.tensor(ARRAYS[f"screen.{name}.{tag}"], dtype)
    assert image.dtype == dtype and tuple(image.shape) == tuple(expected.shape)
    scale = float(expected.max())
    got = image.cpu()
    if "histogram" in name:
        assert torch.allclose(got, expected, rtol=1e-6 if dtype == torch.float32 else 1e-12,
                              atol=scale * 1e-6)
    elif "kde" in name:
        # every particle is spread over the pixels within 7 / 8 bandwidths only: what is dropped
        # is < 2e-11 / 1e-14 of a particle's peak; float32 sums of ~1e3 terms per pixel
        tol = 1e-9 if dtype == torch.float64 else 2e-5
        assert float((got - expected).abs().max()) <= tol * scale
        assert abs(float(got.sum()) - 1.0) < 1e-5
    else:
        tol = 1e-10 if dtype == torch.float64 else 2e-5
        assert float((got - expected).abs().max()) <= tol * scale
    # (a float32 kde image is a quotient of float32 sums: its total is 1 to a few 1e-6)
    assert torch.isclose(got.sum(), expected.sum(),
                         rtol=2e-5 if "kde" in name and dtype == torch.float32 else 1e-6)
    assert screen.reading is image  # cached until the next beam


@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("name", GAUSSIAN_SCREENS)
def test_parameter_beam_screen_images_match_the_reference(name, tag, dtype):
    """ch_screen_gaussian: the analytic image of a ParameterBeam, including the reference's
    float32 arange grid (97 columns for a float32 screen of 96 pixels)."""
    import cheetah_b200 as cb

    mu = gu.tensor(ARRAYS[f"parameter_beam.mu.{tag}"], dtype).to(DEVICE)
    cov = gu.tensor(ARRAYS[f"parameter_beam.cov.{tag}"], dtype).to(DEVICE)
    beam = cb.ParameterBeam(mu, cov, torch.tensor(1e8, device=DEVICE, dtype=dtype),
                            species=cb.Species("electron", device=DEVICE, dtype=dtype))
    screen = cb.Screen(is_active=True, name=name, **screen_kwargs(SCREENS[name], dtype))
    out = screen.track(beam)
    assert torch.equal(out.mu, beam.mu)
    expected = gu.tensor(ARRAYS[f"screen.{name}.{tag}"], dtype)
    image = screen.reading
    assert image.dtype == dtype and tuple(image.shape) == tuple(expected.shape)
    # the grid points are float32 (torch.arange quirk) and ATen's CPU arange evaluates them in
    # SIMD chunks (base + lane * step in float32), so they are only defined to one float32 ulp:
    # 1e-10 m on a 3e-4 m beam = 1e-6 of the density
    tol = 5e-6 if dtype == torch.float64 else 2e-4
    assert float((image.cpu() - expected).abs().max()) <= tol * float(expected.max())
    vectorised = cb.ParameterBeam(mu.expand(3, 7).contiguous(), cov.expand(3, 7, 7).contiguous(),
                                  torch.tensor(1e8, device=DEVICE, dtype=dtype),
                                  species=beam.species)
    screen.track(vectorised)
    with pytest.raises(NotImplementedError, match="vectorization of `ParameterBeam`"):
        screen.reading


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_vectorised_kde_screen(tag, dtype):
    """method="kde" on a vectorised beam (tests/test_vectorized.py:305-352): one normalised image
    per setting."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=dtype)  # noqa: E731
    beam = gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), DEVICE, dtype)
    segment = cb.Segment([
        cb.HorizontalCorrector(length=t(0.1), angle=t([0.0, 1e-3, -2e-3])),
        cb.Drift(length=t(0.5)),
        cb.Screen(is_active=True, name="screen", resolution=(96, 64), pixel_size=t([2.5e-5, 3e-5]),
                  method="kde"),
    ])
    segment.track(beam)
    image = segment.screen.reading.cpu()
    expected = gu.tensor(ARRAYS[f"segment.kde_screen.{tag}"], dtype)
    assert image.shape == expected.shape == (3, 64, 96)
    scale = float(expected.max())
    assert float((image - expected).abs().max()) <= (1e-9 if dtype == torch.float64 else 5e-3) * scale
    assert torch.allclose(image.sum(dim=(-2, -1)), torch.ones(3, dtype=dtype), atol=1e-5)


def test_kde_bandwidths_against_the_oracle():
    """Bandwidths from half a pixel to 12 pixels (windows wider than one warp) and a beam that is
    partly off the screen, float64, against the dense oracle."""
    import cheetah_b200 as cb

    g = torch.Generator().manual_seed(8)
    n = 3000
    particles = torch.zeros(n, 7, dtype=torch.float64)
    particles[:, 0] = 4e-4 * torch.randn(n, generator=g, dtype=torch.float64) + 3e-4
    particles[:, 2] = 2e-4 * torch.randn(n, generator=g, dtype=torch.float64) - 1e-4
    particles[:, 6] = 1.0
    charges = torch.rand(n, generator=g, dtype=torch.float64) * 1e-15
    survival = (torch.rand(n, generator=g, dtype=torch.float64) > 0.1).double()
    oracle_beam = {"particles": particles, "particle_charges": charges,
                   "survival_probabilities": survival}
    beam = cb.ParticleBeam(particles.to(DEVICE), torch.tensor(1e8, device=DEVICE, dtype=torch.float64),
                           particle_charges=charges.to(DEVICE),
                           survival_probabilities=survival.to(DEVICE),
                           species=cb.Species("electron", device=DEVICE, dtype=torch.float64))
    for bandwidth in (1e-5, 2e-5, 9e-5, 2.4e-4):
        spec = {"resolution": (80, 50), "pixel_size": (2e-5, 2e-5), "method": "kde",
                "kde_bandwidth": bandwidth}
        screen = cb.Screen(is_active=True, resolution=(80, 50), method="kde",
                           pixel_size=torch.tensor([2e-5, 2e-5], device=DEVICE, dtype=torch.float64),
                           kde_bandwidth=torch.tensor(bandwidth, device=DEVICE, dtype=torch.float64))
        screen.track(beam)
        expected = diag.screen_reading(spec, oracle_beam)
        got = screen.reading.cpu()
        assert tuple(got.shape) == (50, 80)
        assert float((got - expected).abs().max()) <= 1e-9 * float(expected.max()), bandwidth


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_bpm_and_screen_inside_a_vectorised_segment(tag, dtype):
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=dtype)  # noqa: E731
    beam = gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), DEVICE, dtype)
    segment = cb.Segment([
        cb.HorizontalCorrector(length=t(0.1), angle=t([0.0, 1e-3, -2e-3])),
        cb.Drift(length=t(0.5)),
        cb.BPM(is_active=True, name="bpm", misalignment=t([1e-4, 2e-4])),
        cb.Screen(is_active=True, name="screen", resolution=(96, 64), pixel_size=t([2.5e-5, 3e-5])),
    ])
    out = segment.track(beam)
    assert tuple(out.particles.shape) == tuple(int(v) for v in ARRAYS[f"segment.outgoing_shape.{tag}"])
    expected = gu.tensor(ARRAYS[f"segment.bpm.{tag}"], dtype)
    reading = segment.bpm.reading.cpu()
    assert reading.shape == expected.shape == (3, 2)
    # fp64 accumulation about a pilot particle: closer to the exact mean than the reference's
    # float32 sum, hence the float32 tolerance of a few ulp of the 1e-4 m offsets
    assert torch.allclose(reading, expected, rtol=1e-11 if dtype == torch.float64 else 2e-5,
                          atol=1e-16 if dtype == torch.float64 else 2e-9)
    image = segment.screen.reading.cpu()
    expected = gu.tensor(ARRAYS[f"segment.screen.{tag}"], dtype)
    assert image.shape == expected.shape == (3, 64, 96)
    scale = float(expected.max())
    # float32: the screen sees positions that went through the (float32) linear section
    assert float((image - expected).abs().max()) <= (1e-9 if dtype == torch.float64 else 5e-3) * scale
    assert torch.allclose(image.sum(dim=(-2, -1)), expected.sum(dim=(-2, -1)), rtol=1e-5)


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_bpm_misalignment_and_blocking_screen(tag, dtype):
    import cheetah_b200 as cb

    beam = gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), DEVICE, dtype)
    bpm = cb.BPM(is_active=True, misalignment=torch.tensor([0.1, 0.2], device=DEVICE, dtype=dtype))
    assert torch.isnan(bpm.reading).all()
    bpm.track(beam)
    assert torch.allclose(bpm.reading.cpu(), gu.tensor(ARRAYS[f"bpm.{tag}"], dtype),
                          rtol=1e-12 if dtype == torch.float64 else 1e-6)
    blocking = cb.Screen(is_active=True, is_blocking=True,
                         pixel_size=torch.tensor([1e-3, 1e-3], device=DEVICE, dtype=dtype),
                         misalignment=torch.zeros(2, device=DEVICE, dtype=dtype))
    out = blocking.track(beam)
    assert torch.equal(out.survival_probabilities.cpu(),
                       gu.tensor(ARRAYS[f"blocking.survival.{tag}"], dtype))
    assert torch.equal(out.particles, beam.particles)
    inactive = cb.Screen(resolution=(10, 8),
                         pixel_size=torch.tensor([1e-3, 1e-3], device=DEVICE, dtype=dtype))
    assert tuple(inactive.reading.shape) == (8, 10) and float(inactive.reading.abs().sum()) == 0.0


def test_full_size_screen_conserves_charge_and_matches_the_oracle():
    """1e6 particles on the default 1024 x 1024 screen: every particle is inside, so the image
    sums to the beam charge; a 20 000-particle slice is compared with the oracle."""
    import cheetah_b200 as cb

    torch.manual_seed(4)
    beam = cb.ParticleBeam.from_parameters(
        num_particles=1_000_000, sigma_x=2e-3, sigma_y=1e-3, total_charge=torch.tensor(1e-9),
        device=DEVICE, dtype=torch.float32,
    )
    screen = cb.Screen(is_active=True, pixel_size=torch.tensor([2e-5, 2e-5], device=DEVICE))
    screen.track(beam)
    image = screen.reading
    assert tuple(image.shape) == (1024, 1024)
    assert torch.isclose(image.double().sum(), beam.particle_charges.double().abs().sum(), rtol=1e-5)
    part = slice(0, 20_000)
    small = cb.ParticleBeam(beam.particles[part], beam.energy,
                            particle_charges=beam.particle_charges[part], species=beam.species)
    screen.track(small)
    oracle_beam = {
        "particles": small.particles.cpu().double(),
        "particle_charges": small.particle_charges.cpu().double(),
        "survival_probabilities": small.survival_probabilities.cpu().double(),
    }
    expected = diag.screen_reading({"pixel_size": (2e-5, 2e-5)}, oracle_beam)
    got = screen.reading.cpu().double()
    # float32 bin-space positions (|px| <= 512) carry ~3e-5 of a cell
    assert float((got - expected).abs().max()) <= 2e-4 * float(expected.max())


def test_unsupported_screen_modes_are_loud():
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_parameters(num_particles=100, device=DEVICE)
    vectorised = cb.ParticleBeam(beam.particles.expand(2, 100, 7).contiguous(), beam.energy,
                                 species=beam.species)
    histogram = cb.Screen(is_active=True, method="histogram",
                          pixel_size=torch.tensor([1e-3, 1e-3], device=DEVICE))
    histogram.track(vectorised)
    with pytest.raises(NotImplementedError, match="does not support vectorization"):
        histogram.reading
    with pytest.raises(AssertionError, match="Invalid method"):
        cb.Screen(method="nonsense")
