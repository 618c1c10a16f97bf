"""The host-side beam mirror (cheetah_b200/beam.py) against values computed by the unmodified
reference (tests/golden/beam_properties.npz, oracle/make_golden.py --only-beam-properties):
means, sigmas, the 15 named covariances, emittances, Twiss and dispersion functions
(cheetah/particles/beam.py:262-557), the SI phase-space round trip, ``transformed_to``, the
conversions between the beam types and the Twiss constructor.

CPU beams use the closed formulas (set-up code); on a CUDA device ``second_moments`` is one pass
of the fused covariance kernel (``-m gpu`` test below).  float64, relative 1e-9 unless noted
(emittances subtract nearly equal products of second moments).
"""

import pytest
import torch

from . import golden_utils as gu

ARRAYS = gu.load_npz("beam_properties.npz")
NAMES = sorted(k.split(".", 1)[1] for k in ARRAYS if k.startswith("twiss.") and "." not in k[6:]
               and k not in ("twiss.mu", "twiss.cov"))


def particle_beam(device="cpu"):
    return gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), device, torch.float64)


def close(value, expected, rtol=1e-9):
    expected = gu.tensor(expected)
    value = value.detach().cpu().double()
    assert tuple(value.shape) == tuple(expected.shape), (value.shape, expected.shape)
    scale = expected.abs().max().clamp_min(1e-300)
    assert float((value - expected).abs().max() / scale) < rtol, (value, expected)


def check_scalar_properties(beam, prefix, rtol=1e-9, loose=1e-6):
    assert len(NAMES) == 45
    for name in NAMES:
        derived = "emittance" in name or "beta_" in name or "alpha_" in name
        close(getattr(beam, name), ARRAYS[f"{prefix}.{name}"], max(loose, rtol) if derived else rtol)


def test_particle_beam_properties_cpu():
    beam = particle_beam()
    check_scalar_properties(beam, "particle")
    close(beam.energies, ARRAYS["particle.energies"])
    close(beam.momenta, ARRAYS["particle.momenta"])
    close(beam.to_xyz_pxpypz(), ARRAYS["particle.xyz_pxpypz"], 1e-12)


def test_si_round_trip_and_transformed_to():
    import cheetah_b200 as cb

    beam = particle_beam()
    back = cb.ParticleBeam.from_xyz_pxpypz(beam.to_xyz_pxpypz(), beam.energy, species=beam.species)
    close(back.particles, ARRAYS["particle.round_trip"], 1e-12)
    moved = beam.transformed_to(mu_x=torch.tensor(1e-3, dtype=torch.float64),
                                sigma_py=torch.tensor(5e-5, dtype=torch.float64),
                                total_charge=torch.tensor(3e-11, dtype=torch.float64))
    close(moved.particles, ARRAYS["particle.transformed.particles"])
    close(moved.particle_charges, ARRAYS["particle.transformed.charges"])
    assert moved.survival_probabilities is beam.survival_probabilities
    with pytest.raises(AssertionError, match="cannot set"):
        beam.transformed_to(cov_xpx=1e-9)


def test_as_parameter_beam_and_back():
    beam = particle_beam()
    parameter = beam.as_parameter_beam()
    close(parameter.mu, ARRAYS["parameter.mu"])
    close(parameter.cov, ARRAYS["parameter.cov"])
    check_scalar_properties(parameter, "parameter")
    sampled = parameter.as_particle_beam(200_000, generator=torch.Generator().manual_seed(1))
    assert sampled.num_particles == 200_000
    assert torch.allclose(sampled.sigma_x, parameter.sigma_x, rtol=1e-2)
    assert torch.allclose(sampled.cov_xpx, parameter.cov_xpx, rtol=3e-2)
    clone = parameter.clone()
    assert torch.equal(clone.cov, parameter.cov) and clone.cov is not parameter.cov


def test_parameter_beam_from_twiss_and_transformed_to():
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    twiss = cb.ParameterBeam.from_twiss(
        beta_x=t([1.0, 2.5]), alpha_x=t(-0.7), emittance_x=t(3e-9), beta_y=t(4.0),
        alpha_y=t([0.2, 0.0]), emittance_y=t(2e-9), sigma_tau=t(1e-4), sigma_p=t(1e-3),
        cov_taup=t(2e-8), dispersion_x=t(0.03), dispersion_py=t(-0.02), energy=t(1.2e8),
        dtype=torch.float64,
    )
    close(twiss.mu, ARRAYS["twiss.mu"])
    close(twiss.cov, ARRAYS["twiss.cov"])
    check_scalar_properties(twiss, "twiss")
    changed = twiss.transformed_to(sigma_x=t(2e-4), mu_y=t(1e-4))
    close(changed.mu, ARRAYS["twiss.transformed.mu"])
    close(changed.cov, ARRAYS["twiss.transformed.cov"])
    with pytest.raises(ValueError, match="positive definite"):
        cb.ParameterBeam.from_parameters(sigma_x=t(1e-4), sigma_px=t(1e-5), cov_xpx=t(1e-8))
    with pytest.raises(AssertionError, match="Beta function in x"):
        cb.ParameterBeam.from_twiss(beta_x=t(0.0), beta_y=t(1.0))


def test_indexing_and_subsampling():
    import cheetah_b200 as cb

    beam = particle_beam()
    wide = cb.ParticleBeam(beam.particles.expand(3, 4000, 7) * 1.0, torch.tensor([1e8, 2e8, 3e8]).double(),
                           particle_charges=beam.particle_charges,
                           survival_probabilities=beam.survival_probabilities, species=beam.species)
    one = wide[1]
    assert tuple(one.particles.shape) == (4000, 7) and float(one.energy) == 2e8
    assert tuple(one.particle_charges.shape) == (4000,)
    assert tuple(wide[:2].particles.shape) == (2, 4000, 7)
    sub = beam.randomly_subsampled(500, random_state=torch.Generator().manual_seed(0))
    assert sub.num_particles == 500
    assert torch.allclose(sub.total_charge, beam.total_charge, rtol=1e-12)
    raw = beam.randomly_subsampled(500, adjust_particle_charges=False,
                                   random_state=torch.Generator().manual_seed(0))
    assert float(raw.total_charge.abs()) < float(beam.total_charge.abs())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_particle_beam_properties_from_the_covariance_kernel(dtype):
    """On a CUDA device every moment comes from one pass of the fused covariance kernel."""
    beam = gu.product_beam(gu.beam_dict(ARRAYS, "incoming"), "cuda", dtype)
    mean, covariance = beam.second_moments()
    assert mean.is_cuda and tuple(mean.shape) == (6,) and tuple(covariance.shape) == (6, 6)
    if dtype == torch.float64:
        # float64 beams accumulate in fp64 throughout (sums about the pilot particle)
        check_scalar_properties(beam, "particle", 1e-9, 1e-6)
        parameter = beam.as_parameter_beam()
        close(parameter.mu, ARRAYS["parameter.mu"], 1e-11)
        close(parameter.cov, ARRAYS["parameter.cov"], 1e-10)
    else:
        expected = gu.tensor(ARRAYS["parameter.cov"])[:6, :6]
        sigma = expected.diagonal().sqrt()
        error = (covariance.cpu().double() - expected).abs() / (sigma[:, None] * sigma[None, :])
        assert float(error.max()) < 2e-5
        close(beam.sigma_x, ARRAYS["particle.sigma_x"], 1e-5)
        close(beam.cov_xpx, ARRAYS["particle.cov_xpx"], 1e-4)
    # vectorised beam: one matrix per setting
    wide = type(beam)(beam.particles.expand(3, 4000, 7).contiguous(), beam.energy,
                      particle_charges=beam.particle_charges,
                      survival_probabilities=beam.survival_probabilities, species=beam.species)
    assert tuple(wide.second_moments()[1].shape) == (3, 6, 6)
    assert tuple(wide.emittance_x.shape) == (3,)


def test_waterbag_and_linspaced_generators():
    """uniform_3d_ellipsoid (particle_beam.py:563-666: uniform density inside the ellipsoid, so
    sigma = radius / sqrt(5)) and make_linspaced / linspaced (:668-803, :1180-1210)."""
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.uniform_3d_ellipsoid(
        num_particles=200_000, radius_x=1e-3, radius_y=2e-3, radius_tau=5e-4, sigma_px=1e-5,
        dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    inside = (beam.x / 1e-3) ** 2 + (beam.y / 2e-3) ** 2 + (beam.tau / 5e-4) ** 2
    assert float(inside.max()) <= 1.0 + 1e-12
    for name, radius in (("sigma_x", 1e-3), ("sigma_y", 2e-3), ("sigma_tau", 5e-4)):
        assert abs(float(getattr(beam, name)) / (radius / 5 ** 0.5) - 1.0) < 1e-2
    assert abs(float(beam.sigma_px) / 1e-5 - 1.0) < 1e-2
    t = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    line = cb.ParticleBeam.make_linspaced(num_particles=5, mu_x=t([0.0, 1e-3]), sigma_x=t(2e-4),
                                          sigma_p=t(1e-3), dtype=torch.float64)
    assert tuple(line.particles.shape) == (2, 5, 7)
    assert torch.allclose(line.x[1], t([8e-4, 9e-4, 1e-3, 1.1e-3, 1.2e-3]))
    assert torch.allclose(line.p[0], t([-1e-3, -5e-4, 0.0, 5e-4, 1e-3]))
    assert torch.equal(line.particles[..., 6], torch.ones(2, 5, dtype=torch.float64))
    again = particle_beam().linspaced(11)
    assert again.num_particles == 11
    assert torch.isclose(again.mu_x, particle_beam().mu_x, rtol=1e-9)
    assert torch.isclose(again.total_charge, particle_beam().total_charge, rtol=1e-9)


def test_generated_beams_have_exactly_the_requested_moments():
    """from_parameters / from_twiss / from_distribution whiten one shared sample and map it per
    vector entry (particle_beam.py:357-431, statistics.py:91-150): sample moments equal the
    targets to rounding, for vectorised parameters too."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    g = torch.Generator().manual_seed(0)
    beam = cb.ParticleBeam.from_parameters(
        num_particles=10_000, mu_y=3e-5, sigma_x=t([1e-4, 2e-4]), sigma_px=t(3e-5), cov_xpx=1e-10,
        cov_xp=t([[1e-8], [2e-8], [0.0]]), dtype=torch.float64, generator=g)
    assert tuple(beam.particles.shape) == (3, 2, 10_000, 7)
    assert torch.allclose(beam.sigma_x, t([1e-4, 2e-4]).expand(3, 2), rtol=1e-10)
    assert torch.allclose(beam.cov_xpx, t(1e-10).expand(3, 2), rtol=1e-9)
    assert torch.allclose(beam.cov_xp, t([[1e-8], [2e-8], [0.0]]).expand(3, 2), rtol=1e-9, atol=1e-20)
    assert torch.allclose(beam.mu_y, t(3e-5).expand(3, 2), rtol=1e-10)
    assert torch.allclose(beam.mu_x, torch.zeros(3, 2, dtype=torch.float64), atol=1e-18)
    twiss = cb.ParticleBeam.from_twiss(
        num_particles=5000, beta_x=3.14, alpha_x=-0.5, beta_y=t([42.0, 10.0]), dispersion_x=0.1,
        sigma_p=1e-3, dtype=torch.float64, generator=g)
    # (the dispersive part of sigma_x^2 is 4000 x the betatron part here: the dispersion-corrected
    # quantities subtract nearly equal numbers)
    assert torch.allclose(twiss.beta_x, t(3.14).expand(2), rtol=1e-6)
    assert torch.allclose(twiss.alpha_x, t(-0.5).expand(2), rtol=1e-6)
    assert torch.allclose(twiss.beta_y, t([42.0, 10.0]), rtol=1e-8)
    assert torch.allclose(twiss.dispersion_x, t(0.1).expand(2), rtol=1e-8)
    assert torch.allclose(twiss.emittance_x, t(7.1971891e-13).expand(2), rtol=1e-6)
    cold = cb.ParticleBeam.from_parameters(num_particles=1000, sigma_p=0.0, generator=g)
    assert float(cold.sigma_p) == 0.0 and float(cold.sigma_x) == pytest.approx(175e-6, rel=1e-5)
    with pytest.raises(AssertionError, match="Beta function in x"):
        cb.ParticleBeam.from_twiss(num_particles=10)
