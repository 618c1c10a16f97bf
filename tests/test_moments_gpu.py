"""Fused observables epilogue: moments from the apply kernel must equal the
moments of the fully materialised outgoing beam."""

import pytest
import torch

import workloads

pytestmark = pytest.mark.gpu
DEVICE = "cuda"


def reference_moments(out):
    """mu / sigma exactly as ParticleBeam.mu_* / sigma_* define them, in float64."""
    w = out.survival_probabilities.double()
    if w.dim() < out.particles.dim() - 1:
        w = w.expand(out.particles.shape[:-1])
    u = out.particles.double()[..., :6]
    s0 = w.sum(dim=-1)
    mu = (u * w.unsqueeze(-1)).sum(dim=-2) / s0.unsqueeze(-1)
    correction = s0 - w.square().sum(dim=-1) / s0
    var = (w.unsqueeze(-1) * (u - mu.unsqueeze(-2)).square()).sum(dim=-2) / correction.unsqueeze(-1)
    return mu, var.sqrt(), s0


@pytest.mark.parametrize("n,settings", [(100_000, 6), (4099, 3), (1_000_000, 4)])
def test_track_moments_matches_moments_of_tracked_beam(n, settings):
    description = workloads.ares_config3(settings, torch.float32)
    segment = workloads.product_segment(description, DEVICE, torch.float32)
    particles = workloads.twiss_beam_particles(n)
    particles[:, :4] *= 40.0  # fat beam: the apertures cut through the core
    beam = workloads.product_beam(particles, DEVICE, torch.float32)
    beam.survival_probabilities = (torch.rand(n, device=DEVICE) > 0.2).float()

    out = segment.track(beam)
    mu, sigma, s0 = reference_moments(out)
    observed = segment.track_moments(beam)
    assert observed.mu.shape == (settings, 6) and observed.sigma.shape == (settings, 6)
    assert torch.equal(observed.num_particles_survived.double(), s0)
    scale = sigma.abs().clamp_min(1e-30)
    assert ((observed.mu.double() - mu).abs() / scale).max() < 2e-5
    assert ((observed.sigma.double() - sigma).abs() / scale).max() < 2e-5
    assert torch.allclose(observed.sigma_x.double(), sigma[..., 0], rtol=2e-5)
    assert torch.allclose(observed.s, out.s)

    kept, observed2 = segment.track_moments(beam, keep_particles=True)
    assert torch.equal(kept.particles, out.particles)
    assert torch.equal(kept.survival_probabilities, out.survival_probabilities)
    # particles + moments and moments only are different kernels (scalar / packed pairs): same
    # sums in a different order
    assert ((observed2.mu.double() - observed.mu.double()).abs() / scale).max() < 2e-6
    assert ((observed2.sigma.double() - observed.sigma.double()).abs() / scale).max() < 2e-6
    assert torch.equal(observed2.num_particles_survived, observed.num_particles_survived)


def test_track_moments_without_apertures_and_single_setting():
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    segment = cb.Segment([cb.Quadrupole(length=t(0.2), k1=t(3.0), tilt=t(0.3)),
                          cb.Drift(length=t(1.0)), cb.HorizontalCorrector(length=t(0.1), angle=t(1e-3))])
    beam = cb.ParticleBeam.from_parameters(num_particles=50_001, device=DEVICE, dtype=torch.float32)
    out = segment.track(beam)
    mu, sigma, s0 = reference_moments(out)
    observed = segment.track_moments(beam)
    assert observed.mu.shape == (6,)
    assert float(observed.num_particles_survived) == 50_001
    assert ((observed.mu.double() - mu).abs() / sigma).max() < 2e-5
    assert ((observed.sigma.double() - sigma).abs() / sigma).max() < 2e-5


def reference_covariance(out):
    """unbiased_weighted_covariance_matrix (cheetah/utils/statistics.py:65-88) in float64."""
    w = out.survival_probabilities.double()
    if w.dim() < out.particles.dim() - 1:
        w = w.expand(out.particles.shape[:-1])
    u = out.particles.double()[..., :6]
    normalized = w / w.sum(dim=-1, keepdim=True)
    correction = 1 - normalized.square().sum(dim=-1)
    centred = u - (u * normalized.unsqueeze(-1)).sum(dim=-2, keepdim=True)
    return (normalized.unsqueeze(-1) * centred).mT @ centred / correction[..., None, None]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("keep", [False, True])
def test_track_moments_with_full_covariance(dtype, keep):
    """ch_apply_maps_covariance: the 6x6 covariance of the outgoing beam from the kernel epilogue
    (a tilted quadrupole couples x and y, the dipole adds dispersion)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=dtype)  # noqa: E731
    segment = cb.Segment([
        cb.Quadrupole(length=t(0.2), k1=t([3.0, -2.0, 0.5]), tilt=t(0.3)),
        cb.Drift(length=t(1.0)),
        cb.Dipole(length=t(0.5), angle=t(0.2)),
        cb.Aperture(x_max=t(4e-4), y_max=t(6e-4), shape="elliptical"),
        cb.Drift(length=t(0.7)),
    ])
    torch.manual_seed(2)
    beam = cb.ParticleBeam.from_parameters(num_particles=200_003, device=DEVICE, dtype=dtype)
    out = segment.track(beam)
    expected = reference_covariance(out)
    result = segment.track_moments(beam, keep_particles=keep, covariance=True)
    observed = result[1] if keep else result
    assert observed.cov.shape == (3, 6, 6) and observed.cov.dtype == dtype
    sigma = expected.diagonal(dim1=-2, dim2=-1).sqrt()
    scale = sigma.unsqueeze(-1) * sigma.unsqueeze(-2)
    tol = 3e-5 if dtype == torch.float32 else 1e-6  # fp32 per-tile partial sums in both cases
    assert float(((observed.cov.double().cpu() - expected.cpu()).abs() / scale.cpu()).max()) < tol
    assert torch.allclose(observed.cov, observed.cov.mT)
    assert torch.allclose(observed.cov.diagonal(dim1=-2, dim2=-1).sqrt(), observed.sigma, rtol=1e-5)
    plain = segment.track_moments(beam)
    assert plain.cov is None and torch.allclose(plain.sigma, observed.sigma, rtol=1e-6)
    if keep:
        assert torch.equal(result[0].particles, out.particles)


def test_track_moments_against_the_oracle_beam():
    """Directly against the reference's definitions: the float64 CPU oracle tracks the beam and
    mu / sigma / survivor count are taken from ITS outgoing beam (ParticleBeam.mu_* / sigma_*,
    particle_beam.py:1699-1805), not from this package's own tracked particles."""
    from oracle import lattice_io
    from oracle import track_oracle as oracle

    n, settings = 50_000, 5
    description = workloads.ares_config3(settings, torch.float32)
    particles = workloads.twiss_beam_particles(n)
    particles[:, :4] *= 40.0
    survival = (torch.rand(n, generator=torch.Generator().manual_seed(5)) > 0.2).float()

    truth_beam = workloads.oracle_beam(particles.float(), torch.float64)
    truth_beam["survival_probabilities"] = survival.double()
    truth = oracle.track(
        lattice_io.cast(lattice_io.cast(description, torch.float32), torch.float64), truth_beam)

    class Out:  # what reference_moments reads
        particles = truth["particles"]
        survival_probabilities = truth["survival_probabilities"]

    mu, sigma, s0 = reference_moments(Out)
    segment = workloads.product_segment(description, DEVICE, torch.float32)
    beam = workloads.product_beam(particles, DEVICE, torch.float32)
    beam.survival_probabilities = survival.to(DEVICE)
    observed = segment.track_moments(beam)
    assert torch.equal(observed.num_particles_survived.cpu().double(), s0)
    assert 0.05 < float(s0.mean()) / n < 0.95
    scale = sigma.abs().clamp_min(1e-30)
    assert ((observed.mu.cpu().double() - mu).abs() / scale).max() < 2e-5
    assert ((observed.sigma.cpu().double() - sigma).abs() / scale).max() < 2e-5


@pytest.mark.parametrize("covariance", [False, True], ids=["moments", "cov"])
@pytest.mark.parametrize("shape", ["rectangular", "elliptical"])
@pytest.mark.parametrize("n_apertures", [1, 2, 3, 4])
@pytest.mark.parametrize("beam_per_setting", [False, True], ids=["shared_beam", "beam_per_setting"])
@pytest.mark.parametrize("coupling", ["sparse", "coupled", "dense"])
def test_observables_kernel_variants(n_apertures, shape, covariance, beam_per_setting, coupling):
    """Every dispatch of the observables-only path (apply.cu): the kernel specialised for one
    beam under many settings (0-3 apertures, rectangular-only or with elliptical ones, with and
    without the covariance sums) and the general kernel (four apertures, a beam per setting)
    against the moments of the materialised outgoing beam."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=torch.float32)  # noqa: E731
    tilt = 0.0 if coupling == "sparse" else 0.3  # a tilted quadrupole couples x and y
    elements = [cb.Quadrupole(length=t(0.2), k1=t([4.0, -3.0, 1.5, 0.0, 2.5]), tilt=t(tilt))]
    if coupling == "dense":  # a tau column and a changed delta row: no sparsity flag holds
        matrix = torch.eye(7)
        matrix[0, 4], matrix[1, 4], matrix[2, 4] = 2e-3, -1e-3, 5e-4
        matrix[5, 4], matrix[5, 0], matrix[4, 5] = 3e-3, 1e-3, 0.2
        elements.append(cb.CustomTransferMap(predefined_transfer_map=t(matrix.tolist()),
                                             length=t(0.1)))
    for i in range(n_apertures):
        elements += [cb.Drift(length=t(0.4 + 0.1 * i)),
                     cb.Aperture(x_max=t(2.5e-4 + 5e-5 * i), y_max=t(3e-4 - 2e-5 * i), shape=shape)]
    elements += [cb.Quadrupole(length=t(0.2), k1=t(-2.0)), cb.Drift(length=t(0.5))]
    segment = cb.Segment(elements)
    torch.manual_seed(11)
    n = 70_001
    beam = cb.ParticleBeam.from_parameters(num_particles=n, device=DEVICE, dtype=torch.float32)
    if beam_per_setting:
        beam.particles = beam.particles.unsqueeze(0) * t([1.0, 1.1, 0.9, 1.2, 0.8]).view(5, 1, 1)
        beam.particles[..., 6] = 1.0
    out = segment.track(beam)
    assert 0.05 < float(out.survival_probabilities.mean()) < 0.95
    mu, sigma, s0 = reference_moments(out)
    observed = segment.track_moments(beam, covariance=covariance)
    assert torch.equal(observed.num_particles_survived.double(), s0)
    assert ((observed.mu.double() - mu).abs() / sigma).max() < 2e-5
    assert ((observed.sigma.double() - sigma).abs() / sigma).max() < 2e-5
    if covariance:
        expected = reference_covariance(out)
        scale = sigma.unsqueeze(-1) * sigma.unsqueeze(-2)
        assert float(((observed.cov.double() - expected).abs() / scale).max()) < 3e-5
