"""The kernels specialised for one beam under many consecutive settings (csrc/apply_lean.cu) must
give exactly what the general kernels of csrc/apply.cu give: the same fma chains, so outgoing
coordinates and survival masks are compared bit for bit.  The general kernels are the ones pinned
to the oracle and the reference's pickles in test_linear_gpu.py."""

import pytest
import torch

pytestmark = pytest.mark.gpu
DEVICE = "cuda"


def _segment(n_apertures, shape, coupling, dtype=torch.float32):
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=dtype)  # noqa: E731
    tilt = 0.3 if coupling != "sparse" else 0.0
    elements = [cb.Quadrupole(length=t(0.2), k1=t([4.0, -3.0, 1.5, 0.0, 2.5, -1.0, 0.7]),
                              tilt=t(tilt)),
                cb.HorizontalCorrector(length=t(0.1), angle=t(2e-4))]
    for i in range(n_apertures):
        elements += [cb.Drift(length=t(0.4 + 0.1 * i)),
                     cb.Aperture(x_max=t(2.5e-4 + 5e-5 * i), y_max=t(3e-4 - 2e-5 * i), shape=shape)]
    elements += [cb.Quadrupole(length=t(0.2), k1=t(-2.0)), cb.Drift(length=t(0.5))]
    if coupling == "dense":
        matrix = torch.eye(7)
        matrix[0, 4], matrix[1, 4], matrix[2, 4] = 2e-3, -1e-3, 5e-4
        matrix[5, 4], matrix[5, 0], matrix[4, 5] = 3e-3, 1e-3, 0.2
        elements.append(cb.CustomTransferMap(predefined_transfer_map=t(matrix.tolist()),
                                             length=t(0.1)))
    return cb.Segment(elements)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("coupling", ["sparse", "coupled", "dense"])
@pytest.mark.parametrize("shape", ["rectangular", "elliptical"])
@pytest.mark.parametrize("n_apertures", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n", [70_001, 1024, 37])
def test_shared_beam_track_equals_general_kernel(n, n_apertures, shape, coupling, dtype):
    import cheetah_b200 as cb

    if n_apertures == 0 and shape == "elliptical":
        pytest.skip("no aperture to shape")
    segment = _segment(n_apertures, shape, coupling, dtype)
    torch.manual_seed(3)
    beam = cb.ParticleBeam.from_parameters(num_particles=n, device=DEVICE, dtype=dtype)
    beam.survival_probabilities = (torch.rand(n, device=DEVICE) > 0.1).to(dtype)
    out = segment.track(beam)  # one beam, seven settings: apply_shared_beam_kernel up to 3 apertures

    per_setting = cb.ParticleBeam(
        particles=beam.particles.unsqueeze(0).repeat(7, 1, 1), energy=beam.energy,
        survival_probabilities=beam.survival_probabilities.unsqueeze(0).repeat(7, 1),
        particle_charges=beam.particle_charges, device=DEVICE, dtype=dtype)
    general = segment.track(per_setting)  # a beam per setting: apply_maps_kernel
    assert out.particles.shape == (7, n, 7)
    assert torch.equal(out.particles, general.particles)
    # (without apertures the shared incoming survival vector is passed through un-broadcast)
    assert torch.equal(out.survival_probabilities.expand(7, n), general.survival_probabilities)
    if n_apertures and n > 1000:
        assert 0.02 < float(out.survival_probabilities.mean()) < 0.9


def test_shared_beam_kernel_is_the_one_launched():
    """The dispatch itself: counts of launches do not tell kernels apart, so the library's last
    kernel name is not available -- instead check a call shape the lean kernels refuse (unaligned
    beam view: no TMA tile) still matches."""
    import cheetah_b200 as cb

    segment = _segment(2, "rectangular", "coupled")
    torch.manual_seed(4)
    n = 4099
    base = cb.ParticleBeam.from_parameters(num_particles=n + 1, device=DEVICE, dtype=torch.float32)
    shifted = cb.ParticleBeam(particles=base.particles[1:], energy=base.energy,
                              particle_charges=base.particle_charges[1:], device=DEVICE,
                              dtype=torch.float32)  # rows start 28 bytes into the allocation
    aligned = cb.ParticleBeam(particles=base.particles[1:].clone(), energy=base.energy,
                              particle_charges=base.particle_charges[1:].clone(), device=DEVICE,
                              dtype=torch.float32)
    a, b = segment.track(shifted), segment.track(aligned)
    assert torch.equal(a.particles, b.particles)
    assert torch.equal(a.survival_probabilities, b.survival_probabilities)
