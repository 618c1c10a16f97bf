"""INTEGRATION.md section 2 executed: the dispatcher of ``cheetah_b200.integration`` patched into
the UNMODIFIED reference (``oracle/_ref``, installed by ``oracle/build_ref.py``), fed with the
reference's own elements and beams on the GPU.

Checks: the accelerated path returns a ``cheetah.ParticleBeam`` equal to the
reference's CPU result (2e-6 of the column scale in float32, survival masks exact); CPU beams,
``requires_grad`` inputs and lattices with elements outside the hot path fall through to the
reference implementation (cheetah/accelerator/element.py:159-193, cheetah/utils/cache.py:16-21).
"""

import pytest
import torch

from oracle import reference

from . import golden_utils as gu

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(not reference.available(), reason="oracle/_ref (the reference) is absent"),
]
DEVICE = "cuda"


@pytest.fixture()
def cheetah():
    from cheetah_b200 import integration

    module = reference.load()
    integration.install(module)
    integration.counters.update(accelerated=0, fallback=0)
    yield module
    integration.uninstall(module)


def ares(cheetah, device, dtype=torch.float32):
    from oracle import lattice_io

    description = gu.ares_lattice(dtype)
    for name, key, value in (
        ("AREAMQZM1", "k1", [8.2, -3.0]), ("AREAMQZM2", "k1", [-14.3, 5.0]),
        ("AREAMCVM1", "angle", 9e-5), ("AREAMQZM3", "k1", 3.142), ("AREAMCHM1", "angle", -1e-4),
    ):
        gu.set_attr(description, name, key, torch.tensor(value, dtype=dtype))
    for name in ("ARLISLHG1", "ARBCSLHB1"):
        gu.set_attr(description, name, "x_max", torch.tensor(3e-3, dtype=dtype))
        gu.set_attr(description, name, "y_max", torch.tensor(3e-3, dtype=dtype))
    segment = cheetah.Segment(elements=lattice_io.build(description, cheetah, None, dtype))
    return segment.to(device=device, dtype=dtype)


def beam_pair(cheetah, n=20_000, dtype=torch.float32):
    torch.manual_seed(11)
    cpu = cheetah.ParticleBeam.from_twiss(
        num_particles=n, beta_x=torch.tensor(3.14), beta_y=torch.tensor(42.0),
        emittance_x=torch.tensor(2e-8), emittance_y=torch.tensor(2e-8), energy=torch.tensor(1e8),
        dtype=dtype,
    )
    gpu = cheetah.ParticleBeam(
        particles=cpu.particles.to(DEVICE), energy=cpu.energy.to(DEVICE),
        particle_charges=cpu.particle_charges.to(DEVICE),
        survival_probabilities=cpu.survival_probabilities.to(DEVICE), s=cpu.s.to(DEVICE),
        species=cheetah.Species("electron", device=DEVICE, dtype=dtype), device=DEVICE,
        dtype=dtype,
    )
    return cpu, gpu


def test_reference_segment_and_beam_run_on_the_library(cheetah):
    from cheetah_b200 import _capi, integration

    cpu_beam, gpu_beam = beam_pair(cheetah)
    segment_gpu = ares(cheetah, DEVICE)
    before = _capi.launch_count()
    outgoing = segment_gpu.track(gpu_beam)
    torch.cuda.synchronize()
    assert integration.counters == {"accelerated": 1, "fallback": 0}
    assert _capi.launch_count() - before >= 2
    assert isinstance(outgoing, cheetah.ParticleBeam)
    assert outgoing.particles.is_cuda and outgoing.particles.dtype == torch.float32
    assert outgoing.particles.shape == (2, 20_000, 7)

    # the reference itself, on the CPU, in float64: ground truth for the same inputs
    segment_cpu = ares(cheetah, "cpu", torch.float64)
    truth = segment_cpu.track(
        cheetah.ParticleBeam(
            particles=cpu_beam.particles.double(), energy=cpu_beam.energy.double(),
            particle_charges=cpu_beam.particle_charges.double(),
            survival_probabilities=cpu_beam.survival_probabilities.double(),
            s=cpu_beam.s.double(), dtype=torch.float64,
        )
    )
    assert integration.counters["fallback"] == 1  # CPU beam: the reference's own path
    assert gu.column_scaled_error(outgoing.particles, truth.particles) < 2e-6
    assert torch.equal(outgoing.survival_probabilities.cpu().double(),
                       truth.survival_probabilities)
    assert 0.02 < float(truth.survival_probabilities.mean()) < 0.98
    assert torch.allclose(outgoing.s.cpu().double(), truth.s, rtol=1e-6)
    assert outgoing.species is not gpu_beam.species


def test_changed_settings_are_seen(cheetah):
    """Rebinding and in-place edits of the reference's buffers invalidate the cached lowering."""
    _, gpu_beam = beam_pair(cheetah, n=4000)
    segment = ares(cheetah, DEVICE)
    first = segment.track(gpu_beam).particles.clone()
    segment.AREAMQZM1.k1 = torch.tensor([1.0, 2.0], device=DEVICE)  # rebound
    second = segment.track(gpu_beam).particles.clone()
    assert not torch.equal(first, second)
    segment.AREAMQZM1.k1.mul_(0.5)  # in place
    third = segment.track(gpu_beam).particles.clone()
    assert not torch.equal(second, third)
    segment.AREAMQZM1.k1 = torch.tensor([8.2, -3.0], device=DEVICE)
    assert torch.equal(segment.track(gpu_beam).particles, first)


def test_fallbacks_use_the_reference_path(cheetah):
    from cheetah_b200 import _capi, integration

    cpu_beam, gpu_beam = beam_pair(cheetah, n=3000)
    segment = ares(cheetah, DEVICE)

    # requires_grad on a lattice setting: autograd must keep working, so the reference tracks
    k1 = torch.tensor(4.0, device=DEVICE, requires_grad=True)
    segment.AREAMQZM3.k1 = k1
    before = _capi.launch_count()
    out = segment.track(gpu_beam)
    assert _capi.launch_count() == before and integration.counters["fallback"] == 1
    out.particles[..., 0].square().mean().backward()
    assert k1.grad is not None and torch.isfinite(k1.grad)

    # ... unless gradients are switched off
    with torch.no_grad():
        segment.track(gpu_beam)
    assert integration.counters["accelerated"] == 1

    # requires_grad on the beam
    segment.AREAMQZM3.k1 = torch.tensor(4.0, device=DEVICE)
    grad_beam = cheetah.ParticleBeam(
        particles=gpu_beam.particles.clone().requires_grad_(True), energy=gpu_beam.energy,
        particle_charges=gpu_beam.particle_charges, device=DEVICE, dtype=torch.float32,
    )
    segment.track(grad_beam)
    assert integration.counters["fallback"] == 2

    # an element outside the hot path (a sextupole tracked with its non-default kick method would
    # be one; here: an element type the lowering does not know) -> the reference tracks it all
    class Strange(cheetah.Drift):
        pass

    odd = cheetah.Segment([cheetah.Drift(length=torch.tensor(0.3)),
                           Strange(length=torch.tensor(0.2))]).to(DEVICE)
    before = _capi.launch_count()
    out = odd.track(gpu_beam)
    assert integration.counters["fallback"] == 3 and isinstance(out, cheetah.ParticleBeam)

    # CPU beam and CPU lattice
    ares(cheetah, "cpu").track(cpu_beam)
    assert integration.counters["fallback"] == 4


def test_space_charge_kick_of_the_reference_runs_on_the_library(cheetah):
    from cheetah_b200 import integration

    torch.manual_seed(2)
    n = 50_000
    cpu = cheetah.ParticleBeam.from_parameters(
        num_particles=n, total_charge=torch.tensor(1e-9, dtype=torch.float64),
        energy=torch.tensor(5e7, dtype=torch.float64), dtype=torch.float64,
    )
    elements = lambda: [  # noqa: E731
        cheetah.Drift(length=torch.tensor(0.5)),
        cheetah.SpaceChargeKick(effect_length=torch.tensor(1.0), grid_shape=(32, 32, 32)),
        cheetah.Drift(length=torch.tensor(0.5)),
    ]
    truth = cheetah.Segment(elements()).to(torch.float64).track(cpu)
    segment = cheetah.Segment(elements()).to(DEVICE)
    gpu = cheetah.ParticleBeam(
        particles=cpu.particles.float().to(DEVICE), energy=cpu.energy.float().to(DEVICE),
        particle_charges=cpu.particle_charges.float().to(DEVICE), device=DEVICE,
        dtype=torch.float32,
    )
    out = segment.track(gpu)
    assert integration.counters["accelerated"] == 1
    assert isinstance(out, cheetah.ParticleBeam)
    start = cpu.particles
    moved = (truth.particles - start).abs().amax(dim=-2, keepdim=True)
    err = (out.particles.cpu().double() - gpu.particles.cpu().double() + start
           - truth.particles).abs() / moved.clamp_min(1e-300)
    assert float(err[..., (1, 3, 5)].max()) < 3e-3
