"""Host-buffer entry points (cheetah_b200/host.py): CPU tensors in, pinned host tensors out.

The compact output of ``ch_apply_maps_compact`` (6 coordinates per row, survival as a byte mask)
must carry exactly what ``Segment.track`` leaves on the device (element.py:181-191,
aperture.py:108-132), for unit and non-unit incoming survival probabilities, tail tiles and
rebinding of lattice tensors on the host."""

import pytest
import torch

import workloads

pytestmark = pytest.mark.gpu
DEVICE = "cuda:0"


def _host_case(n_particles, n_settings, dtype=torch.float32):
    import cheetah_b200 as cb
    from cheetah_b200 import lattice_description

    description = workloads.ares_config3(n_settings, dtype)
    host_segment = cb.Segment(elements=lattice_description.build(description, dtype=dtype))
    device_segment = workloads.product_segment(description, DEVICE, dtype)
    particles = workloads.twiss_beam_particles(n_particles)
    host_beam = cb.ParticleBeam(
        particles=particles.to(dtype), energy=torch.tensor(1e8, dtype=dtype),
        species=cb.Species("electron", dtype=dtype),
    )
    return host_segment, device_segment, host_beam, particles


class Collect:
    def __init__(self, n_settings, n, survival_dtype):
        self.coordinates = torch.zeros((n_settings, n, 6))
        self.survival = torch.zeros((n_settings, n), dtype=survival_dtype)
        self.calls = 0

    def __call__(self, begin, end, coordinates, survival):
        self.coordinates[begin:end] = coordinates
        self.survival[begin:end] = survival
        self.calls += 1


@pytest.mark.parametrize("n_particles,n_settings,chunk", [(5000, 37, 8), (1024, 3, 64), (999, 17, 4)])
def test_host_tracker_compact_output_equals_device_track(n_particles, n_settings, chunk):
    from cheetah_b200.host import HostTracker

    host_segment, device_segment, host_beam, particles = _host_case(n_particles, n_settings)
    tracker = HostTracker(host_segment, n_particles, n_settings, device=DEVICE, chunk_settings=chunk)
    got = Collect(n_settings, n_particles, torch.uint8)
    tracker.track(host_beam, consumer=got)
    expected = device_segment.track(workloads.product_beam(particles, DEVICE, torch.float32))
    assert got.calls == -(-n_settings // min(chunk, n_settings))
    assert torch.equal(got.coordinates, expected.particles[..., :6].cpu())
    assert torch.equal(got.survival.float(), expected.survival_probabilities.cpu())
    assert tracker.d2h_bytes == n_settings * n_particles * 25
    assert 0 < got.survival.float().mean() < 1


def test_host_tracker_non_unit_survival_travels_as_floats():
    from cheetah_b200.host import HostTracker
    import cheetah_b200 as cb

    n, b = 3000, 5
    host_segment, device_segment, host_beam, particles = _host_case(n, b)
    survival = torch.rand(n, generator=torch.Generator().manual_seed(3))
    host_beam = cb.ParticleBeam(
        particles=host_beam.particles, energy=host_beam.energy, survival_probabilities=survival,
        species=host_beam.species,
    )
    tracker = HostTracker(host_segment, n, b, device=DEVICE, chunk_settings=2)
    got = Collect(b, n, torch.float32)
    tracker.track(host_beam, consumer=got)
    device_beam = workloads.product_beam(particles, DEVICE, torch.float32)
    device_beam = cb.ParticleBeam(
        device_beam.particles, device_beam.energy, survival_probabilities=survival.to(DEVICE),
        species=device_beam.species,
    )
    expected = device_segment.track(device_beam)
    assert torch.equal(got.coordinates, expected.particles[..., :6].cpu())
    assert torch.equal(got.survival, expected.survival_probabilities.cpu())
    assert tracker.d2h_bytes == b * n * 28


def test_host_tracker_sees_rebound_and_in_place_settings():
    """`quad.k1 = tensor` on the host rebinds the buffer; `quad.k1.mul_()` changes it in place:
    both must reach the device at the next call (uploads resolved by name)."""
    from cheetah_b200.host import HostTracker

    n, b = 2000, 4
    host_segment, device_segment, host_beam, particles = _host_case(n, b)
    tracker = HostTracker(host_segment, n, b, device=DEVICE, chunk_settings=4)
    first = tracker.track_moments(host_beam).mu.clone()
    host_segment.AREAMQZM1.k1 = host_segment.AREAMQZM1.k1 * 0.5 + 1.0  # rebinding
    second = tracker.track_moments(host_beam).mu.clone()
    host_segment.AREAMQZM2.k1.mul_(-1.0)  # in place
    third = tracker.track_moments(host_beam).mu.clone()
    first, second, third = (t.nan_to_num(nan=-1.0) for t in (first, second, third))
    assert not torch.equal(first, second) and not torch.equal(second, third)
    device_segment.AREAMQZM1.k1 = host_segment.AREAMQZM1.k1.to(DEVICE)
    device_segment.AREAMQZM2.k1 = host_segment.AREAMQZM2.k1.to(DEVICE)
    expected = device_segment.track_moments(workloads.product_beam(particles, DEVICE, torch.float32))
    assert torch.equal(third, expected.mu.cpu().nan_to_num(nan=-1.0))


def test_host_tracker_keeps_its_plan_between_calls():
    """Uploading the lattice block must not look like an edit of the lattice: the lowered program
    of the device lattice is reused from call to call (value-watched tensors -- the cavity
    voltages of ARES -- live outside the flat block), yet a voltage changed on the host is seen."""
    from cheetah_b200 import tracking
    from cheetah_b200.host import HostTracker

    n, b = 2000, 4
    host_segment, device_segment, host_beam, particles = _host_case(n, b)
    tracker = HostTracker(host_segment, n, b, device=DEVICE, chunk_settings=4)
    assert tracker.separate, "ARES has cavities: their voltages shape the lowering"
    tracker.track_moments(host_beam)
    plan = tracking._plan(list(tracker.device_segment.elements), torch.device(DEVICE), (),
                          tracker.device_segment)
    before = tracker.track_moments(host_beam).mu.clone()
    assert tracking._plan(list(tracker.device_segment.elements), torch.device(DEVICE), (),
                          tracker.device_segment) is plan
    cavity = next(e for e in host_segment.elements if type(e).__name__ == "Cavity")
    cavity.voltage = torch.tensor(2e6)
    cavity.phase = torch.tensor(-20.0)
    getattr(device_segment, cavity.name).voltage = torch.tensor(2e6, device=DEVICE)
    getattr(device_segment, cavity.name).phase = torch.tensor(-20.0, device=DEVICE)
    after = tracker.track_moments(host_beam)
    expected = device_segment.track_moments(workloads.product_beam(particles, DEVICE, torch.float32))
    assert not torch.equal(before.nan_to_num(nan=-1.0), after.mu.nan_to_num(nan=-1.0))
    assert torch.equal(after.mu.nan_to_num(nan=-1.0), expected.mu.cpu().nan_to_num(nan=-1.0))


def test_track_host_space_charge_lattice_round_trip():
    """Any lattice through the general host path: FODO cell with two space-charge kicks."""
    from cheetah_b200.host import track_host
    import cheetah_b200 as cb

    dtype = torch.float32
    description = workloads.fodo_space_charge(1, 32, dtype)
    segment = workloads.product_segment(description, DEVICE, dtype)
    n = 20_000
    particles = workloads.parameters_beam_particles(n).to(dtype)
    charges = torch.full((n,), 1e-10 / n, dtype=dtype)
    out, survival, buffers = track_host(segment, particles, 1e8, charges, device=DEVICE)
    beam = cb.ParticleBeam(
        particles.to(DEVICE), torch.tensor(1e8, device=DEVICE), particle_charges=charges.to(DEVICE),
        species=cb.Species("electron", device=DEVICE, dtype=dtype),
    )
    expected = segment.track(beam)
    assert out.is_pinned() and out.shape == (n, 7)
    # the deposit's float atomics are order dependent: equal to ~1e-5 of the kick, not bit-equal
    kick = (expected.particles[:, 1] - beam.particles[:, 1]).abs().max().item()
    assert (out - expected.particles.cpu()).abs().max().item() < 1e-3 * max(kick, 1e-12) + 1e-12
    out2, _, buffers2 = track_host(segment, particles, 1e8, charges, device=DEVICE, buffers=buffers)
    assert buffers2 is buffers and out2.data_ptr() == out.data_ptr()
