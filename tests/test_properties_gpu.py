"""Size-independent properties at BASELINE.json's full sizes (1e6 particles), where the CPU
oracle would take too long: round trips, composition, checksums, idempotence."""

import pytest
import torch

import workloads

pytestmark = pytest.mark.gpu
DEVICE = "cuda"
N = 1_000_000


@pytest.fixture(scope="module")
def beam():
    return workloads.product_beam(workloads.twiss_beam_particles(N), DEVICE, torch.float32)


def test_negative_length_drift_inverts(beam):
    """tests/test_drift.py:95-115 at 1e6 particles: Drift(L) then Drift(-L) is the identity."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    there = cb.Drift(length=t(1.7)).track(beam)
    back = cb.Drift(length=t(-1.7)).track(there)
    scale = beam.particles.abs().amax(dim=0)
    assert ((back.particles - beam.particles).abs().amax(dim=0) <= 2e-7 * scale + 1e-30).all()
    # as one merged segment the maps cancel exactly in the fp64 composer
    both = cb.Segment([cb.Drift(length=t(1.7)), cb.Drift(length=t(-1.7))]).track(beam)
    assert torch.equal(both.particles, beam.particles)


def test_unpowered_quadrupole_is_a_drift(beam):
    """tests/test_quadrupole.py:7-27, :265-292 (k1 = 0 quadrupole == drift), full size."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    a = cb.Quadrupole(length=t(0.8), k1=t(0.0)).track(beam)
    b = cb.Drift(length=t(0.8)).track(beam)
    assert torch.equal(a.particles, b.particles)


def test_segment_equals_composition_of_its_halves_config3_shape():
    """ARES with 16 vectorised settings x 1e6 particles: tracking the whole lattice in one fused
    pass equals tracking its two halves one after the other (survival masks identical)."""
    description = workloads.ares_config3(16, torch.float32)
    beam = workloads.product_beam(workloads.twiss_beam_particles(N), DEVICE, torch.float32)
    whole = workloads.product_segment(description, DEVICE, torch.float32).track(beam)
    cut = 100
    first = workloads.product_segment(description[:cut], DEVICE, torch.float32).track(beam)
    second = workloads.product_segment(description[cut:], DEVICE, torch.float32).track(first)
    assert whole.particles.shape == second.particles.shape == (16, N, 7)
    scale = whole.particles.abs().amax(dim=-2, keepdim=True)
    err = ((whole.particles - second.particles).abs() / scale.clamp_min(1e-30))[..., :6].max()
    # the split run rounds the intermediate beam to fp32 once more; the second half's maps
    # (|R| up to ~10 with these random optics) amplify that rounding
    assert err < 2e-5, float(err)
    mismatches = int((whole.survival_probabilities != second.survival_probabilities).sum())
    # edge-band flips are possible in principle; report, tolerate a handful
    assert mismatches <= 4, mismatches
    assert torch.allclose(whole.s, second.s)
    assert 0.2 < float(whole.survival_probabilities.mean()) < 0.8


def test_apertures_are_idempotent_and_monotone(beam):
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    aperture = cb.Aperture(x_max=t(1e-6), y_max=t(4e-6), shape="elliptical")
    once = aperture.track(beam)
    twice = aperture.track(once)
    assert torch.equal(once.survival_probabilities, twice.survival_probabilities)
    wider = cb.Aperture(x_max=t(2e-6), y_max=t(8e-6), shape="elliptical").track(beam)
    assert bool((wider.survival_probabilities >= once.survival_probabilities).all())
    frac = float(once.survival_probabilities.mean())
    assert 0.05 < frac < 0.95, frac


def test_space_charge_checksums_full_size():
    """64^3 grid, 1e6 particles (BASELINE configs[3] sizes): deposited charge equals the charge
    inside the grid, positions are untouched, and the kick matches the float64 CPU oracle."""
    import cheetah_b200 as cb
    from cheetah_b200 import space_charge

    particles = workloads.parameters_beam_particles(N).to(torch.float32)
    dev = lambda x: x.to(DEVICE)  # noqa: E731
    charges = torch.full((N,), 1e-10 / N)
    survival = (torch.rand(N, generator=torch.Generator().manual_seed(3)) > 0.1).float()
    args = dict(
        energy=dev(torch.tensor(1e8)), charges=dev(charges), survival=dev(survival),
        mass_eV=dev(torch.tensor(510998.95069)), effect_length=dev(torch.tensor(1.0)),
        extents=tuple(dev(torch.tensor(3.0)) for _ in range(3)), grid_shape=(64, 64, 64),
    )
    out, ws = space_charge.kick(dev(particles), **args)
    grid = ws.charge_grid().double()
    half = ws.params[0, 0:3]
    p = dev(particles).double()
    z = -ws.params[0, 7] * p[:, 4]
    inside = (p[:, 0].abs() <= half[0]) & (p[:, 2].abs() <= half[1]) & (z.abs() <= half[2])
    expected = (dev(charges).double() * dev(survival).double() * inside).sum()
    # CIC loses the part of edge clouds that falls off the grid: compare within the edge share
    assert abs(float(grid.sum() / expected) - 1.0) < 2e-3
    assert float(grid.min()) >= 0.0
    for col in (0, 2, 4, 6):
        assert torch.equal(out[0, :, col], dev(particles)[:, col])
    # the float64 CPU oracle is fast enough for ONE kick at the full size: compare the kicks
    from oracle import track_oracle as oracle

    beam64 = oracle.make_beam(particles.double(), torch.tensor(1e8, dtype=torch.float64),
                              particle_charges=charges.double(),
                              survival_probabilities=survival.double())
    el = {"type": "SpaceChargeKick", "effect_length": torch.tensor(1.0, dtype=torch.float64),
          "grid_shape": (64, 64, 64)}
    expected = oracle.track_space_charge(el, beam64)["particles"]
    kick_ref = (expected - particles.double())[:, [1, 3, 5]]
    kick_ours = (out[0].cpu().double() - particles.double())[:, [1, 3, 5]]
    err = (kick_ours - kick_ref).abs().amax(dim=0) / kick_ref.abs().amax(dim=0)
    assert (err < 3e-3).all(), err


def test_space_charge_odd_particle_count_and_survival_weighting():
    """N not a multiple of 4 (non-TMA code path) and lost particles carry no charge:
    tests/test_space_charge_kick.py:369-409."""
    from oracle import track_oracle as oracle
    from . import golden_utils as gu

    n = 5003
    g = torch.Generator().manual_seed(9)
    particles = workloads.parameters_beam_particles(n, seed=9)
    survival = (torch.rand(n, generator=g) > 0.3).double()
    beam = oracle.make_beam(particles, torch.tensor(3e7, dtype=torch.float64),
                            particle_charges=torch.full((n,), 2e-10 / n, dtype=torch.float64),
                            survival_probabilities=survival)
    el = {"type": "SpaceChargeKick", "name": "sc", "effect_length": torch.tensor(0.4, dtype=torch.float64),
          "grid_shape": (16, 32, 8)}
    expected = oracle.track([el], beam)
    for dtype, tol in ((torch.float64, 1e-7), (torch.float32, 3e-3)):
        beam_t = {k: v.to(dtype) for k, v in beam.items()}
        from oracle import lattice_io

        out = gu.product_segment(lattice_io.cast([el], dtype), DEVICE, dtype).track(
            gu.product_beam(beam_t, DEVICE, dtype))
        moved = (expected["particles"] - particles).abs().amax(dim=0)
        ours = out.particles.cpu().double() - beam_t["particles"].double() + particles
        err = ((ours - expected["particles"]).abs().amax(dim=0) / moved.clamp_min(1e-300))[[1, 3, 5]]
        assert (err < tol).all(), err


# ---- non-linear tracking methods at full size ---------------------------
def _column_error(a, b, reference):
    scale = reference.abs().amax(dim=-2, keepdim=True).clamp_min(1e-30)
    return float(((a - b).abs() / scale)[..., :6].max())


def test_drift_kick_drift_round_trip_and_limits(beam):
    """1e6 particles: the exact drift inverts under L -> -L, a k1 = 0 quadrupole and a zero-angle
    dipole are exact drifts, an unpowered TDC is two half drifts, and a run of elements equals
    tracking them one by one (tests/test_drift.py:43-69, test_dipole.py:153-175)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    dkd = "drift_kick_drift"
    there = cb.Drift(length=t(1.7), tracking_method=dkd).track(beam)
    back = cb.Drift(length=t(-1.7), tracking_method=dkd).track(there)
    assert _column_error(back.particles, beam.particles, beam.particles) < 5e-7
    assert torch.equal(back.particles[..., 5], beam.particles[..., 5])  # delta untouched
    drift = cb.Drift(length=t(0.8), tracking_method=dkd).track(beam)
    quad = cb.Quadrupole(length=t(0.8), k1=t(0.0), num_steps=4, tracking_method=dkd).track(beam)
    dipole = cb.Dipole(length=t(0.8), angle=t(0.0), tracking_method=dkd).track(beam)
    tdc = cb.TransverseDeflectingCavity(length=t(0.8), voltage=t(0.0)).track(beam)
    for other in (quad, dipole, tdc):
        assert _column_error(other.particles, drift.particles, drift.particles) < 1e-6
    # linear drift agrees to second order in the (small) angles
    linear = cb.Drift(length=t(0.8)).track(beam)
    assert _column_error(drift.particles[..., :4], linear.particles[..., :4],
                         linear.particles[..., :4]) < 1e-4
    elements = [
        cb.Drift(length=t(0.3), tracking_method=dkd),
        cb.Quadrupole(length=t(0.2), k1=t(4.0), num_steps=3, tracking_method=dkd),
        cb.Marker(),
        cb.Dipole(length=t(0.4), angle=t(0.05), tracking_method=dkd),
        cb.Sextupole(length=t(0.1), k2=t(20.0)),
        cb.Drift(length=t(0.3), tracking_method="second_order"),
    ]
    fused = cb.Segment(elements).track(beam)
    step = beam
    for element in elements:
        step = element.track(step)
    assert _column_error(fused.particles, step.particles, step.particles) < 2e-6
    assert torch.allclose(fused.s, step.s)


def test_second_order_reduces_to_linear_for_small_amplitudes(beam):
    """Second-order maps = linear map + O(amplitude^2): shrinking the beam by 1e-3 leaves only
    the linear part (sextupole.py:84-116, quadrupole.py:93-143)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    small = workloads.product_beam(workloads.twiss_beam_particles(N), DEVICE, torch.float32)
    small.particles[..., :6] *= 1e-3
    for second, linear in (
        (cb.Quadrupole(length=t(0.3), k1=t(5.0), tilt=t(0.2), tracking_method="second_order"),
         cb.Quadrupole(length=t(0.3), k1=t(5.0), tilt=t(0.2))),
        (cb.Sextupole(length=t(0.3), k2=t(50.0)), cb.Drift(length=t(0.3))),
        (cb.Dipole(length=t(0.5), angle=t(0.1), tracking_method="second_order"),
         cb.Dipole(length=t(0.5), angle=t(0.1))),
    ):
        a, b = second.track(small), linear.track(small)
        assert _column_error(a.particles, b.particles, b.particles) < 5e-5


def test_screen_conserves_the_surviving_charge_at_full_size():
    """4 settings x 1e6 particles behind an aperture: every image sums to the charge that
    survived and lies on the screen (screen.py:316-340)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    torch.manual_seed(8)
    beam = cb.ParticleBeam.from_parameters(num_particles=N, total_charge=torch.tensor(1e-9),
                                           sigma_x=3e-4, sigma_y=2e-4, device=DEVICE)
    segment = cb.Segment([
        cb.Quadrupole(length=t(0.2), k1=t([0.0, 4.0, -4.0, 8.0])),
        cb.Drift(length=t(1.0)),
        cb.Aperture(x_max=t(4e-4), y_max=t(4e-4)),
        cb.Screen(is_active=True, name="screen", resolution=(512, 512), pixel_size=t([4e-6, 4e-6])),
    ])
    out = segment.track(beam)
    image = segment.screen.reading
    assert tuple(image.shape) == (4, 512, 512)
    on_screen = (out.particles[..., 0].abs() < 1.0e-3) & (out.particles[..., 2].abs() < 1.0e-3)
    expected = (out.particle_charges.abs() * out.survival_probabilities * on_screen).double().sum(-1)
    assert torch.allclose(image.double().sum(dim=(-2, -1)), expected, rtol=2e-4)


@pytest.mark.parametrize("n", [1, 2, 3, 1023, 1025])
def test_tiny_and_ragged_beams(n):
    """One to a few particles and tile-boundary sizes through every kernel family against the
    oracle / closed formulas (ragged last tiles, single-particle statistics)."""
    import cheetah_b200 as cb
    from oracle import diagnostics_oracle as diag
    from oracle import track_oracle as oracle
    from . import golden_utils as gu

    g = torch.Generator().manual_seed(n)
    particles = torch.randn(n, 7, generator=g, dtype=torch.float64) * 1e-4
    particles[:, 6] = 1.0
    t = lambda v: torch.tensor(v, device="cuda", dtype=torch.float64)  # noqa: E731
    lattice = [
        {"type": "Quadrupole", "name": "q", "length": torch.tensor(0.2, dtype=torch.float64),
         "k1": torch.tensor([3.0, -2.0], dtype=torch.float64)},
        {"type": "Aperture", "name": "a", "x_max": torch.tensor(1.5e-4, dtype=torch.float64),
         "y_max": torch.tensor(2e-4, dtype=torch.float64), "shape": "rectangular", "is_active": True},
        {"type": "Drift", "name": "d", "length": torch.tensor(1.0, dtype=torch.float64)},
    ]
    beam = oracle.make_beam(particles, torch.tensor(1e8, dtype=torch.float64))
    expected = oracle.track(lattice, beam)
    segment = gu.product_segment(lattice, "cuda", torch.float64)
    product = gu.product_beam(beam, "cuda", torch.float64)
    out = segment.track(product)
    assert tuple(out.particles.shape) == (2, n, 7)
    assert torch.allclose(out.particles.cpu(), expected["particles"], rtol=1e-11, atol=1e-18)
    assert torch.equal(out.survival_probabilities.cpu(), expected["survival_probabilities"])
    observed = segment.track_moments(product, covariance=True)
    survivors = expected["survival_probabilities"].sum(dim=-1)
    assert torch.equal(observed.num_particles_survived.cpu(), survivors)
    w = expected["survival_probabilities"]
    mean = (expected["particles"][..., 0] * w).sum(-1) / w.sum(-1)
    ok = survivors > 0
    assert torch.allclose(observed.mu[..., 0].cpu()[ok], mean[ok], rtol=1e-10, atol=1e-18)
    # non-linear run and a screen on the same beam
    nonlinear = [{"type": "Drift", "name": "d", "length": torch.tensor(0.7, dtype=torch.float64),
                  "tracking_method": "drift_kick_drift"},
                 {"type": "Sextupole", "name": "s", "length": torch.tensor(0.1, dtype=torch.float64),
                  "k2": torch.tensor(20.0, dtype=torch.float64), "tracking_method": "second_order"}]
    got = gu.product_segment(nonlinear, "cuda", torch.float64).track(product)
    assert gu.column_scaled_error(got.particles, oracle.track(nonlinear, beam)["particles"]) < 1e-11
    for method in ("cloud-in-cell", "histogram", "kde"):
        screen = cb.Screen(is_active=True, resolution=(40, 30), method=method,
                           pixel_size=t([2e-5, 2e-5]))
        screen.track(product)
        spec = {"resolution": (40, 30), "pixel_size": (2e-5, 2e-5), "method": method}
        reference = diag.screen_reading(spec, beam)
        image = screen.reading.cpu()
        assert float((image - reference).abs().max()) <= 1e-9 * float(reference.max()) + 1e-300


def test_config3_workload_against_the_oracle_at_full_beam_size():
    """The bench workload itself (BASELINE configs[2]: ARES, 1e6 particles, settings 0..3 of the
    4096 drawn by workloads.ares_config3) on the GPU in float32 against the float64 CPU oracle:
    2e-6 of the column scale, survival masks exactly equal (what bench.py reports as
    `parity_check` for settings 0..7)."""
    from oracle import lattice_io
    from oracle import track_oracle as oracle
    from tests import golden_utils as gu

    dtype = torch.float32
    description = workloads.ares_config3(4096, dtype, 0, 4)
    particles = workloads.twiss_beam_particles(N)
    segment = workloads.product_segment(description, DEVICE, dtype)
    out = segment.track(workloads.product_beam(particles, DEVICE, dtype))
    assert out.particles.shape == (4, N, 7)
    truth = oracle.track(lattice_io.cast(lattice_io.cast(description, dtype), torch.float64),
                         workloads.oracle_beam(particles.to(dtype), torch.float64))
    assert gu.column_scaled_error(out.particles, truth["particles"]) < 2e-6
    mask = out.survival_probabilities.cpu().double()
    assert torch.equal(mask, truth["survival_probabilities"].expand_as(mask))
    assert 0.1 < float(mask.mean()) < 0.9
