"""GPU parity of the per-particle non-linear tracking methods (ch_nonlinear_constants +
ch_track_nonlinear) through the public API.

Tolerances:
  * float64: the reference's golden pickles at the reference's tolerance (allclose defaults), the
    Bmad-X fixtures at the reference's 1e-14 (tests/test_drift.py:63-69), fresh reference outputs
    and the float64 oracle at 1e-11 relative;
  * float32: 3e-6 x max(column maximum, largest element misalignment) against the float64
    oracle on the same float32-rounded inputs.  The misalignment enters because a float32
    coordinate shifted by an offset of 1-2 cm (the reference's consistency cases) is rounded at
    the offset's magnitude, whatever the implementation.  The reference's own float32 run of
    the same cases is 10-1000x farther from float64 in the longitudinal columns, see
    test_float32_is_closer_to_float64_than_the_reference.
"""

import pytest
import torch

from oracle import lattice_io
from oracle import track_oracle as oracle

from . import golden_utils as gu
from .test_nonlinear_oracle import ARRAYS, CASES, incoming_beam

pytestmark = pytest.mark.gpu
DEVICE = "cuda"
F32_TOL = 3e-6


def run_product(case: str, dtype):
    beam, rows = incoming_beam(case)
    segment = gu.product_segment(CASES[case]["lattice"], DEVICE, dtype)
    out = segment.track(gu.product_beam(beam, DEVICE, dtype))
    return out, rows


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("consistency.")))
def test_reference_pickles_f64(case):
    out, rows = run_product(case, torch.float64)
    expected = gu.beam_dict(ARRAYS, f"{case}.expected")
    assert out.particles.shape[:-2] == expected["particles"].shape[:-2]
    assert torch.allclose(out.particles.cpu()[..., rows, :], expected["particles"])
    assert torch.allclose(out.energy.cpu(), expected["energy"])
    assert torch.allclose(out.s.cpu(), expected["s"])


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("bmadx.")))
def test_bmadx_fixtures_f64(case):
    out, rows = run_product(case, torch.float64)
    expected = gu.tensor(ARRAYS[f"{case}.expected.particles"])
    assert torch.allclose(out.particles.cpu()[..., rows, :], expected, atol=1e-14, rtol=1e-14)


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("bmadx.")))
def test_bmadx_fixtures_f32(case):
    """The reference's own float32 tolerance for these fixtures (atol 1e-5, rtol 1e-6)."""
    out, rows = run_product(case, torch.float32)
    expected = gu.tensor(ARRAYS[f"{case}.expected.particles"])
    assert torch.allclose(out.particles.cpu().double()[..., rows, :], expected, atol=1e-5, rtol=1e-6)


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("fresh.")))
def test_fresh_reference_outputs_f64(case):
    out, rows = run_product(case, torch.float64)
    expected = gu.beam_dict(ARRAYS, f"{case}.f64")
    got = out.particles.cpu()[..., rows, :]
    assert got.shape == expected["particles"].shape
    assert torch.allclose(got, expected["particles"], rtol=1e-11, atol=1e-15)
    survival = out.survival_probabilities.cpu()[..., rows]
    assert torch.equal(survival.expand(expected["survival_probabilities"].shape),
                       expected["survival_probabilities"])
    assert torch.allclose(out.energy.cpu(), expected["energy"], rtol=1e-14)
    assert torch.allclose(out.s.cpu(), expected["s"])


def _oracle_on_f32_inputs(case: str) -> dict:
    beam, _ = incoming_beam(case)
    lattice32 = lattice_io.cast(lattice_io.cast(CASES[case]["lattice"], torch.float32), torch.float64)
    beam32 = {k: v.to(torch.float32).to(torch.float64) for k, v in beam.items()}
    return oracle.track(lattice32, beam32)


def _largest_misalignment(lattice: list) -> float:
    return max(
        [float(d["misalignment"].abs().max()) for d in lattice if "misalignment" in d] + [0.0]
    )


@pytest.mark.parametrize("case", sorted(CASES))
def test_float32_against_float64_oracle(case):
    out, _ = run_product(case, torch.float32)
    expected = _oracle_on_f32_inputs(case)
    assert out.particles.dtype == torch.float32
    truth = expected["particles"]
    scale = truth.abs().amax(dim=-2, keepdim=True).clamp_min(
        _largest_misalignment(CASES[case]["lattice"])
    )
    error = ((out.particles.cpu().double() - truth).abs() / scale)[..., :6].max()
    assert float(error) < F32_TOL
    assert torch.equal(
        out.survival_probabilities.cpu().double().expand(expected["survival_probabilities"].shape),
        expected["survival_probabilities"],
    )


@pytest.mark.parametrize("case", ["fresh.drift_dkd_vector", "fresh.quadrupole_dkd_steps",
                                  "fresh.dipole_dkd_entrance", "fresh.tdc_vector"])
def test_float32_is_closer_to_float64_than_the_reference(case):
    """The reference's float32 run cancels in (p - p0c) / p0c and in the path-length
    differences; the kernel's cancellation-free forms stay at float32 rounding."""
    out, rows = run_product(case, torch.float32)
    truth = gu.beam_dict(ARRAYS, f"{case}.f64")["particles"]
    reference32 = gu.beam_dict(ARRAYS, f"{case}.f32")["particles"]
    ours = gu.column_scaled_error(out.particles.cpu()[..., rows, :], truth)
    theirs = gu.column_scaled_error(reference32, truth)
    assert ours <= theirs * 1.05 + 1e-7, (ours, theirs)


def test_parameter_beam_and_error_behaviour():
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    mu = torch.zeros(7, device=DEVICE)
    mu[6] = 1.0
    parameter_beam = cb.ParameterBeam(mu, torch.eye(7, device=DEVICE) * 1e-8, energy=t(1e8))
    with pytest.raises(AssertionError, match="Drift-kick-drift tracking"):
        cb.Drift(length=t(1.0), tracking_method="drift_kick_drift").track(parameter_beam)
    with pytest.raises(AssertionError, match="Second-order tracking"):
        cb.Sextupole(length=t(1.0), k2=t(2.0)).track(parameter_beam)
    beam = cb.ParticleBeam.from_parameters(num_particles=1000, device=DEVICE)
    drift = cb.Drift(length=t(1.0), tracking_method="drift_kick_drift")
    out = drift.track(beam)
    assert out.particles.data_ptr() != beam.particles.data_ptr()  # input never mutated
    assert out.survival_probabilities is beam.survival_probabilities
    assert out.particle_charges is beam.particle_charges
    assert torch.equal(out.s, beam.s + 1.0)


def test_long_run_is_split_and_matches_single_elements():
    """> CH_NL_MAX_OPS consecutive non-linear elements: several launches, same result as
    tracking element by element."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=torch.float64)  # noqa: E731
    torch.manual_seed(3)
    beam = cb.ParticleBeam.from_parameters(num_particles=5000, sigma_p=1e-3, device=DEVICE,
                                           dtype=torch.float64)
    elements = []
    for i in range(40):
        elements.append(cb.Drift(length=t(0.1), tracking_method="drift_kick_drift"))
        elements.append(cb.Quadrupole(length=t(0.05), k1=t(3.0 if i % 2 else -3.0), num_steps=2,
                                      tracking_method="drift_kick_drift"))
    fused = cb.Segment(elements).track(beam)
    step = beam
    for element in elements:
        step = element.track(step)
    assert torch.allclose(fused.particles, step.particles, rtol=1e-10, atol=1e-16)
    assert torch.allclose(fused.s, step.s)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("k1,length,num_steps", [
    (4.2, 0.2, 5),      # |k1| step^2 = 0.0067: power series
    (-35.0, 0.5, 2),    # 2.19: series, just inside the switch-over
    (35.0, 0.5, 2),
    (40.0, 0.5, 2),     # 2.5: closed forms (cosh / sinh, cos / sin)
    (-120.0, 0.6, 1),   # 43: closed forms, k l = 6.6 rad
    (0.0, 0.3, 3),      # k1 = 0: a drift
])
def test_drift_kick_drift_quadrupole_series_and_closed_forms(k1, length, num_steps, dtype):
    """quadrupole_plane evaluates cos / cosh and sin / sinh by a power series in k1 l^2 for
    |k1 l^2| <= 2.25 and by the closed forms (quadrupole.py:168-251, bmadx.py:223-260) above;
    both against the oracle, with vectorised k1 straddling the switch-over."""
    g = torch.Generator().manual_seed(11)
    n = 4000
    particles = torch.randn(n, 7, generator=g, dtype=torch.float64)
    particles[:, :6] *= torch.tensor([3e-4, 4e-5, 3e-4, 4e-5, 1e-4, 2e-2], dtype=torch.float64)
    particles[:, 6] = 1.0
    k1s = torch.tensor([k1, 0.97 * k1, 1.06 * k1], dtype=torch.float64)
    lattice = [{"type": "Quadrupole", "name": "q", "length": torch.tensor(length, dtype=torch.float64),
                "k1": k1s, "num_steps": num_steps, "tracking_method": "drift_kick_drift"}]
    beam = oracle.make_beam(particles, torch.tensor(6e6, dtype=torch.float64))
    if dtype == torch.float32:
        lattice = lattice_io.cast(lattice_io.cast(lattice, torch.float32), torch.float64)
        beam = {k: v.to(torch.float32).to(torch.float64) for k, v in beam.items()}
    expected = oracle.track(lattice, beam)
    out = gu.product_segment(lattice, DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
    assert tuple(out.particles.shape) == (3, n, 7)
    tolerance = 1e-11 if dtype == torch.float64 else F32_TOL
    assert gu.column_scaled_error(out.particles, expected["particles"]) < tolerance


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["drift", "quadrupole", "sextupole", "dipole", "rbend"])
def test_dense_second_order_transfer_map(name, dtype):
    """``Element.second_order_transfer_map``: the fp64 coefficient table of
    ``ch_nonlinear_constants`` scattered into the reference's dense (7, 7, 7) layout with the
    frames folded in, against tensors from the unmodified reference."""
    import cheetah_b200 as cb

    from .test_nonlinear_oracle import SECOND_ORDER_MAPS, second_order_element

    element = gu.product_segment([second_order_element(name)], DEVICE, dtype).elements[0]
    energy = torch.tensor([6.3e7, 1.2e8], device=DEVICE, dtype=dtype)
    got = element.second_order_transfer_map(energy, cb.Species("electron", device=DEVICE, dtype=dtype))
    expected = gu.tensor(SECOND_ORDER_MAPS[name])
    assert got.dtype == dtype and tuple(got.shape) == (2, 7, 7, 7)
    error = (got.cpu().double() - expected).abs()
    # entry by entry, relative to the largest coefficient that shares its output row
    scale = expected.abs().amax(dim=(-2, -1), keepdim=True).clamp_min(1e-300)
    assert float((error / scale).max()) < (1e-10 if dtype == torch.float64 else 2e-6)
    with pytest.raises(NotImplementedError):
        cb.Marker().second_order_transfer_map(energy, cb.Species("electron", device=DEVICE, dtype=dtype))
