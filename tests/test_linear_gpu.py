"""GPU parity of the linear path (ch_compose_maps + ch_apply_maps) through the public API.

Tolerances (north_star: "within a stated fp32 tolerance, bit-exact for aperture masks"):
  * float64 beams: the reference's own golden tolerance, torch.allclose rtol 1e-5 / atol 1e-8,
    and 1e-12 relative to the column maximum against the float64 oracle;
  * float32 beams: 2e-6 x column maximum against the float64 oracle (the reference's own
    fp32-vs-fp64 distance on ARES is 5e-7);
  * survival masks: exactly equal to the oracle's on identical inputs.
"""

import pytest
import torch

from oracle import track_oracle as oracle

from . import golden_utils as gu
from .test_oracle_golden import ARES, CONSISTENCY, LATTICES, ROW_STRIDE, ares_case

pytestmark = pytest.mark.gpu
DEVICE = "cuda"
F32_TOL = 2e-6
F64_TOL = 1e-12

LINEAR_CASES = sorted(k for k in LATTICES if not k.startswith("SpaceChargeKick"))


@pytest.mark.parametrize("case", LINEAR_CASES)
def test_reference_consistency_pickles_f64(case):
    """The reference's own golden pickles at the reference's own tolerance (float64)."""
    incoming = gu.beam_dict(CONSISTENCY, "incoming")
    segment = gu.product_segment(LATTICES[case], DEVICE, torch.float64)
    out = segment.track(gu.product_beam(incoming, DEVICE, torch.float64))
    rows = slice(None, None, ROW_STRIDE)
    expected = gu.beam_dict(CONSISTENCY, f"{case}.expected")
    assert torch.allclose(out.particles.cpu()[..., rows, :], expected["particles"])
    assert torch.allclose(
        out.survival_probabilities.cpu()[..., rows], expected["survival_probabilities"]
    )
    assert torch.allclose(out.s.cpu(), expected["s"])
    assert torch.allclose(out.energy.cpu(), expected["energy"])
    assert out.particles.shape[:-2] == expected["particles"].shape[:-2]


@pytest.mark.parametrize("case", LINEAR_CASES)
def test_reference_consistency_pickles_f32(case):
    incoming = gu.beam_dict(CONSISTENCY, "incoming")
    lattice = [dict(d) for d in LATTICES[case]]
    segment = gu.product_segment(lattice, DEVICE, torch.float32)
    out = segment.track(gu.product_beam(incoming, DEVICE, torch.float32))
    # oracle in float64 on the SAME float32-rounded inputs
    from oracle import lattice_io

    lattice32 = lattice_io.cast(lattice_io.cast(LATTICES[case], torch.float32), torch.float64)
    beam32 = {
        k: v.to(torch.float32).to(torch.float64) for k, v in incoming.items()
    }
    expected = oracle.track(lattice32, beam32)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < F32_TOL
    assert torch.equal(
        out.survival_probabilities.cpu().double(), expected["survival_probabilities"]
    )


@pytest.mark.parametrize("case", LINEAR_CASES)
def test_parameter_beam_goldens_f64(case):
    import cheetah_b200 as cb

    if f"{case}.expected.mu" not in CONSISTENCY:
        pytest.skip("no ParameterBeam pickle")
    incoming = gu.beam_dict(CONSISTENCY, "incoming")
    segment = gu.product_segment(LATTICES[case], DEVICE, torch.float64)
    beam = cb.ParameterBeam(
        mu=gu.tensor(CONSISTENCY["incoming.mu"]).to(DEVICE),
        cov=gu.tensor(CONSISTENCY["incoming.cov"]).to(DEVICE),
        energy=incoming["energy"].to(DEVICE),
        total_charge=gu.tensor(CONSISTENCY["incoming.total_charge"]).to(DEVICE),
        species=cb.Species(
            "custom",
            num_elementary_charges=incoming["num_elementary_charges"].to(DEVICE),
            mass_eV=incoming["mass_eV"].to(DEVICE),
        ),
    )
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = segment.track(beam)
    assert torch.allclose(out.mu.cpu(), gu.tensor(CONSISTENCY[f"{case}.expected.mu"]))
    assert torch.allclose(out.cov.cpu(), gu.tensor(CONSISTENCY[f"{case}.expected.cov"]))


@pytest.mark.parametrize("case", ["default", "config2", "vectorised"])
@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
def test_ares_against_reference_outputs(case, tag, dtype):
    """ARES (195 elements) against outputs of the unmodified reference (tests/golden/ares.npz)."""
    lattice, beam = ares_case(case, dtype)
    segment = gu.product_segment(lattice, DEVICE, dtype)
    out = segment.track(gu.product_beam(beam, DEVICE, dtype))
    rows = slice(None, None, 4)
    expected = gu.beam_dict(ARES, f"{case}.{tag}", dtype)
    # against the reference run in the same dtype: both carry their own rounding (the
    # reference's fp32 chain of 7x7 products is the noisier side: our fp32 result is ~1e-7
    # from the fp64 truth, the reference's fp32 up to ~7e-6 on the vectorised case)
    tol = 1e-11 if dtype == torch.float64 else 2e-5
    assert gu.column_scaled_error(out.particles[..., rows, :], expected["particles"]) < tol
    # against the float64 reference: our float32 path must be at the reference's own noise floor
    truth = gu.beam_dict(ARES, f"{case}.f64", torch.float64)
    assert gu.column_scaled_error(out.particles[..., rows, :], truth["particles"]) < (
        F64_TOL * 10 if dtype == torch.float64 else F32_TOL
    )
    flips = (
        out.survival_probabilities.cpu().double()[..., rows] != truth["survival_probabilities"]
    ).sum()
    assert flips == 0, f"{int(flips)} survival mask mismatches against the float64 reference"
    assert torch.allclose(out.s.cpu().double(), truth["s"], rtol=1e-6)
    assert out.particles.shape[:-2] == expected["particles"].shape[:-2]
    assert out.survival_probabilities.shape == (
        *expected["survival_probabilities"].shape[:-1],
        beam["particles"].shape[-2],
    )


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", ["rectangular", "elliptical"])
@pytest.mark.parametrize("n", [1000, 1001, 4099])
def test_aperture_masks_bit_exact(dtype, shape, n):
    """Aperture in isolation on identical inputs: masks equal the oracle's exactly, including
    particles placed on and one ulp around the edge."""
    import cheetah_b200 as cb

    g = torch.Generator().manual_seed(n)
    particles = torch.randn(3, n, 7, generator=g, dtype=torch.float64) * 1e-3
    particles[..., 6] = 1.0
    x_max = torch.tensor([[1e-3], [5e-4]], dtype=dtype)
    y_max = torch.tensor(8e-4, dtype=dtype)
    particles = particles.to(dtype)
    # edge cases: exactly on the edge and the neighbouring representable values
    edge = x_max[1, 0]
    particles[0, 0, 0] = edge
    particles[0, 1, 0] = torch.nextafter(edge, torch.tensor(0.0, dtype=dtype))
    particles[0, 2, 0] = torch.nextafter(edge, torch.tensor(1.0, dtype=dtype))
    particles[0, 3, 0] = -edge
    particles[0, :4, 2] = 0.0
    particles[1, 0, 2] = y_max
    particles[1, 0, 0] = 0.0
    survival = torch.rand(n, generator=g, dtype=torch.float64).to(dtype)

    beam = oracle.make_beam(particles, torch.tensor(1e8, dtype=dtype), survival_probabilities=survival)
    expected = oracle.track_aperture(
        {"type": "Aperture", "x_max": x_max, "y_max": y_max, "shape": shape}, beam
    )
    aperture = cb.Aperture(x_max=x_max.to(DEVICE), y_max=y_max.to(DEVICE), shape=shape)
    out = aperture.track(gu.product_beam(beam, DEVICE, dtype))
    assert out.survival_probabilities.shape == expected["survival_probabilities"].shape
    assert torch.equal(out.survival_probabilities.cpu(), expected["survival_probabilities"])
    # particles pass through an aperture unchanged, bit for bit, and are NOT widened by the
    # vectorised aperture limits (aperture.py:126-132 hands `incoming.particles` on)
    assert torch.equal(out.particles.cpu(), particles)


ELEMENT_CASES = {
    "Drift": {"length": [1.0, -1.0]},
    "Quadrupole": {"length": 0.7, "k1": [1.0, -2.0, 0.0], "tilt": 0.42, "misalignment": (0.01, -0.02)},
    "Dipole": {
        "length": 1.0, "angle": [1.0, -2.0], "k1": 0.3, "dipole_e1": 0.1, "dipole_e2": -0.2,
        "tilt": 0.42, "gap": 0.03, "fringe_integral": 0.5, "fringe_integral_exit": 0.4,
    },
    "RBend": {"length": 1.0, "angle": [1.0, -2.0, 0.0], "rbend_e1": 0.05, "tilt": 0.1},
    "HorizontalCorrector": {"length": 0.2, "angle": [1e-3, -2e-3]},
    "VerticalCorrector": {"length": 0.2, "angle": [1e-3, -2e-3]},
    "CombinedCorrector": {"length": 0.2, "horizontal_angle": [1e-3, -2e-3], "vertical_angle": 5e-4},
    "Solenoid": {"length": 0.5, "k": [1.0, -2.0, 0.0], "misalignment": (0.01, -0.02)},
    "Undulator": {"length": 1.0, "period": 0.1, "kx": 1.3, "ky": 0.4},
    "Cavity": {"length": 1.3},
    "Sextupole": {"length": 0.3, "k2": 2.0, "tracking_method": "linear"},
    "Marker": {},
}


@pytest.mark.parametrize("kind", sorted(ELEMENT_CASES))
@pytest.mark.parametrize("energy", [1e8, [6e6, 2.5e8]])
def test_first_order_transfer_map_matches_oracle(kind, energy):
    import cheetah_b200 as cb

    dtype = torch.float64
    kwargs = {
        k: (v if isinstance(v, str) else torch.tensor(v, dtype=dtype)) for k, v in ELEMENT_CASES[kind].items()
    }
    energy_t = torch.tensor(energy, dtype=dtype)
    if energy_t.dim() == 1:  # make the energy dimension broadcast against the parameter dimension
        energy_t = energy_t.unsqueeze(-1)
    description = {"type": kind, **kwargs}
    expected = oracle.first_order_map(
        description, energy_t, torch.tensor(oracle.ELECTRON_MASS_EV, dtype=dtype), torch.tensor(-1.0, dtype=dtype)
    )
    element = getattr(cb, kind)(
        **{k: (v.to(DEVICE) if isinstance(v, torch.Tensor) else v) for k, v in kwargs.items()}
    )
    species = cb.Species("electron", device=DEVICE, dtype=dtype)
    tm = element.first_order_transfer_map(energy_t.to(DEVICE), species)
    expected = expected.expand(tm.shape)
    assert tm.shape[-2:] == (7, 7)
    assert torch.allclose(tm.cpu(), expected, rtol=1e-11, atol=1e-13), (tm.cpu() - expected).abs().max()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_segment_of_every_linear_element_and_odd_sizes(dtype):
    """A lattice with every linear element type, tilted/misaligned magnets, two apertures, a
    particle count that is not a multiple of 4 (non-TMA path) and a non-unit 7th column."""
    import cheetah_b200 as cb

    g = torch.Generator().manual_seed(7)
    lattice = []
    for kind, params in ELEMENT_CASES.items():
        entry = {"type": kind, "name": kind}
        for k, v in params.items():
            if isinstance(v, str):
                entry[k] = v
            else:
                t = torch.tensor(v, dtype=torch.float64)
                entry[k] = t if t.dim() == 0 or k == "misalignment" else t[0]
        lattice.append(entry)
    lattice.insert(4, {"type": "Aperture", "name": "a1", "x_max": torch.tensor(2e-3), "y_max": torch.tensor(3e-3), "shape": "rectangular", "is_active": True})
    lattice.append({"type": "Aperture", "name": "a2", "x_max": torch.tensor(0.9), "y_max": torch.tensor(0.8), "shape": "elliptical", "is_active": True})
    lattice.append({"type": "Drift", "name": "tail", "length": torch.tensor(0.25)})
    from oracle import lattice_io

    lattice = lattice_io.cast(lattice, dtype)
    for n, unit in ((1003, True), (2048, False), (5, True)):
        particles = torch.randn(n, 7, generator=g, dtype=torch.float64) * 1e-3
        particles[:, 6] = 1.0 if unit else 0.5
        beam = oracle.make_beam(particles.to(dtype), torch.tensor(8e7, dtype=dtype))
        truth_lattice = lattice_io.cast(lattice, torch.float64)
        truth_beam = {k: v.to(torch.float64) for k, v in beam.items()}
        expected = oracle.track(truth_lattice, truth_beam)
        out = gu.product_segment(lattice, DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
        tol = F32_TOL if dtype == torch.float32 else F64_TOL
        assert gu.column_scaled_error(out.particles, expected["particles"]) < tol
        assert torch.equal(out.particles[..., 6].cpu().double(), expected["particles"][..., 6])
        mismatches = (out.survival_probabilities.cpu().double() != expected["survival_probabilities"]).sum()
        assert mismatches == 0
        assert torch.allclose(out.s.cpu().double(), expected["s"], rtol=1e-6)


def test_broadcast_shapes_follow_the_reference():
    """Vector dims: particles (2,1,N,7) x quad k1 (3,) -> (2,3,N,7); survival only widens up to
    the last aperture; `s` only carries the length dims (tests/test_vectorized.py:339-371)."""
    import cheetah_b200 as cb

    dtype = torch.float32
    n = 512
    g = torch.Generator().manual_seed(11)
    particles = torch.randn(2, 1, n, 7, generator=g, dtype=torch.float64) * 1e-3
    particles[..., 6] = 1.0
    lattice = [
        {"type": "Drift", "name": "d0", "length": torch.tensor(0.5)},
        {"type": "Aperture", "name": "ap", "x_max": torch.tensor(1e-3), "y_max": torch.tensor(1e-3), "shape": "rectangular", "is_active": True},
        {"type": "Quadrupole", "name": "q", "length": torch.tensor(0.2), "k1": torch.tensor([4.0, -3.0, 0.5]), "misalignment": torch.zeros(2), "tilt": torch.tensor(0.0)},
        {"type": "Drift", "name": "d1", "length": torch.tensor([1.0, 2.0, 3.0])},
    ]
    beam = oracle.make_beam(particles.to(dtype), torch.tensor(1e8, dtype=dtype))
    from oracle import lattice_io

    expected = oracle.track(lattice_io.cast(lattice, torch.float64), {k: v.double() for k, v in beam.items()})
    out = gu.product_segment(lattice_io.cast(lattice, dtype), DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
    assert out.particles.shape == expected["particles"].shape == (2, 3, n, 7)
    assert out.survival_probabilities.shape == expected["survival_probabilities"].shape == (2, 1, n)
    assert out.s.shape == expected["s"].shape == (3,)
    assert gu.column_scaled_error(out.particles, expected["particles"]) < F32_TOL
    assert torch.equal(out.survival_probabilities.cpu().double(), expected["survival_probabilities"])


def test_input_beam_is_not_mutated_and_outputs_are_new():
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_parameters(num_particles=4096, device=DEVICE, dtype=torch.float32)
    before = beam.particles.clone()
    segment = cb.Segment([cb.Drift(length=torch.tensor(1.0, device=DEVICE))])
    out = segment.track(beam)
    assert torch.equal(beam.particles, before)
    assert out.particles.data_ptr() != beam.particles.data_ptr()
    assert out.survival_probabilities is beam.survival_probabilities  # passed through (element.py:186-188)
    assert out.particle_charges is beam.particle_charges
    assert out.species is not beam.species and out.species.name == beam.species.name


def test_in_place_setting_updates_are_seen_without_relowering():
    import cheetah_b200 as cb

    quad = cb.Quadrupole(length=torch.tensor(0.2, device=DEVICE), k1=torch.tensor(1.0, device=DEVICE))
    segment = cb.Segment([quad, cb.Drift(length=torch.tensor(1.0, device=DEVICE))])
    beam = cb.ParticleBeam.from_parameters(num_particles=1024, device=DEVICE, dtype=torch.float32)
    a = segment.track(beam).particles.clone()
    quad.k1.fill_(-3.0)  # in place: same storage, program keeps pointing at it
    b = segment.track(beam).particles.clone()
    quad.k1 = torch.tensor(-3.0, device=DEVICE)  # re-assignment: new tensor -> re-lowered
    c = segment.track(beam).particles
    assert not torch.equal(a, b)
    assert torch.equal(b, c)


def test_errors_are_loud():
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_parameters(num_particles=16, dtype=torch.float32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cb.Drift(length=torch.tensor(1.0)).track(beam)
    class Wiggler(cb.Element):  # an element type this backend does not know
        @property
        def is_skippable(self) -> bool:
            return False

    with pytest.raises(NotImplementedError, match="outside the accelerated hot path"):
        Wiggler(name="w").track(
            cb.ParticleBeam.from_parameters(num_particles=16, device=DEVICE, dtype=torch.float32)
        )
    with pytest.raises(ValueError, match="move the lattice"):
        cb.Drift(length=torch.tensor(1.0)).track(
            cb.ParticleBeam.from_parameters(num_particles=16, device=DEVICE, dtype=torch.float32)
        )
    with pytest.raises(TypeError):
        cb.Drift(length=torch.tensor(1.0, device=DEVICE)).track("not a beam")


# ---- active cavities -------------------------------------------------------
@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
@pytest.mark.parametrize("case", ["standing", "traveling", "decelerating", "vectorised", "segment"])
def test_active_cavity_against_reference_outputs(case, tag, dtype):
    """Active Cavity.track (cheetah/accelerator/cavity.py:100-251): linear R + exact delta update
    + second-order tau terms, and the energy change seen by the elements downstream."""
    from .test_oracle_golden import CAVITY, cavity_lattices

    lattice = cavity_lattices(dtype)[case]
    beam = gu.beam_dict(CAVITY, "incoming", dtype)
    out = gu.product_segment(lattice, DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
    truth = gu.beam_dict(CAVITY, f"{case}.f64", torch.float64)
    rows = slice(None, None, 4)
    assert out.particles.shape[:-2] == truth["particles"].shape[:-2]
    # float32: compared with the float64 reference; delta carries the cos(phi + e) - cos(phi)
    # cancellation in the reference's own float32 path, ours is evaluated without it
    tol = 1e-10 if dtype == torch.float64 else 5e-6
    assert gu.column_scaled_error(out.particles[..., rows, :], truth["particles"]) < tol
    assert torch.allclose(out.energy.cpu().double(), truth["energy"], rtol=1e-12 if dtype == torch.float64 else 1e-6)
    assert torch.allclose(out.s.cpu().double(), truth["s"], rtol=1e-6)
    assert torch.equal(
        out.survival_probabilities.cpu().double()[..., rows], truth["survival_probabilities"]
    )


def test_transfer_maps_merged_and_beam_along_segment():
    """Segment.transfer_maps_merged / CustomTransferMap.from_merging_elements /
    beam_along_segment_generator (segment.py:179-229, :631-656, custom_transfer_map.py:60-109)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=torch.float64)  # noqa: E731
    segment = cb.Segment([
        cb.Drift(length=t(0.5), name="d1"),
        cb.Quadrupole(length=t(0.2), k1=t(4.0), tilt=t(0.1), name="q1"),
        cb.Drift(length=t(0.3), name="d2"),
        cb.Aperture(x_max=t(3e-4), y_max=t(3e-4), name="a1"),
        cb.HorizontalCorrector(length=t(0.1), angle=t(2e-4), name="h1"),
        cb.Quadrupole(length=t(0.2), k1=t(-4.0), name="q2"),
        cb.Drift(length=t(0.7), name="d3"),
    ])
    torch.manual_seed(1)
    beam = cb.ParticleBeam.from_parameters(num_particles=20_000, device=DEVICE, dtype=torch.float64)
    expected = segment.track(beam)
    merged = segment.transfer_maps_merged(beam)
    assert [type(e).__name__ for e in merged.elements] == [
        "CustomTransferMap", "Aperture", "CustomTransferMap"]
    assert merged.elements[0].name == "combined_d1_q1_d2"
    assert torch.allclose(merged.elements[0].length, t(1.0))
    out = merged.track(beam)
    assert torch.allclose(out.particles, expected.particles, rtol=1e-12, atol=1e-18)
    assert torch.equal(out.survival_probabilities, expected.survival_probabilities)
    assert torch.allclose(out.s, expected.s)
    kept = segment.transfer_maps_merged(beam, except_for=["q2"])
    # like the reference, a trailing run is wrapped even when it holds a single element
    assert [e.name for e in kept.elements] == ["combined_d1_q1_d2", "a1", "h1", "q2", "combined_d3"]
    assert torch.allclose(kept.track(beam).particles, expected.particles, rtol=1e-12, atol=1e-18)
    beams = list(segment.beam_along_segment_generator(beam))
    assert len(beams) == len(segment.elements) + 1 and beams[0] is beam
    assert torch.allclose(beams[-1].particles, expected.particles, rtol=1e-11, atol=1e-17)
    assert torch.allclose(torch.stack([b.s for b in beams])[-1], expected.s)


def test_superimposed_and_split_against_the_reference_pickle():
    """cheetah/accelerator/superimposed.py: a BPM at the centre of a quadrupole (the reference's
    Superimposed_ParticleBeam_default.pkl), and Segment.beam_along_segment_generator(resolution)."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=torch.float64)  # noqa: E731
    incoming = gu.product_beam(gu.beam_dict(CONSISTENCY, "incoming"), DEVICE, torch.float64)
    element = cb.Superimposed(
        base_element=cb.Quadrupole(length=t(1.0), k1=t(0.5)),
        superimposed_element=cb.BPM(misalignment=t([0.0, 0.0])), name="default",
    )
    assert [type(e).__name__ for e in element.flattened().elements] == [
        "Quadrupole", "BPM", "Quadrupole"]
    out = element.track(incoming)
    rows = slice(None, None, ROW_STRIDE)
    expected = gu.beam_dict(CONSISTENCY, "Superimposed_default.expected")
    assert torch.allclose(out.particles.cpu()[..., rows, :], expected["particles"])
    assert torch.allclose(out.s.cpu(), expected["s"])
    assert torch.allclose(
        element.first_order_transfer_map(incoming.energy, incoming.species),
        cb.Quadrupole(length=t(1.0), k1=t(0.5)).first_order_transfer_map(
            incoming.energy, incoming.species), rtol=1e-12, atol=1e-15)
    segment = cb.Segment([cb.Drift(length=t(1.0)), element, cb.Solenoid(length=t(0.5), k=t(0.3),
                                                                        misalignment=t([0.0, 0.0]))])
    beams = list(segment.beam_along_segment_generator(incoming, resolution=0.3))
    # 4 drift slices + the Superimposed (not splittable, element.py:338-347) + 2 solenoid slices
    assert len(beams) == 1 + 4 + 1 + 2
    assert torch.allclose(beams[-1].particles, segment.track(incoming).particles,
                          rtol=1e-10, atol=1e-16)
    assert torch.allclose(beams[-1].s, t(2.5))


@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
@pytest.mark.parametrize("case", ["standing", "traveling", "decelerating", "vectorised", "segment"])
def test_active_cavity_parameter_beam(case, tag, dtype):
    """ParameterBeam branch of Cavity.track (cavity.py:108-110, :129-135, :203-217) in
    ch_apply_maps_parameter, against the unmodified reference."""
    import warnings

    import cheetah_b200 as cb

    from .test_oracle_golden import CAVITY, cavity_lattices

    lattice = cavity_lattices(dtype)[case]
    incoming = gu.beam_dict(CAVITY, "incoming", dtype)
    to = lambda a: gu.tensor(a, dtype).to(DEVICE)  # noqa: E731
    beam = cb.ParameterBeam(
        mu=to(CAVITY["incoming.mu"]), cov=to(CAVITY["incoming.cov"]),
        energy=incoming["energy"].to(DEVICE),
        species=cb.Species("electron", device=DEVICE, dtype=dtype),
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the segment holds an aperture
        out = gu.product_segment(lattice, DEVICE, dtype).track(beam)
    truth_mu = gu.tensor(CAVITY[f"{case}.f64.mu"])
    truth_cov = gu.tensor(CAVITY[f"{case}.f64.cov"])
    assert out.mu.shape == truth_mu.shape and out.cov.shape == truth_cov.shape
    rtol = 1e-9 if dtype == torch.float64 else 2e-5
    scale_mu = truth_mu.abs().amax(dim=-1, keepdim=True)
    assert float(((out.mu.cpu().double() - truth_mu).abs() / scale_mu).max()) < rtol
    sigma = truth_cov.diagonal(dim1=-2, dim2=-1).abs().sqrt()
    scale_cov = (sigma.unsqueeze(-1) * sigma.unsqueeze(-2)).clamp_min(1e-300)
    assert float(((out.cov.cpu().double() - truth_cov).abs() / scale_cov)[..., :6, :6].max()) < (
        1e-8 if dtype == torch.float64 else 2e-4)
    assert torch.allclose(out.energy.cpu().double(),
                          gu.tensor(CAVITY[f"{case}.f64.parameter_energy"]),
                          rtol=1e-12 if dtype == torch.float64 else 1e-6)


@pytest.mark.parametrize("shape,expected", [("elliptical", [0.0235, 0.42, 0.552]),
                                            ("rectangular", [0.029, 0.495, 0.629])])
def test_vectorised_aperture_survival_fractions(shape, expected):
    """Known-answer test of the reference (tests/test_vectorized.py:461-504): a Gaussian beam
    (sigma_px = 2e-4, sigma_py = 1e-4) through Drift(0.5) - Aperture(x_max (3, 1), y_max 2e-4) -
    Drift(0.5) with two beam energies: broadcast shapes and the surviving fractions to 5e-3."""
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE)  # noqa: E731
    incoming = cb.ParticleBeam.from_parameters(
        num_particles=100_000, sigma_px=2e-4, sigma_py=1e-4, energy=t([154e6, 14e9]),
        device=DEVICE, generator=torch.Generator().manual_seed(7),
    )
    segment = cb.Segment([
        cb.Drift(length=t(0.5)),
        cb.Aperture(x_max=t([[1e-5], [2e-4], [3e-4]]), y_max=t(2e-4), shape=shape),
        cb.Drift(length=t(0.5)),
    ])
    outgoing = segment.track(incoming)
    assert tuple(outgoing.particles.shape) == (2, 100_000, 7)
    assert tuple(outgoing.energy.shape) == (2,)
    assert tuple(outgoing.particle_charges.shape) == (100_000,)
    assert tuple(outgoing.survival_probabilities.shape) == (3, 2, 100_000)
    fractions = outgoing.survival_probabilities.mean(dim=-1)[:, 0].cpu()
    assert torch.allclose(fractions, torch.tensor(expected), atol=5e-3), fractions


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_incoming_survival_with_vector_dims_of_its_own(dtype):
    """``survival_probabilities`` may carry vector dimensions that neither the particles nor the
    lattice have (reference contract: tests/test_vectorized.py:339-371, :488-491 -- the survival
    of the outgoing beam is the broadcast of everything, the particles keep their own shape)."""
    from oracle import lattice_io

    n = 3000
    g = torch.Generator().manual_seed(3)
    particles = torch.randn(n, 7, generator=g, dtype=torch.float64) * 1e-3
    particles[..., 6] = 1.0
    survival = torch.rand(3, 1, n, generator=g, dtype=torch.float64)
    lattice = [
        {"type": "Drift", "name": "d1", "length": torch.tensor(0.4)},
        {"type": "Quadrupole", "name": "q", "length": torch.tensor(0.2),
         "k1": torch.tensor([2.0, -3.0])},
        {"type": "Aperture", "name": "a", "x_max": torch.tensor(1.1e-3),
         "y_max": torch.tensor(9e-4), "shape": "elliptical"},
        {"type": "Drift", "name": "d2", "length": torch.tensor(0.3)},
    ]
    lattice = lattice_io.cast(lattice, dtype)
    beam = oracle.make_beam(particles.to(dtype), torch.tensor(1e8, dtype=dtype),
                            survival_probabilities=survival.to(dtype))
    expected = oracle.track(lattice, beam)
    out = gu.product_segment(lattice, DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
    assert out.particles.shape == expected["particles"].shape == (2, n, 7)
    assert out.survival_probabilities.shape == expected["survival_probabilities"].shape == (3, 2, n)
    assert torch.equal(out.survival_probabilities.cpu(), expected["survival_probabilities"])
    err = gu.column_scaled_error(out.particles, expected["particles"])
    assert err < (1e-12 if dtype == torch.float64 else 2e-6)


def test_survey_ranges_k1_30_parity():
    """SURVEY 8d's first recipe for config 3 -- k1 ~ U(-30, 30) 1/m^2, corrector angles ~
    U(-1e-3, 1e-3) rad -- which the bench narrows (these ranges over-focus ARES until the beam is
    kilometres wide).  As a parity case it exercises map entries up to ~1e9: the apertures are
    opened to 100 m instead (mean survival 56 %), float32 on the GPU against the
    float64 oracle on the same inputs."""
    import workloads
    from oracle import lattice_io

    n_settings, n = 64, 5000
    particles = workloads.twiss_beam_particles(n)
    lattice32 = workloads.ares_survey_ranges(n_settings, torch.float32, x_max=100.0)
    truth = oracle.track(
        lattice_io.cast(lattice_io.cast(lattice32, torch.float32), torch.float64),
        workloads.oracle_beam(particles.float(), torch.float64),
    )
    segment = workloads.product_segment(lattice32, DEVICE, torch.float32)
    out = segment.track(workloads.product_beam(particles, DEVICE, torch.float32))
    survival = truth["survival_probabilities"]
    assert 0.3 < float(survival.mean()) < 0.7
    assert torch.equal(out.survival_probabilities.cpu().double(), survival)
    assert float(truth["particles"].abs().max()) > 1e3  # the regime this test is about
    # per setting: error relative to the largest coordinate of that setting and column
    scale = truth["particles"].abs().amax(dim=-2, keepdim=True).clamp_min(1e-30)
    err = ((out.particles.cpu().double() - truth["particles"]).abs() / scale)[..., :6].max()
    assert float(err) < 2e-6, float(err)
