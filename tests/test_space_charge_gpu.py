"""GPU parity of the SpaceChargeKick kernel family against the reference's own outputs
(tests/golden/space_charge.npz, cloud_in_cell.npz, consistency.npz) and the float64 oracle.

Tolerances.  The reference's own float32 result differs from its float64 result by ~1e-3 of
the kick (its IGF loses digits to cancellation in float32), so
  * float64 beams: every stage and the final kick within 1e-9 of the reference (relative to
    the stage's maximum);
  * float32 beams: kicks within 2e-3 x the largest kick of the float64 oracle evaluated on the
    same float32-rounded inputs (our Green function and SI conversions run in float64, so this
    path is CLOSER to the float64 truth than the reference's float32 path);
  * cloud-in-cell with particles at bin centres: exactly torch.histogramdd.
"""

import pytest
import torch

from oracle import lattice_io
from oracle import track_oracle as oracle

from . import golden_utils as gu
from .test_oracle_golden import CONSISTENCY, LATTICES, ROW_STRIDE, SPACE_CHARGE

pytestmark = pytest.mark.gpu
DEVICE = "cuda"


def rel_err(actual, expected) -> float:
    actual = actual.detach().cpu().double()
    expected = expected.detach().cpu().double()
    return float((actual - expected).abs().max() / expected.abs().max())


def incoming(dtype):
    return gu.beam_dict(SPACE_CHARGE, "incoming", dtype)


@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
def test_every_stage_matches_the_reference(tag, dtype):
    from cheetah_b200 import space_charge

    beam = incoming(dtype)
    to = lambda t: t.to(DEVICE)  # noqa: E731
    out, ws = space_charge.kick(
        to(beam["particles"]), to(beam["energy"]), to(beam["particle_charges"]),
        to(beam["survival_probabilities"]), to(beam["mass_eV"]),
        torch.tensor(0.7, dtype=dtype, device=DEVICE),
        tuple(torch.tensor(3.0, dtype=dtype, device=DEVICE) for _ in range(3)),
        (16, 16, 16), want_intermediates=True,
    )
    torch.cuda.synchronize()
    f64 = dtype == torch.float64
    golden = lambda name: gu.tensor(SPACE_CHARGE[f"kick16.{tag}.{name}"])  # noqa: E731
    truth = lambda name: gu.tensor(SPACE_CHARGE[f"kick16.f64.{name}"])  # noqa: E731

    assert rel_err(ws.params[:, 0:3], golden("grid_dimensions")) < (1e-12 if f64 else 1e-6)
    density = ws.charge_grid().double() * ws.params[:, 9, None, None, None]
    assert rel_err(density, golden("rho_padded")[:, :16, :16, :16]) < (1e-12 if f64 else 1e-5)
    # Green function: ours is evaluated in float64, so compare with the float64 reference
    assert rel_err(ws.green, truth("green")) < (1e-10 if f64 else 1e-6)
    assert rel_err(ws.phi, truth("potential")) < (1e-10 if f64 else 2e-5)
    rows = slice(None, None, 4)
    assert rel_err(ws.forces[:, rows], truth("forces")) < (1e-9 if f64 else 2e-4)
    expected = gu.beam_dict(SPACE_CHARGE, "kick16.f64")
    kick_ours = out.cpu().double()[0, rows] - beam["particles"].double()[rows]
    kick_ref = expected["particles"] - gu.beam_dict(SPACE_CHARGE, "incoming")["particles"][rows]
    for col in (1, 3, 5):
        err = (kick_ours[:, col] - kick_ref[:, col]).abs().max() / kick_ref[:, col].abs().max()
        assert err < (1e-9 if f64 else 2e-3), (col, float(err))
    for col in (0, 2, 4, 6):  # positions untouched
        assert torch.equal(out.cpu()[0, :, col], beam["particles"][:, col])


def _track_case(case, dtype):
    import cheetah_b200 as cb

    beam = incoming(dtype)
    if case == "kick16":
        lattice = [{"type": "SpaceChargeKick", "name": "sc", "effect_length": torch.tensor(0.7),
                    "grid_shape": (16, 16, 16)}]
    elif case == "kick32vec":
        lattice = [{"type": "SpaceChargeKick", "name": "sc", "effect_length": torch.tensor(0.2)}]
        beam["energy"] = torch.tensor(5e6, dtype=dtype)
        beam["particle_charges"] = beam["particle_charges"] * torch.tensor([[1.0], [3.0]], dtype=dtype)
    else:
        lattice = lattice_io.load(gu.GOLDEN / "fodo_space_charge_lattice.json", dtype)
        beam["energy"] = torch.tensor(5e7, dtype=dtype)
        beam["particle_charges"] = beam["particle_charges"] * 0.1
    lattice = lattice_io.cast(lattice, dtype)
    out = gu.product_segment(lattice, DEVICE, dtype).track(gu.product_beam(beam, DEVICE, dtype))
    return lattice, beam, out


@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
@pytest.mark.parametrize("case", ["kick16", "kick32vec", "fodo"])
def test_track_matches_reference_outputs(case, tag, dtype):
    lattice, beam, out = _track_case(case, dtype)
    rows = slice(None, None, 4)
    f64 = dtype == torch.float64
    truth = gu.beam_dict(SPACE_CHARGE, f"{case}.f64")
    assert out.particles.shape[:-2] == truth["particles"].shape[:-2]
    # compare DISPLACEMENTS: ours from the (float32-rounded) input it was given, the float64
    # reference's from the float64 input, so that input rounding does not count as kick error
    start = beam["particles"].double()[rows]
    start64 = gu.beam_dict(SPACE_CHARGE, "incoming")["particles"][rows]
    ours = out.particles.cpu().double()[..., rows, :] - start + start64
    moved = (truth["particles"] - start64).abs().amax(dim=-2, keepdim=True)
    # coordinates a kick leaves alone only "move" by the reference's rounding (tau -> -z/beta)
    floor = 1e-6 * truth["particles"].abs().amax(dim=-2, keepdim=True)
    err = ((ours - truth["particles"]).abs() / torch.maximum(moved, floor).clamp_min(1e-300))
    err = err[..., :6].max()
    # errors are relative to how far each coordinate moved through the whole lattice
    # float64: the reference's own delta = (gamma' - gamma0) / (beta0 gamma0) cancels ~8 digits
    assert err < (1e-7 if f64 else 3e-3), float(err)
    assert torch.equal(
        out.survival_probabilities.cpu().double()[..., rows], truth["survival_probabilities"]
    )
    assert torch.allclose(out.s.cpu().double(), truth["s"], rtol=1e-6)


def test_reference_consistency_pickle_space_charge_kick():
    """The reference's own golden pickle for SpaceChargeKick (32^3 default grid), float64,
    at the reference's own tolerance (tests/test_elements.py:356-431)."""
    beam = gu.beam_dict(CONSISTENCY, "incoming")
    case = "SpaceChargeKick_default"
    out = gu.product_segment(LATTICES[case], DEVICE, torch.float64).track(
        gu.product_beam(beam, DEVICE, torch.float64)
    )
    expected = gu.beam_dict(CONSISTENCY, f"{case}.expected")
    rows = slice(None, None, ROW_STRIDE)
    assert torch.allclose(out.particles.cpu()[rows], expected["particles"])
    assert out.survival_probabilities is not None and out.particles.shape == (3000, 7)


def test_float32_track_is_at_least_as_close_to_float64_truth_as_the_reference():
    """Our float32 kick vs the reference's float32 kick, both measured against float64."""
    _, beam, out = _track_case("kick16", torch.float32)
    rows = slice(None, None, 4)
    truth = gu.beam_dict(SPACE_CHARGE, "kick16.f64")["particles"]
    ref32 = gu.beam_dict(SPACE_CHARGE, "kick16.f32")["particles"]
    ours = out.particles.cpu().double()[rows]
    for col in (1, 3, 5):
        ours_err = (ours[:, col] - truth[:, col]).abs().max()
        ref_err = (ref32[:, col] - truth[:, col]).abs().max()
        assert ours_err <= 1.5 * ref_err + 1e-12, (col, float(ours_err), float(ref_err))


def test_input_not_mutated_and_passthrough():
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_parameters(
        num_particles=20_000, total_charge=torch.tensor(1e-9), device=DEVICE, dtype=torch.float32
    )
    before = beam.particles.clone()
    kick = cb.SpaceChargeKick(effect_length=torch.tensor(1.0, device=DEVICE))
    out = kick.track(beam)
    assert torch.equal(beam.particles, before)  # tests/test_space_charge_kick.py:171-199
    assert out.particle_charges is beam.particle_charges
    assert out.energy is beam.energy and out.s is beam.s
    assert not torch.equal(out.particles[:, 1], beam.particles[:, 1])
    with pytest.raises(AssertionError, match="only supported for `ParticleBeam`"):
        kick.track(cb.ParameterBeam(mu=torch.zeros(7, device=DEVICE), cov=torch.zeros(7, 7, device=DEVICE),
                                    energy=torch.tensor(1e8, device=DEVICE)))
    with pytest.raises(NotImplementedError, match="grid sizes in"):
        cb.SpaceChargeKick(effect_length=torch.tensor(1.0, device=DEVICE), grid_shape=(300, 32, 32)).track(beam)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_cloud_in_cell_matches_reference_and_histogram(dtype):
    from cheetah_b200.space_charge import cloud_in_cell_charge_deposition

    data = gu.load_npz("cloud_in_cell.npz")
    tag = "f64" if dtype == torch.float64 else "f32"
    grid = cloud_in_cell_charge_deposition(
        gu.tensor(data["positions"], dtype).to(DEVICE), tuple(int(b) for b in data["bins"]),
        gu.tensor(data["extent"], dtype).to(DEVICE), gu.tensor(data["charges"], dtype).to(DEVICE),
    )
    expected = gu.tensor(data[f"grid.{tag}"], dtype)
    assert grid.shape == expected.shape
    assert rel_err(grid, expected) < (1e-13 if dtype == torch.float64 else 2e-6)

    # particles on bin centres: exactly the histogram (tests/test_cloud_in_cell.py:98-150)
    g = torch.Generator().manual_seed(0)
    bins = (8, 4, 16)
    idx = torch.stack([torch.randint(0, b, (5000,), generator=g) for b in bins], dim=-1)
    extent = torch.tensor([[-1.0, 1.0], [0.0, 2.0], [-4.0, 4.0]], dtype=dtype)
    width = (extent[:, 1] - extent[:, 0]) / torch.tensor(bins, dtype=dtype)
    positions = extent[:, 0] + (idx.to(dtype) + 0.5) * width
    hist = torch.histogramdd(positions.double(), bins=list(bins), range=extent.double().flatten().tolist()).hist
    grid = cloud_in_cell_charge_deposition(positions.to(DEVICE), bins, extent.to(DEVICE))
    assert torch.equal(grid.cpu().double(), hist)
    # particles outside the extent contribute nothing (tests/test_cloud_in_cell.py:244-259)
    outside = torch.tensor([[0.0, 1.0, 0.0], [5.0, 1.0, 0.0], [0.0, -1.0, 0.0], [0.5, 0.5, 1.0]], dtype=dtype)
    grid = cloud_in_cell_charge_deposition(outside.to(DEVICE), bins, extent.to(DEVICE))
    assert float(grid.sum()) == 2.0


@pytest.mark.parametrize("dims", [1, 2])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_cloud_in_cell_1d_and_2d(dims, dtype):
    """cloud_in_cell_charge_deposition for 1 and 2 position dimensions (cloud_in_cell.py:67-241):
    reference outputs, exact histogram at bin centres, missing extent / charges, exact total."""
    from cheetah_b200.space_charge import cloud_in_cell_charge_deposition

    data = gu.load_npz("cloud_in_cell.npz")
    tag = "f64" if dtype == torch.float64 else "f32"
    bins = tuple(int(b) for b in data["bins"])[:dims]
    grid = cloud_in_cell_charge_deposition(
        gu.tensor(data["positions"], dtype)[..., :dims].to(DEVICE), bins,
        gu.tensor(data["extent"], dtype)[..., :dims, :].to(DEVICE),
        gu.tensor(data["charges"], dtype).to(DEVICE),
    )
    expected = gu.tensor(data[f"grid{dims}d.{tag}"], dtype)
    assert grid.shape == expected.shape
    assert rel_err(grid, expected) < (1e-13 if dtype == torch.float64 else 2e-6)

    g = torch.Generator().manual_seed(1)
    bins = (8, 16)[:dims]
    idx = torch.stack([torch.randint(0, b, (4000,), generator=g) for b in bins], dim=-1)
    extent = torch.tensor([[-1.0, 1.0], [0.0, 4.0]], dtype=dtype)[:dims]
    width = (extent[:, 1] - extent[:, 0]) / torch.tensor(bins, dtype=dtype)
    positions = extent[:, 0] + (idx.to(dtype) + 0.5) * width
    hist = torch.histogramdd(positions.double(), bins=list(bins),
                             range=extent.double().flatten().tolist()).hist
    grid = cloud_in_cell_charge_deposition(positions.to(DEVICE), bins, extent.to(DEVICE))
    assert torch.equal(grid.cpu().double(), hist)
    # extent inferred from the positions, integer bin count (cloud_in_cell.py:31-38)
    inferred = cloud_in_cell_charge_deposition(positions.to(DEVICE), 8)
    assert tuple(inferred.shape) == (8,) * dims
    span = torch.stack([positions.amin(dim=-2), positions.amax(dim=-2)], dim=-1)
    expected = oracle.cic_deposit_nd(positions, (8,) * dims, span, torch.ones(4000, dtype=dtype))
    # ~500 float32 additions per bin in a different order than scatter_add_
    assert rel_err(inferred, expected) < (1e-13 if dtype == torch.float64 else 2e-5)


# ---- fused stages: kick + following linear section + moments of the next kick ------------------
def _fodo_with_kicks(dtype, vector_k1=None, aperture=False, grid=(16, 16, 16)):
    import cheetah_b200 as cb

    t = lambda v: torch.tensor(v, device=DEVICE, dtype=dtype)  # noqa: E731
    elements = []
    for cell in range(2):
        for sign in (1.0, -1.0):
            k1 = t(sign * 4.2) if vector_k1 is None else t([sign * k for k in vector_k1])
            elements += [
                cb.Quadrupole(length=t(0.2), k1=k1),
                cb.Drift(length=t(0.5)),
                cb.SpaceChargeKick(effect_length=t(1.0), grid_shape=grid),
                cb.Drift(length=t(0.5)),
            ]
            if aperture and sign > 0:
                elements.append(cb.Aperture(x_max=t(5e-4), y_max=t(5e-4)))
    return cb.Segment(elements)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", ["plain", "vector_beams", "vector_lattice", "aperture"])
def test_fused_kick_pipeline_equals_the_stage_by_stage_one(case, dtype):
    """ch_sc_gather_kick_fused (kick + next linear map + next kick's moments in one pass) against
    the unfused stage sequence.  The only arithmetic difference is the reference point of the
    moment sums (origin instead of a pilot particle), i.e. rounding of sigma."""
    import cheetah_b200 as cb
    from cheetah_b200 import _capi, tracking

    torch.manual_seed(9)
    kwargs = {}
    if case in ("vector_beams", "vector_lattice"):
        kwargs["total_charge"] = torch.tensor([1e-10, 3e-10, 5e-10])
    beam = cb.ParticleBeam.from_parameters(
        num_particles=20_000, energy=torch.tensor(5e7), device=DEVICE, dtype=dtype, **kwargs
    )
    if case == "vector_beams":  # per-beam particles
        scale = torch.tensor([1.0, 1.1, 0.9], device=DEVICE, dtype=dtype).reshape(3, 1, 1)
        particles = beam.particles.unsqueeze(0) * scale
        particles[..., 6] = 1.0
        beam = cb.ParticleBeam(particles, beam.energy, particle_charges=beam.particle_charges,
                               species=beam.species)
    segment = _fodo_with_kicks(
        dtype, vector_k1=[4.2, 3.0, 5.0] if case == "vector_lattice" else None,
        aperture=case == "aperture",
    )
    launches = []
    results = []
    for fuse in (False, True):
        tracking.fuse_space_charge = fuse
        try:
            before = _capi.launch_count()
            results.append(segment.track(beam))
            torch.cuda.synchronize()
            launches.append(_capi.launch_count() - before)
        finally:
            tracking.fuse_space_charge = True
    unfused, fused = results
    assert fused.particles.shape == unfused.particles.shape
    assert launches[1] < launches[0]  # fewer passes over the particles
    tol = 1e-9 if dtype == torch.float64 else 5e-5
    scale = unfused.particles.abs().amax(dim=-2, keepdim=True).clamp_min(1e-30)
    assert float(((fused.particles - unfused.particles).abs() / scale).max()) < tol
    assert torch.equal(fused.survival_probabilities, unfused.survival_probabilities)
    assert torch.allclose(fused.s, unfused.s)


@pytest.mark.parametrize("energy", [2.5e8, 1e6], ids=["ultra-relativistic", "non-relativistic"])
def test_cold_uniform_bunch_doubles_in_size(energy):
    """Known-answer test of the reference (tests/test_space_charge_kick.py:14-69; free expansion
    of a cold uniform bunch): after the drift length
    L = beta gamma kappa sqrt(R0^3 / (N r_e)), kappa = 1 + sqrt(2)/4 ln(3 + 2 sqrt(2)), with three
    space-charge kicks, all three beam sizes have doubled within 2 %."""
    import math

    import cheetah_b200 as cb

    r0, total_charge = 1e-3, 1e-8
    electron_radius = 2.8179403205e-15
    gamma = energy / 510998.95069
    beta = math.sqrt(1 - 1 / gamma ** 2)
    incoming = cb.ParticleBeam.uniform_3d_ellipsoid(
        num_particles=100_000, radius_x=r0, radius_y=r0, radius_tau=r0 / gamma / beta,
        sigma_px=1e-15, sigma_py=1e-15, sigma_p=1e-15, energy=torch.tensor(energy),
        total_charge=torch.tensor(total_charge), device=DEVICE, dtype=torch.float32,
        generator=torch.Generator().manual_seed(2),
    )
    kappa = 1 + math.sqrt(2) / 4 * math.log(3 + 2 * math.sqrt(2))
    n_electrons = total_charge / 1.602176634e-19
    length = beta * gamma * kappa * math.sqrt(r0 ** 3 / (n_electrons * electron_radius))
    t = lambda v: torch.tensor(v, device=DEVICE, dtype=torch.float32)  # noqa: E731
    segment = cb.Segment([
        cb.Drift(length=t(length / 6)), cb.SpaceChargeKick(effect_length=t(length / 3)),
        cb.Drift(length=t(length / 3)), cb.SpaceChargeKick(effect_length=t(length / 3)),
        cb.Drift(length=t(length / 3)), cb.SpaceChargeKick(effect_length=t(length / 3)),
        cb.Drift(length=t(length / 6)),
    ])
    outgoing = segment.track(incoming)
    for name in ("sigma_x", "sigma_y", "sigma_tau"):
        ratio = float(getattr(outgoing, name) / getattr(incoming, name))
        assert abs(ratio / 2.0 - 1.0) < 2e-2, (name, ratio)


@pytest.mark.parametrize("grid", [(16, 16, 16), (32, 8, 64)])
def test_brick_gather_equals_node_gather(grid):
    """float32: ch_sc_field_bricks + the quad-cooperative gather against ch_sc_field + the
    one-thread-per-particle gather (same trilinear weights and node fields; the sums run in a
    different order), including particles outside the grid and on its boundary cells."""
    from cheetah_b200 import space_charge

    torch.manual_seed(4)
    dtype = torch.float32
    n = 30_011  # ragged last tile
    sigma = torch.tensor([2e-4, 3e-6, 1e-4, 5e-6, 1e-5, 1e-3], dtype=dtype)
    particles = torch.randn((2, n, 7), dtype=dtype) * torch.cat([sigma, torch.ones(1)])
    particles[..., 6] = 1.0
    particles[0, :64, 0] *= 5.0  # far outside the 3-sigma grid
    to = lambda t: t.to(DEVICE)  # noqa: E731
    args = (
        to(particles), torch.tensor(4e7, dtype=dtype, device=DEVICE),
        torch.full((2, n), 1e-10 / n, dtype=dtype, device=DEVICE),
        torch.ones((2, n), dtype=dtype, device=DEVICE),
        torch.tensor(510998.95, dtype=dtype, device=DEVICE),
        torch.tensor(0.4, dtype=dtype, device=DEVICE),
        tuple(torch.tensor(3.0, dtype=dtype, device=DEVICE) for _ in range(3)), grid,
    )
    results = {}
    for bricks in (True, False):
        space_charge.use_field_bricks = bricks
        try:
            out, ws = space_charge.kick(*args, want_intermediates=True)
            torch.cuda.synchronize()
            assert ws.bricks == bricks
            results[bricks] = (out.clone(), ws.forces.clone())
        finally:
            space_charge.use_field_bricks = True
    force_scale = results[False][1].abs().max()
    assert float((results[True][1] - results[False][1]).abs().max() / force_scale) < 5e-6
    kick_nodes = results[False][0] - to(particles)
    kick_bricks = results[True][0] - to(particles)
    for col in (1, 3, 5):
        scale = kick_nodes[..., col].abs().max()
        assert float((kick_bricks[..., col] - kick_nodes[..., col]).abs().max() / scale) < 1e-5
    for col in (0, 2, 4, 6):
        assert torch.equal(results[True][0][..., col], to(particles)[..., col])


def test_far_field_green_function_matches_the_exact_one(monkeypatch):
    """float32 beams take the Green function of far cells from the 4th-order series of the cell
    integral; compare the mirrored array with the one built from exact 8-corner differences
    (float64 beam, same geometry)."""
    from cheetah_b200 import space_charge

    torch.manual_seed(5)
    n = 20_000
    grid = (32, 32, 32)
    greens = {}
    sigma = torch.tensor([2e-4, 3e-6, 1.5e-4, 5e-6, 8e-6, 1e-3], dtype=torch.float64)
    particles = torch.randn((n, 7), dtype=torch.float64) * torch.cat([sigma, torch.ones(1)])
    particles[..., 6] = 1.0
    for dtype in (torch.float32, torch.float64):
        out, ws = space_charge.kick(
            particles.to(DEVICE, dtype), torch.tensor(1e8, dtype=dtype, device=DEVICE),
            torch.full((n,), 1e-10 / n, dtype=dtype, device=DEVICE),
            torch.ones((n,), dtype=dtype, device=DEVICE),
            torch.tensor(510998.95, dtype=dtype, device=DEVICE),
            torch.tensor(1.0, dtype=dtype, device=DEVICE),
            tuple(torch.tensor(3.0, dtype=dtype, device=DEVICE) for _ in range(3)), grid,
            want_intermediates=True,
        )
        torch.cuda.synchronize()
        greens[dtype] = ws.green.double().cpu()
        cells = ws.params[0, 3:6].cpu()
    # the two beams have (almost) the same sigma, hence the same cells up to float32 rounding:
    # compare point by point relative to the local value (the far field is much smaller than G[0])
    exact, approx = greens[torch.float64], greens[torch.float32]
    mask = exact != 0
    rel = ((approx - exact).abs() / exact.abs().clamp_min(1e-300))[mask]
    assert float(rel.max()) < 5e-6, float(rel.max())  # sigma rounding dominates (1e-6)
    # aspect ratio of this beam: most of the lattice must have been far field
    gamma = 1e8 / 510998.95
    h = torch.tensor([float(cells[0]), float(cells[1]), float(cells[2]) * gamma])
    idx = torch.stack(torch.meshgrid(*(torch.arange(32.0),) * 3, indexing="ij"), dim=-1)
    far = ((idx * h) ** 2).sum(-1) >= (6.0 * h.max()) ** 2
    assert float(far.float().mean()) > 0.5


@pytest.mark.parametrize("tag,dtype", [("f64", torch.float64), ("f32", torch.float32)])
@pytest.mark.parametrize("grid", [(20, 12, 48), (33, 17, 5), (3, 7, 130)])
def test_arbitrary_grid_shapes_match_the_oracle(grid, tag, dtype):
    """The reference takes any ``grid_shape`` (space_charge_kick.py:57, :148); here the FFT length
    of an axis is the next power of two >= 2 n, which leaves the aperiodic convolution -- and so
    the kick -- unchanged."""
    import cheetah_b200 as cb

    torch.manual_seed(13)
    n = 40_000
    sigma = torch.tensor([2e-4, 3e-6, 1.5e-4, 5e-6, 2e-5, 1e-3], dtype=torch.float64)
    particles = torch.randn((n, 7), dtype=torch.float64) * torch.cat([sigma, torch.ones(1)])
    particles[..., 6] = 1.0
    particles = particles.to(dtype)
    charges = torch.full((n,), 2e-10 / n, dtype=dtype)
    element = {"type": "SpaceChargeKick", "name": "sc", "effect_length": torch.tensor(0.5),
               "grid_shape": grid}
    beam64 = oracle.make_beam(particles.double(), torch.tensor(4e7, dtype=torch.float64),
                              particle_charges=charges.double())
    truth = oracle.track_space_charge(lattice_io.cast([element], torch.float64)[0], beam64)
    kick = cb.SpaceChargeKick(effect_length=torch.tensor(0.5, dtype=dtype, device=DEVICE),
                              grid_shape=grid)
    beam = cb.ParticleBeam(particles.to(DEVICE), torch.tensor(4e7, dtype=dtype, device=DEVICE),
                           particle_charges=charges.to(DEVICE),
                           species=cb.Species("electron", device=DEVICE, dtype=dtype))
    out = kick.track(beam)
    ours = out.particles.cpu().double() - particles.double()
    ref = truth["particles"] - particles.double()
    for col in (1, 3, 5):
        err = (ours[:, col] - ref[:, col]).abs().max() / ref[:, col].abs().max()
        assert err < (1e-8 if dtype == torch.float64 else 3e-3), (col, float(err))
    for col in (0, 2, 4, 6):
        assert torch.equal(out.particles.cpu()[:, col], particles[:, col])


@pytest.fixture()
def deterministic_algorithms():
    previous = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True, warn_only=True)
    yield
    torch.use_deterministic_algorithms(previous)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_deterministic_deposit_is_bit_reproducible(deterministic_algorithms, dtype):
    """With torch.use_deterministic_algorithms (what the reference's test suite sets,
    tests/conftest.py:204) the deposits accumulate in fixed point: every run gives the same bits,
    whatever order the atomics retire in; the values agree with the float atomics to rounding."""
    from cheetah_b200 import space_charge

    torch.manual_seed(21)
    n = 200_000
    sigma = torch.tensor([2e-4, 3e-6, 1.5e-4, 5e-6, 2e-5, 1e-3], dtype=torch.float64)
    particles = (torch.randn((3, n, 7), dtype=torch.float64)
                 * torch.cat([sigma, torch.ones(1)])).to(dtype)
    particles[..., 6] = 1.0
    charges = (torch.rand((3, n), dtype=torch.float64) * 2e-15).to(dtype)
    args = (
        particles.to(DEVICE), torch.tensor(4e7, dtype=dtype, device=DEVICE), charges.to(DEVICE),
        torch.rand((3, n), dtype=torch.float64).to(dtype).to(DEVICE),
        torch.tensor(510998.95, dtype=dtype, device=DEVICE),
        torch.tensor(0.4, dtype=dtype, device=DEVICE),
        tuple(torch.tensor(3.0, dtype=dtype, device=DEVICE) for _ in range(3)), (32, 32, 32),
    )
    grids, outs = [], []
    for _ in range(3):
        out, ws = space_charge.kick(*args)
        torch.cuda.synchronize()
        grids.append(ws.charge_grid().clone())
        outs.append(out.clone())
    assert torch.equal(grids[0], grids[1]) and torch.equal(grids[0], grids[2])
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    torch.use_deterministic_algorithms(False)
    _, ws = space_charge.kick(*args)
    atomic = ws.charge_grid()
    scale = atomic.abs().max()
    assert float((atomic - grids[0]).abs().max() / scale) < (1e-5 if dtype == torch.float32 else 1e-12)
    assert torch.allclose(grids[0].sum(dim=(1, 2, 3)).double(), atomic.sum(dim=(1, 2, 3)).double(),
                          rtol=1e-5)

    # general cloud-in-cell entry (1, 2 and 3 dimensions)
    torch.use_deterministic_algorithms(True, warn_only=True)
    for dims in (1, 2, 3):
        positions = torch.rand((2, 50_000, dims), dtype=dtype, device=DEVICE)
        weights = torch.rand((2, 50_000), dtype=dtype, device=DEVICE) - 0.3  # signed charges
        bins = (17, 9, 12)[:dims]
        runs = [space_charge.cloud_in_cell_charge_deposition(positions, bins, None, weights)
                for _ in range(2)]
        assert torch.equal(runs[0], runs[1])
        torch.use_deterministic_algorithms(False)
        reference_run = space_charge.cloud_in_cell_charge_deposition(positions, bins, None, weights)
        torch.use_deterministic_algorithms(True, warn_only=True)
        tol = 2e-5 if dtype == torch.float32 else 1e-12
        assert float((runs[0] - reference_run).abs().max() / reference_run.abs().max()) < tol


def test_concurrent_tracks_on_two_streams_keep_their_own_maps():
    """The fused section's maps travel through ONE constant-memory buffer per device; it belongs
    to the first stream that uses it and other streams fall back to shared memory, so two host
    threads tracking different lattices at the same time cannot see each other's maps."""
    import threading

    import cheetah_b200 as cb

    dtype = torch.float32
    torch.manual_seed(17)
    beam = cb.ParticleBeam.from_parameters(
        num_particles=200_000, total_charge=torch.tensor(5e-10), energy=torch.tensor(5e7),
        device=DEVICE, dtype=dtype,
    )
    segments = [_fodo_with_kicks(dtype, vector_k1=[k1]) for k1 in (4.2, -6.0)]
    expected = [segment.track(beam).particles.clone() for segment in segments]  # default stream
    torch.cuda.synchronize()
    results = [None, None]
    streams = [torch.cuda.Stream(DEVICE) for _ in range(2)]

    def work(i):
        with torch.cuda.stream(streams[i]):
            for _ in range(3):
                results[i] = segments[i].track(beam).particles
        streams[i].synchronize()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for got, want in zip(results, expected):
        scale = want.abs().amax(dim=-2, keepdim=True).clamp_min(1e-30)
        assert float(((got - want).abs() / scale).max()) < 5e-5
    # the two lattices really differ
    assert float((expected[0] - expected[1]).abs().max()) > 1e-6
