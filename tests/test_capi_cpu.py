"""CPU-side checks of the boundary: the library loads and exports every declared symbol, and
the host logic (lowering, shapes, errors) behaves without a GPU."""

import re
from pathlib import Path

import pytest
import torch

import cheetah_b200 as cb
from cheetah_b200 import _capi, lowering

REPO = Path(__file__).resolve().parent.parent


def declared_symbols() -> set:
    header = (REPO / "include" / "cheetah_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return set(re.findall(r"\b(ch_[a-z0-9_]+)\s*\(", header))


def test_library_exports_every_declared_symbol():
    from cheetah_b200 import build

    build.build()
    lib = _capi.lib()
    symbols = declared_symbols()
    assert symbols, "no symbols parsed from the header"
    assert symbols == set(_capi.SIGNATURES), symbols ^ set(_capi.SIGNATURES)
    for name in symbols:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.ch_abi_version() == _capi.ABI_VERSION == 3
    assert lib.ch_kernel_launch_count() == 0  # nothing computes on a CPU-only box


def test_program_create_validates_arguments_without_a_gpu():
    import ctypes

    lib = _capi.lib()
    handle = ctypes.c_void_p()
    opcodes = (ctypes.c_int32 * 1)(99)
    flags = (ctypes.c_int32 * 1)(0)
    begin = (ctypes.c_int32 * 2)(0, 0)
    status = lib.ch_program_create(opcodes, flags, begin, 1, None, None, None, 0, None,
                                   ctypes.byref(handle))
    assert status == -1
    assert b"unknown opcode" in lib.ch_last_error()
    with pytest.raises(_capi.BackendError, match="unknown opcode"):
        _capi.check(status)


def test_lowering_groups_runs_apertures_and_barriers():
    t = torch.tensor
    elements = [
        cb.Drift(length=t(0.5)),
        cb.Quadrupole(length=t(0.2), k1=t([1.0, 2.0, 3.0])),
        cb.Marker(),
        cb.Aperture(x_max=t(1e-3), y_max=t(2e-3), shape="elliptical"),
        cb.Segment([cb.Drift(length=t(0.1)), cb.BPM()]),
        cb.SpaceChargeKick(effect_length=t(1.0)),
        cb.Aperture(is_active=False),
        cb.Drift(length=t(0.5), tracking_method="drift_kick_drift"),
        cb.HorizontalCorrector(length=t(0.1), angle=t(1e-4)),
    ]
    program = lowering.lower(elements, torch.device("cpu"))
    kinds = [type(s).__name__ for s in program.stages]
    assert kinds == ["LinearSection", "Barrier", "LinearSection", "NonlinearRun", "LinearSection"]
    first = program.stages[0]
    assert (first.op_begin, first.op_end) == (0, 6)
    assert first.n_apertures == 1 and first.elliptical_mask == 1
    assert first.lattice_shape == (3,) and first.length_shape == () and first.survival_shape == (3,)
    assert program.stages[1].kind == "space_charge"
    assert (program.stages[3].op_begin, program.stages[3].op_end) == (7, 8)
    assert program.ops[7].opcode == _capi.OP_DKD_DRIFT
    assert [op.opcode for op in program.ops[:6]] == [
        _capi.OP_DRIFT, _capi.OP_QUADRUPOLE, _capi.OP_IDENTITY, _capi.OP_APERTURE,
        _capi.OP_DRIFT, _capi.OP_IDENTITY,
    ]
    # vectorised k1 is referenced in place (stride 1), scalars broadcast (stride 0)
    quad = program.ops[1]
    assert [stride for _, stride, _ in quad.resolved] == [0, 1, 0, 0, 0]
    assert quad.resolved[1][0].data_ptr() == elements[1].k1.data_ptr()
    assert [offset for _, _, offset in quad.resolved] == [0, 0, 0, 0, 1]


def test_lowering_groups_nonlinear_runs():
    """drift_kick_drift / second_order elements form runs; identity elements in between join the
    run, trailing ones open the next linear section."""
    t = torch.tensor
    elements = [
        cb.Drift(length=t(0.5), tracking_method="drift_kick_drift"),
        cb.Marker(),
        cb.Quadrupole(length=t(0.2), k1=t([1.0, 2.0]), num_steps=4,
                      tracking_method="drift_kick_drift"),
        cb.Sextupole(length=t(0.1), k2=t(3.0)),
        cb.BPM(),
        cb.Marker(),
        cb.Drift(length=t(0.3)),
        cb.Dipole(length=t(0.4), angle=t(0.1), fringe_at="exit",
                  tracking_method="drift_kick_drift"),
        cb.TransverseDeflectingCavity(length=t(0.2), voltage=t(1e6)),
        cb.Dipole(length=t(0.4), angle=t(0.1), tracking_method="second_order"),
    ]
    program = lowering.lower(elements, torch.device("cpu"))
    assert [type(s).__name__ for s in program.stages] == [
        "NonlinearRun", "LinearSection", "NonlinearRun"]
    first, middle, last = program.stages
    assert (first.op_begin, first.op_end) == (0, 4)
    assert (middle.op_begin, middle.op_end) == (4, 7) and middle.has_maps
    assert (last.op_begin, last.op_end) == (7, 10)
    assert first.lattice_shape == (2,) and first.length_shape == ()
    assert first.methods == ("drift_kick_drift", "drift_kick_drift", "second_order")
    assert [op.opcode for op in program.ops] == [
        _capi.OP_DKD_DRIFT, _capi.OP_IDENTITY, _capi.OP_DKD_QUADRUPOLE, _capi.OP_SECOND_ORDER,
        _capi.OP_IDENTITY, _capi.OP_IDENTITY, _capi.OP_DRIFT,
        _capi.OP_DKD_DIPOLE, _capi.OP_DKD_TDC, _capi.OP_SECOND_ORDER,
    ]
    assert program.ops[2].flags == 4                     # num_steps
    assert program.ops[7].flags == 2                     # fringe at the exit only
    assert program.ops[9].flags == 1                     # bend
    assert len(program.ops[3].resolved) == 12 and len(program.ops[8].resolved) == 7


def test_lowering_rejects_grad_and_wrong_device():
    k1 = torch.tensor(1.0, requires_grad=True)
    with pytest.raises(NotImplementedError, match="forward-only"):
        lowering.lower([cb.Quadrupole(length=torch.tensor(1.0), k1=k1)], torch.device("cpu"))
    with pytest.raises(ValueError, match="move the lattice"):
        lowering.lower([cb.Drift(length=torch.tensor(1.0))], torch.device("cuda:0"))


def test_cavity_skippability_is_watched():
    cavity = cb.Cavity(length=torch.tensor(1.0))
    program = lowering.lower([cavity], torch.device("cpu"))
    assert isinstance(program.stages[0], lowering.LinearSection) and not program.is_stale()
    cavity.voltage.fill_(1e6)
    assert program.is_stale()
    # an active cavity closes its linear section and contributes the non-linear tail
    active = lowering.lower([cb.Drift(length=torch.tensor(1.0)), cavity, cb.Marker()],
                            torch.device("cpu"))
    assert [type(s).__name__ for s in active.stages] == ["LinearSection", "LinearSection"]
    assert active.stages[0].cavity[0] is cavity and active.stages[1].cavity is None
    assert [op.opcode for op in active.ops] == [_capi.OP_DRIFT, _capi.OP_CAVITY, _capi.OP_IDENTITY]


def test_element_api_mirrors_the_reference():
    quad = cb.Quadrupole(length=torch.tensor(1.0), k1=torch.tensor(2.0), name="Q1")
    assert quad.tracking_method == "linear" and quad.is_skippable
    with pytest.warns(cb.PhysicsWarning):
        quad.tracking_method = "nonsense"
    assert quad.tracking_method == "linear"
    quad.tracking_method = "drift_kick_drift"
    assert not quad.is_skippable
    segment = cb.Segment([quad, cb.Drift(length=torch.tensor(0.5), name="D1")], name="cell")
    assert segment.Q1 is quad
    assert torch.equal(segment.length, torch.tensor(1.5))
    rbend = cb.RBend(length=torch.tensor(1.0), angle=torch.tensor(0.2), rbend_e1=torch.tensor(0.05))
    assert torch.allclose(rbend.dipole_e1, torch.tensor(0.15))
    assert torch.allclose(rbend.rbend_e1, torch.tensor(0.05))
    dipole = cb.Dipole(length=torch.tensor(1.0), fringe_integral=torch.tensor(0.5))
    assert torch.equal(dipole.fringe_integral_exit, torch.tensor(0.5))
    with pytest.raises(AssertionError):
        cb.CustomTransferMap(torch.ones(7, 7))
    assert cb.Sextupole(length=torch.tensor(1.0)).tracking_method == "second_order"
    beam = cb.ParticleBeam.from_twiss(num_particles=100, beta_x=3.14, beta_y=42.0)
    assert beam.particles.shape == (100, 7) and bool((beam.particles[:, 6] == 1).all())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        segment.track(beam)


def test_epoch_invalidation_on_attribute_assignment():
    from cheetah_b200.elements import lattice_epoch

    quad = cb.Quadrupole(length=torch.tensor(1.0))
    before = lattice_epoch()
    quad.k1 = torch.tensor(3.0)
    assert lattice_epoch() > before
    before = lattice_epoch()
    quad.to(torch.float64)
    assert lattice_epoch() > before and quad.k1.dtype == torch.float64


def test_segment_container_helpers_mirror_the_reference():
    """segment.py:73-229, :599-629: names, index, subcell, reversed, partition_at."""
    t = torch.tensor
    segment = cb.Segment([
        cb.Drift(length=t(1.0), name="d1"), cb.Quadrupole(length=t(0.2), k1=t(1.0), name="q1"),
        cb.Marker(name="m"), cb.Drift(length=t(0.5), name="d2"),
    ], name="cell")
    assert segment.element_names == ["d1", "q1", "m", "d2"]
    assert segment.element_index("q1") == 1
    with pytest.raises(ValueError, match="not found in segment"):
        segment.element_index("nope")
    assert segment.subcell("q1", "m").element_names == ["q1", "m"]
    assert segment.subcell(end="q1", include_end=False).element_names == ["d1"]
    assert segment.subcell(start="m", include_start=False).element_names == ["d2"]
    with pytest.raises(ValueError, match="not part of the segment"):
        segment.subcell(start="nope")
    assert segment.reversed().element_names == ["d2", "m", "q1", "d1"]
    pre, element, post = segment.partition_at("q1")
    assert pre.element_names == ["d1"] and element.name == "q1" and post.element_names == ["m", "d2"]
    pre, post = segment.partition_at("q1", mode="before")
    assert pre.element_names == ["d1"] and post.element_names == ["q1", "m", "d2"]
    with pytest.raises(AssertionError, match="not skippable"):
        cb.CustomTransferMap.from_merging_elements(
            [cb.Drift(length=t(1.0), tracking_method="drift_kick_drift")], incoming_beam=None)


def test_from_lattice_json_reads_the_reference_format(tmp_path):
    """Segment.from_lattice_json (segment.py:370-384, latticejson.py:156-260)."""
    import json
    from pathlib import Path

    lattice = {
        "version": "cheetah-0.7", "title": "test", "info": "", "root": "cell",
        "elements": {
            "d1": ["Drift", {"length": 0.5, "tracking_method": "linear"}],
            "q1": ["Quadrupole", {"length": 0.2, "k1": [1.0, -2.0], "misalignment": [0.0, 1e-4],
                                  "tilt": 0.1, "num_steps": 3, "tracking_method": "drift_kick_drift",
                                  "metadata": {"pv": "Q1:K1"}}],
            "a1": ["Aperture", {"x_max": 1e-3, "y_max": 2e-3, "shape": "elliptical",
                                "is_active": True}],
            "s1": ["Screen", {"resolution": [64, 48], "pixel_size": [1e-5, 2e-5], "binning": 2,
                              "method": "histogram", "is_active": False}],
            "bpm": ["BPM", {"is_active": False}],
            "sup": ["Superimposed", {"base_element": "q1", "superimposed_element": "bpm"}],
        },
        "lattices": {"cell": ["d1", "inner", "s1"], "inner": ["q1", "a1", "sup"]},
    }
    path = tmp_path / "lattice.json"
    path.write_text(json.dumps(lattice))
    segment = cb.Segment.from_lattice_json(path, dtype=torch.float64)
    assert segment.name == "cell" and [e.name for e in segment.elements] == ["d1", "inner", "s1"]
    inner = segment.elements[1]
    assert isinstance(inner, cb.Segment) and inner.q1.tracking_method == "drift_kick_drift"
    assert inner.q1.k1.dtype == torch.float64 and inner.q1.k1.tolist() == [1.0, -2.0]
    assert inner.q1.num_steps == 3 and inner.q1.metadata == {"pv": "Q1:K1"}
    assert inner.a1.shape == "elliptical" and float(inner.a1.y_max) == 2e-3
    assert segment.s1.resolution == (64, 48) and segment.s1.method == "histogram"
    assert isinstance(inner.sup, cb.Superimposed) and len(inner.sup.flattened().elements) == 3
    assert torch.allclose(segment.length, torch.tensor(0.5 + 0.2 + 0.2, dtype=torch.float64))

    ares = Path("/root/reference/docs/examples/ARESlatticeStage3v1_9.json")
    if ares.exists():  # only in the build container; the GPU box has the converted fixture
        from tests import golden_utils as gu

        loaded = cb.Segment.from_lattice_json(ares).flattened().elements
        golden = gu.ares_lattice(torch.float32)
        assert [type(e).__name__ for e in loaded] == [d["type"] for d in golden]
        assert [e.name for e in loaded] == [d["name"] for d in golden]
        for element, description in zip(loaded, golden):
            if "length" in description:
                assert torch.equal(element.length, description["length"])
