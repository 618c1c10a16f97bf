"""The lowering is duck-typed: it must accept the REFERENCE's own element objects
(INTEGRATION.md section 2).  Runs only where /root/reference exists (the build container)."""

import sys
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
REPO = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not available")


@pytest.fixture(scope="module")
def cheetah():
    sys.path.insert(0, str(REPO / "oracle" / "refshim"))
    sys.path.insert(0, str(REF))
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import cheetah as reference
    yield reference
    sys.path.remove(str(REF))
    sys.path.remove(str(REPO / "oracle" / "refshim"))


def test_reference_ares_lowers_to_the_same_program(cheetah):
    from cheetah_b200 import lowering

    from . import golden_utils as gu

    reference_segment = cheetah.Segment.from_lattice_json(
        str(REF / "docs" / "examples" / "ARESlatticeStage3v1_9.json")
    )
    reference_segment.AREAMQZM1.k1 = torch.tensor([8.2, 1.0])
    mirror_description = gu.ares_lattice(torch.float32)
    gu.set_attr(mirror_description, "AREAMQZM1", "k1", torch.tensor([8.2, 1.0]))
    mirror_segment = gu.product_segment(mirror_description, "cpu", torch.float32)

    a = lowering.lower([reference_segment], torch.device("cpu"))
    b = lowering.lower([mirror_segment], torch.device("cpu"))
    assert [op.opcode for op in a.ops] == [op.opcode for op in b.ops]
    assert [op.flags for op in a.ops] == [op.flags for op in b.ops]
    assert len(a.ops) == 195
    assert [type(s).__name__ for s in a.stages] == [type(s).__name__ for s in b.stages]
    sa, sb = a.stages[0], b.stages[0]
    assert (sa.n_apertures, sa.lattice_shape, sa.length_shape) == (3, (2,), ())
    assert (sb.n_apertures, sb.lattice_shape, sb.length_shape) == (3, (2,), ())
    for op_a, op_b in zip(a.ops, b.ops):
        for (ta, stride_a, off_a), (tb, stride_b, off_b) in zip(op_a.resolved, op_b.resolved):
            assert stride_a == stride_b and off_a == off_b
            assert torch.equal(ta.double(), tb.double())


def test_reference_space_charge_and_superimposed_lower(cheetah):
    from cheetah_b200 import lowering

    t = torch.tensor
    elements = [
        cheetah.Superimposed(
            base_element=cheetah.Quadrupole(length=t(1.0), k1=t(0.5)),
            superimposed_element=cheetah.BPM(),
        ),
        cheetah.SpaceChargeKick(effect_length=t(1.0)),
        cheetah.RBend(length=t(1.0), angle=t(0.2), rbend_e1=t(0.05)),
        cheetah.Cavity(length=t(1.0), voltage=t(1e6)),
    ]
    program = lowering.lower(elements, torch.device("cpu"))
    kinds = [getattr(s, "kind", "linear") for s in program.stages]
    # the active cavity joins (and closes) the bend's linear section
    assert kinds == ["linear", "space_charge", "linear"]
    assert program.stages[2].cavity[0] is elements[3]
    assert len(program.ops) == 5  # two half quadrupoles + BPM, the bend, the cavity
    dipole = program.ops[3]
    assert torch.allclose(dipole.resolved[3][0], t(0.15))  # e1 = rbend_e1 + angle / 2


def test_clone_and_lattice_simplifications():
    """Element / Segment.clone (element.py:323-336) and the reference's lattice simplifications
    (segment.py:231-330), host-side only."""
    import cheetah_b200 as cb

    t = torch.tensor
    quad = cb.Quadrupole(length=t(0.2), k1=t([1.0, -2.0]), tilt=t(0.1), num_steps=3,
                         tracking_method="drift_kick_drift", name="q1")
    copy = quad.clone()
    assert copy is not quad and copy.name == "q1" and copy.num_steps == 3
    assert copy.tracking_method == "drift_kick_drift"
    assert torch.equal(copy.k1, quad.k1) and copy.k1.data_ptr() != quad.k1.data_ptr()
    segment = cb.Segment([
        cb.Drift(length=t(0.5), name="d1"), cb.Marker(name="m1"), quad,
        cb.BPM(name="bpm_off"), cb.BPM(name="bpm_on", is_active=True),
        cb.Screen(name="screen_off"), cb.Cavity(length=t(1.0), voltage=t(0.0), name="cav_off"),
        cb.Segment([cb.Marker(name="inner_marker"), cb.HorizontalCorrector(length=t(0.1), name="hc")],
                   name="inner"),
    ], name="line")
    cloned = segment.clone()
    assert [e.name for e in cloned.elements] == [e.name for e in segment.elements]
    assert cloned.q1 is not segment.q1 and torch.equal(cloned.q1.k1, segment.q1.k1)
    assert [e.name for e in segment.without_inactive_markers().elements] == [
        "d1", "q1", "bpm_off", "bpm_on", "screen_off", "cav_off", "inner"]
    assert [e.name for e in segment.without_inactive_markers(except_for=["m1"]).elements][1] == "m1"
    assert [e.name for e in segment.without_inactive_zero_length_elements().elements] == [
        "d1", "q1", "bpm_on", "cav_off", "inner"]
    as_drifts = segment.inactive_elements_as_drifts(except_for=["q1"])
    kinds = {e.name: type(e).__name__ for e in as_drifts.elements}
    assert kinds["q1"] == "Quadrupole" and kinds["cav_off"] == "Drift" and kinds["d1"] == "Drift"
    assert kinds["bpm_on"] == "BPM" and kinds["m1"] == "Marker"
    assert torch.equal(as_drifts.length, segment.length)
    segment.set_attrs_on_every_element(cb.BPM, is_active=True)
    assert segment.bpm_off.is_active
    segment.set_attrs_on_every_element(cb.HorizontalCorrector, angle=t(1e-3))
    assert abs(float(segment.inner.hc.angle) - 1e-3) < 1e-9


def test_merging_consecutive_elements():
    """Element.merge / Segment.with_consecutive_elements_merged (drift.py:175-187,
    quadrupole.py:280-301, sextupole.py:133-152, solenoid.py:143-157, segment.py:326-367)."""
    import cheetah_b200 as cb

    t = torch.tensor
    segment = cb.Segment([
        cb.Drift(length=t(0.5), name="D1"), cb.Drift(length=t(0.25), name="D2"),
        cb.Quadrupole(length=t(0.1), k1=t(2.0), name="Q_in", num_steps=2),
        cb.Quadrupole(length=t(0.3), k1=t(4.0), name="Q_out", num_steps=3),
        cb.Quadrupole(length=t(0.3), k1=t(4.0), tilt=t(0.1), name="QT"),
        cb.Sextupole(length=t(0.1), k2=t(3.0), name="S1"), cb.Sextupole(length=t(0.1), k2=t(3.0), name="S2"),
        cb.Sextupole(length=t(0.1), k2=t(4.0), name="S3"),
        cb.Solenoid(length=t(0.2), k=t(1.0), name="sol_a"), cb.Solenoid(length=t(0.2), k=t(3.0), name="sol_b"),
        cb.Drift(length=t(0.1), name="D3"), cb.Drift(length=t(0.1), name="keep"),
    ], name="line")
    merged = segment.with_consecutive_elements_merged(except_for=["keep"])
    assert [e.name for e in merged.elements] == ["D", "Q_", "QT", "S", "S3", "sol_", "D3", "keep"]
    assert torch.isclose(merged.length, segment.length)
    assert torch.isclose(merged.Q_.k1, t(3.5)) and merged.Q_.num_steps == 5
    assert torch.isclose(merged.sol_.k, t(2.0))
    assert torch.isclose(merged.D.length, t(0.75))
    # different tracking methods or frames do not merge
    a = cb.Drift(length=t(0.1), tracking_method="drift_kick_drift")
    assert a.merge(cb.Drift(length=t(0.1))) is None
    assert a.merge(cb.Drift(length=t(0.2), tracking_method="drift_kick_drift")).tracking_method == \
        "drift_kick_drift"
    assert cb.Marker().merge(cb.Marker()) is None
    both = cb.Segment([a], name="cell_1").merge(cb.Segment([cb.Marker(name="m")], name="cell_2"))
    assert both.name == "cell_" and len(both.elements) == 2


def _sample_lattice():
    import cheetah_b200 as cb

    t = torch.tensor
    return cb.Segment([
        cb.Drift(length=t(0.5), name="d1", tracking_method="drift_kick_drift"),
        cb.Quadrupole(length=t(0.2), k1=t([1.0, -2.0]), tilt=t(0.1), misalignment=t([1e-4, 0.0]),
                      num_steps=3, name="q1"),
        cb.Dipole(length=t(0.4), angle=t(0.1), dipole_e1=t(0.02), fringe_integral=t(0.5), gap=t(0.03),
                  fringe_at="entrance", name="b1"),
        cb.Segment([cb.Marker(name="m1"), cb.HorizontalCorrector(length=t(0.1), angle=t(1e-3), name="h1")],
                   name="inner"),
        cb.Aperture(x_max=t(1e-3), y_max=t(2e-3), shape="elliptical", name="a1"),
        cb.Screen(resolution=(64, 48), pixel_size=t([1e-5, 2e-5]), binning=2, method="kde",
                  is_active=True, name="s1"),
        cb.SpaceChargeKick(effect_length=t(0.3), grid_shape=(16, 16, 8), name="sc1"),
        cb.Superimposed(cb.Drift(length=t(1.0), name="long_drift"), cb.Marker(name="mid"), name="sup1"),
    ], name="line")


def test_lattice_json_round_trip(tmp_path):
    """Segment.to_lattice_json / from_lattice_json (segment.py:370-396, latticejson.py)."""
    import cheetah_b200 as cb

    segment = _sample_lattice()
    path = tmp_path / "line.json"
    segment.to_lattice_json(path, title="test lattice")
    loaded = cb.Segment.from_lattice_json(path)
    assert loaded.name == "line"
    assert [type(e).__name__ for e in loaded.elements] == [type(e).__name__ for e in segment.elements]
    assert [e.name for e in loaded.elements] == [e.name for e in segment.elements]
    assert loaded.d1.tracking_method == "drift_kick_drift"
    assert torch.equal(loaded.q1.k1, segment.q1.k1) and loaded.q1.num_steps == 3
    assert loaded.b1.fringe_at == "entrance" and torch.equal(loaded.b1.dipole_e1, segment.b1.dipole_e1)
    assert [e.name for e in loaded.inner.elements] == ["m1", "h1"]
    assert loaded.a1.shape == "elliptical" and float(loaded.a1.y_max) == pytest.approx(2e-3)
    assert tuple(loaded.s1.resolution) == (64, 48) and loaded.s1.method == "kde" and loaded.s1.is_active
    assert tuple(loaded.sc1.grid_shape) == (16, 16, 8)
    assert type(loaded.sup1.superimposed_element).__name__ == "Marker"
    assert torch.isclose(loaded.length, segment.length)


def test_the_reference_reads_our_lattice_json(tmp_path):
    """The file we write loads in the unmodified reference (skipped where it is not installed)."""
    cheetah = pytest.importorskip("cheetah")
    segment = _sample_lattice()
    path = tmp_path / "line.json"
    segment.to_lattice_json(path)
    theirs = cheetah.Segment.from_lattice_json(str(path))
    assert [type(e).__name__ for e in theirs.elements] == [type(e).__name__ for e in segment.elements]
    assert torch.equal(theirs.q1.k1, segment.q1.k1)
    assert theirs.d1.tracking_method == "drift_kick_drift"
    assert torch.isclose(theirs.length, segment.length)
