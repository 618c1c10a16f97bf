"""Pin the non-linear part of the CPU oracle (drift_kick_drift / second_order tracking,
SURVEY.md 8f ranks 3-4) against the reference's golden pickles, the Bmad-X fixtures and fresh
outputs of the unmodified reference (tests/golden/nonlinear.npz, oracle/make_golden.py)."""

import pytest
import torch

from oracle import track_oracle as oracle

from . import golden_utils as gu

ARRAYS, CASES, ROW_STRIDE = gu.nonlinear_cases(torch.float64)
CONSISTENCY = gu.load_npz("consistency.npz")


def incoming_beam(case: str, dtype=torch.float64) -> tuple[dict, slice]:
    """Incoming beam of a case and the row slice its expected output was stored with."""
    prefix = CASES[case]["beam"]
    if prefix == "incoming":  # the reference's consistency beam
        return gu.beam_dict(CONSISTENCY, "incoming", dtype), slice(None, None, ROW_STRIDE)
    if prefix == "bmadx.incoming":  # stored subsampled, in and out
        return gu.beam_dict(ARRAYS, prefix, dtype), slice(None)
    return gu.beam_dict(ARRAYS, prefix, dtype), slice(None, None, ROW_STRIDE)


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("consistency.")))
def test_reference_pickles(case):
    """tests/test_elements.py:356-431 tolerance (allclose defaults, float64)."""
    beam, rows = incoming_beam(case)
    out = oracle.track(CASES[case]["lattice"], beam)
    expected = gu.beam_dict(ARRAYS, f"{case}.expected")
    assert torch.allclose(out["particles"][..., rows, :], expected["particles"])
    assert torch.allclose(out["energy"], expected["energy"])
    assert torch.allclose(out["s"], expected["s"])


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("bmadx.")))
def test_bmadx_fixtures(case):
    """Bmad-X results at the reference's float64 tolerance (1e-14, tests/test_drift.py:63-69)."""
    beam, rows = incoming_beam(case)
    out = oracle.track(CASES[case]["lattice"], beam)
    expected = gu.tensor(ARRAYS[f"{case}.expected.particles"])
    assert torch.allclose(out["particles"][..., rows, :], expected, atol=1e-14, rtol=1e-14)


@pytest.mark.parametrize("case", sorted(k for k in CASES if k.startswith("fresh.")))
def test_fresh_reference_outputs_float64(case):
    beam, rows = incoming_beam(case)
    out = oracle.track(CASES[case]["lattice"], beam)
    expected = gu.beam_dict(ARRAYS, f"{case}.f64")
    assert out["particles"][..., rows, :].shape == expected["particles"].shape
    assert torch.allclose(out["particles"][..., rows, :], expected["particles"], rtol=1e-11, atol=1e-15)
    assert torch.equal(
        out["survival_probabilities"][..., rows].expand(expected["survival_probabilities"].shape),
        expected["survival_probabilities"],
    )
    assert torch.allclose(out["energy"], expected["energy"], rtol=1e-14)
    assert torch.allclose(out["s"], expected["s"])


SECOND_ORDER_MAPS = gu.load_npz("second_order_maps.npz")
SECOND_ORDER_ELEMENTS = {
    "drift": {"type": "Drift", "length": 0.7},
    "quadrupole": {"type": "Quadrupole", "length": 0.2, "k1": [4.2, -3.1], "tilt": 0.3,
                   "misalignment": [2e-4, -1e-4]},
    "sextupole": {"type": "Sextupole", "length": 0.15, "k2": 25.0, "tilt": -0.2,
                  "misalignment": [1e-4, 3e-4]},
    "dipole": {"type": "Dipole", "length": 0.5, "angle": 0.2, "k1": 0.4, "dipole_e1": 0.05,
               "dipole_e2": 0.08, "tilt": 0.1, "fringe_integral": 0.5, "fringe_integral_exit": 0.4,
               "gap": 0.03},
    "rbend": {"type": "RBend", "length": 0.4, "angle": -0.15, "fringe_integral": 0.3, "gap": 0.02},
}


def second_order_element(name: str) -> dict:
    return {k: (torch.tensor(v, dtype=torch.float64) if not isinstance(v, str) else v)
            for k, v in SECOND_ORDER_ELEMENTS[name].items()} | {"name": name}


@pytest.mark.parametrize("name", sorted(SECOND_ORDER_ELEMENTS))
def test_dense_second_order_maps_match_the_reference(name):
    """``Element.second_order_transfer_map`` (element.py:134-147 and the per-class assemblies):
    the oracle's dense T against tensors computed by the unmodified reference."""
    from oracle import nonlinear_oracle

    energy = torch.tensor([6.3e7, 1.2e8], dtype=torch.float64)
    mass = torch.tensor(510998.95069, dtype=torch.float64)
    got = nonlinear_oracle.second_order_map(second_order_element(name), energy, mass)
    expected = gu.tensor(SECOND_ORDER_MAPS[name])
    assert got.shape == expected.shape == (2, 7, 7, 7)
    assert torch.allclose(got, expected, rtol=1e-10, atol=1e-13 * float(expected.abs().max()))
