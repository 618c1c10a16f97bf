"""Helpers to read the committed golden fixtures (see oracle/make_golden.py)."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import torch

from oracle import lattice_io

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_npz(name: str) -> dict:
    with np.load(GOLDEN / name) as data:
        return {key: data[key] for key in data.files}


def tensor(array, dtype=torch.float64) -> torch.Tensor:
    return torch.as_tensor(np.asarray(array), dtype=dtype)


def beam_dict(arrays: dict, prefix: str, dtype=torch.float64) -> dict:
    """Oracle-style beam dict from ``<prefix>.*`` arrays."""
    return {
        key: tensor(arrays[f"{prefix}.{key}"], dtype)
        for key in (
            "particles",
            "energy",
            "particle_charges",
            "survival_probabilities",
            "s",
            "mass_eV",
            "num_elementary_charges",
        )
    }


def consistency_lattices(dtype=torch.float64) -> tuple[dict, int]:
    with (GOLDEN / "consistency.json").open() as f:
        raw = json.load(f)
    lattices = {
        key: lattice_io._from_json(value, dtype) for key, value in raw["lattices"].items()
    }
    return lattices, raw["row_stride"]


def ares_lattice(dtype=torch.float32) -> list:
    return lattice_io.load(GOLDEN / "ares_lattice.json", dtype)


def set_attr(description: list, name: str, attr: str, value) -> None:
    for element in description:
        if element["name"] == name:
            element[attr] = value
            return
    raise KeyError(name)


# ---- product-side helpers (need the CUDA library + a GPU) ---------------------------------
def product_segment(description: list, device, dtype):
    import cheetah_b200

    return cheetah_b200.Segment(
        elements=lattice_io.build(description, cheetah_b200, device=device, dtype=dtype)
    )


def product_beam(beam: dict, device, dtype):
    """cheetah_b200.ParticleBeam from an oracle-style beam dict."""
    import cheetah_b200

    charges = float(beam["num_elementary_charges"])
    mass = float(beam["mass_eV"])
    known = {(-1.0, 510998.95069): "electron", (1.0, 938272089.4300001): "proton"}
    name = known.get((charges, mass))
    if name is not None:
        species = cheetah_b200.Species(name, device=device, dtype=dtype)
    else:
        species = cheetah_b200.Species(
            "custom",
            num_elementary_charges=torch.tensor(charges, device=device, dtype=dtype),
            mass_eV=torch.tensor(mass, device=device, dtype=torch.float64).to(dtype),
        )
    to = lambda t: t.to(device=device, dtype=dtype)  # noqa: E731
    return cheetah_b200.ParticleBeam(
        particles=to(beam["particles"]),
        energy=to(beam["energy"]),
        particle_charges=to(beam["particle_charges"]),
        survival_probabilities=to(beam["survival_probabilities"]),
        s=to(beam["s"]),
        species=species,
    )


def column_scaled_error(actual: torch.Tensor, expected: torch.Tensor) -> float:
    """max |actual - expected| / max|expected| per phase-space column (the SURVEY 7.2 metric)."""
    actual = actual.detach().cpu().to(torch.float64)
    expected = expected.detach().cpu().to(torch.float64)
    scale = expected.abs().amax(dim=-2, keepdim=True).clamp_min(1e-300)
    return float(((actual - expected).abs() / scale)[..., :6].max())


def nonlinear_cases(dtype=torch.float64) -> tuple[dict, dict, int]:
    """(arrays, {case: {"lattice", "beam prefix"}}, row stride) of tests/golden/nonlinear.*"""
    arrays = load_npz("nonlinear.npz")
    with (GOLDEN / "nonlinear.json").open() as f:
        raw = json.load(f)
    cases = {}
    for key, value in raw["lattices"].items():
        if key.startswith("fresh."):
            cases[key] = {
                "lattice": lattice_io._from_json(value["lattice"], dtype),
                "beam": "fresh." + value["beam"],
            }
        elif key.startswith("bmadx."):
            cases[key] = {"lattice": lattice_io._from_json(value, dtype), "beam": "bmadx.incoming"}
        else:
            cases[key] = {"lattice": lattice_io._from_json(value, dtype), "beam": "incoming"}
    return arrays, cases, raw["row_stride"]
