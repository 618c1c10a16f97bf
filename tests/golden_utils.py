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
