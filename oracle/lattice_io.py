"""Plain-dict lattice descriptions shared by the oracle, the golden generator and tests.

TEST INFRASTRUCTURE ONLY (see ``oracle/track_oracle.py``).  ``describe`` turns an element
object -- the reference's ``cheetah.Element`` or the product's ``cheetah_b200.Element``,
both expose the same attribute names -- into ``{"type": ..., "name": ..., <params>}``;
``dump``/``load`` persist a list of such dicts as JSON (tensors as nested lists).
"""

from __future__ import annotations

import json

import torch

# attribute names read per element type (reference: each class's `defining_features`)
_TENSOR_FIELDS = {
    "Drift": ["length"],
    "Quadrupole": ["length", "k1", "misalignment", "tilt"],
    "Dipole": [
        "length", "angle", "k1", "dipole_e1", "dipole_e2", "tilt", "gap", "gap_exit",
        "fringe_integral", "fringe_integral_exit",
    ],
    "RBend": [
        "length", "angle", "k1", "rbend_e1", "rbend_e2", "tilt", "gap", "gap_exit",
        "fringe_integral", "fringe_integral_exit",
    ],
    "HorizontalCorrector": ["length", "angle"],
    "VerticalCorrector": ["length", "angle"],
    "CombinedCorrector": ["length", "horizontal_angle", "vertical_angle"],
    "Solenoid": ["length", "k", "misalignment"],
    "Undulator": ["length", "period", "kx", "ky"],
    "Cavity": ["length", "voltage", "phase", "frequency"],
    "Marker": [],
    "BPM": [],
    "Screen": [],
    "Aperture": ["x_max", "y_max"],
    "Sextupole": ["length", "k2", "misalignment", "tilt"],
    "SpaceChargeKick": [
        "effect_length", "grid_extent_x", "grid_extent_y", "grid_extent_tau",
    ],
    "CustomTransferMap": ["length", "predefined_transfer_map"],
    "TransverseDeflectingCavity": [
        "length", "voltage", "phase", "frequency", "misalignment", "tilt",
    ],
}
_PLAIN_FIELDS = {
    "Drift": ["tracking_method"],
    "Quadrupole": ["tracking_method", "num_steps"],
    "Dipole": ["tracking_method", "fringe_at"],
    "RBend": ["tracking_method", "fringe_at"],
    "TransverseDeflectingCavity": ["num_steps"],
    "Sextupole": ["tracking_method"],
    "Cavity": ["cavity_type"],
    "BPM": ["is_active"],
    "Screen": ["is_active"],
    "Aperture": ["shape", "is_active"],
    "SpaceChargeKick": ["grid_shape"],
}


def describe(element) -> dict:
    kind = type(element).__name__
    if kind == "Segment":
        return {
            "type": "Segment",
            "name": element.name,
            "elements": [describe(sub) for sub in element.elements],
        }
    if kind == "Superimposed":
        return describe(element._segment)
    if kind not in _TENSOR_FIELDS:
        raise NotImplementedError(f"no description for element type {kind}")
    out = {"type": kind, "name": element.name}
    for field in _TENSOR_FIELDS[kind]:
        out[field] = getattr(element, field).detach().cpu().clone()
    for field in _PLAIN_FIELDS.get(kind, []):
        value = getattr(element, field)
        out[field] = list(value) if isinstance(value, tuple) else value
    return out


def cast(description, dtype: torch.dtype):
    """Deep-copy a description with every tensor cast to ``dtype``."""
    if isinstance(description, list):
        return [cast(d, dtype) for d in description]
    out = {}
    for key, value in description.items():
        if key == "elements":
            out[key] = cast(value, dtype)
        elif isinstance(value, torch.Tensor):
            out[key] = value.to(dtype)
        else:
            out[key] = value
    return out


def _to_json(description):
    if isinstance(description, list):
        return [_to_json(d) for d in description]
    out = {}
    for key, value in description.items():
        if key == "elements":
            out[key] = _to_json(value)
        elif isinstance(value, torch.Tensor):
            out[key] = {"__tensor__": value.to(torch.float64).tolist()}
        else:
            out[key] = value
    return out


def _from_json(obj, dtype):
    if isinstance(obj, list):
        return [_from_json(o, dtype) for o in obj]
    out = {}
    for key, value in obj.items():
        if key == "elements":
            out[key] = _from_json(value, dtype)
        elif isinstance(value, dict) and "__tensor__" in value:
            out[key] = torch.tensor(value["__tensor__"], dtype=dtype)
        else:
            out[key] = value
    return out


def dump(description: list, path) -> None:
    with open(path, "w") as f:
        json.dump(_to_json(description), f, indent=0, separators=(",", ":"))


def load(path, dtype: torch.dtype = torch.float32) -> list:
    """Load a lattice; values were stored as float64 and are rounded to ``dtype``."""
    with open(path) as f:
        return _from_json(json.load(f), dtype)


def build(description: list, namespace, device=None, dtype=None) -> list:
    """Instantiate element objects from a description using ``namespace``'s classes.

    ``namespace`` is a module exposing the reference-compatible constructors
    (``cheetah`` or ``cheetah_b200``).
    """
    elements = []
    for d in description:
        kind = d["type"]
        if kind == "Segment":
            elements.append(
                namespace.Segment(
                    elements=build(d["elements"], namespace, device, dtype), name=d["name"]
                )
            )
            continue
        kwargs = {}
        for key, value in d.items():
            if key == "type":
                continue
            if isinstance(value, torch.Tensor):
                value = value.to(device=device, dtype=dtype or value.dtype)
            elif key == "grid_shape":
                value = tuple(value)
            kwargs[key] = value
        kwargs["sanitize_name"] = False
        elements.append(getattr(namespace, kind)(**kwargs))
    return elements
