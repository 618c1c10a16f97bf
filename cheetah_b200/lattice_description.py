"""Plain-dict lattice descriptions -> element objects.

A description is a list of ``{"type": "Quadrupole", "name": ..., "length": tensor, "k1": tensor,
...}`` dicts (nested ``{"type": "Segment", "elements": [...]}`` allowed) -- the form in which
``tests/golden/ares_lattice.json`` stores the ARES lattice (converted once from the reference's
``docs/examples/ARESlatticeStage3v1_9.json``) and in which ``workloads.py`` builds the benchmark
lattices.  ``load`` reads the JSON form (tensors as ``{"__tensor__": nested lists}``, stored in
float64), ``build`` instantiates ``cheetah_b200`` elements on a device.
"""

from __future__ import annotations

import json

import torch


def _decode(obj, dtype):
    if isinstance(obj, list):
        return [_decode(item, dtype) for item in obj]
    out = {}
    for key, value in obj.items():
        if key == "elements":
            out[key] = _decode(value, dtype)
        elif isinstance(value, dict) and "__tensor__" in value:
            out[key] = torch.tensor(value["__tensor__"], dtype=dtype)
        else:
            out[key] = value
    return out


def load(path, dtype: torch.dtype = torch.float32) -> list:
    """Read a JSON lattice description; values are rounded from float64 to ``dtype``."""
    with open(path) as f:
        return _decode(json.load(f), dtype)


def build(description: list, device=None, dtype=None) -> list:
    """Element objects of this package for ``description`` (tensors moved to ``device``)."""
    import cheetah_b200 as cb

    elements = []
    for entry in description:
        kind = entry["type"]
        if kind == "Segment":
            elements.append(
                cb.Segment(elements=build(entry["elements"], device, dtype), name=entry["name"])
            )
            continue
        kwargs = {}
        for key, value in entry.items():
            if key == "type":
                continue
            if isinstance(value, torch.Tensor):
                value = value.to(device=device, dtype=dtype or value.dtype)
            elif key in ("grid_shape", "resolution"):
                value = tuple(value)
            kwargs[key] = value
        kwargs["sanitize_name"] = False
        elements.append(getattr(cb, kind)(**kwargs))
    return elements
