"""Reader for the reference's LatticeJSON files (cheetah/latticejson.py:156-260):

    {"version": ..., "title": ..., "root": "<lattice name>",
     "elements": {"<name>": ["<ClassName>", {<parameters>}], ...},
     "lattices": {"<lattice name>": ["<element or lattice name>", ...], ...}}

Numbers and number lists become tensors on the requested device / dtype; strings, booleans,
integers, dicts (``metadata``) and lists of those are passed through; a parameter whose value is
the name of another element is parsed recursively (``Superimposed``).  Only element types this
package mirrors can be loaded (anything else raises ``AttributeError`` naming the class).
"""

from __future__ import annotations

import json

import torch


def _feature(value, device, dtype):
    if value is None or isinstance(value, (str, bool, int, dict)):
        return value
    if isinstance(value, (tuple, list)) and all(isinstance(v, (str, bool, int)) for v in value):
        return value
    return torch.tensor(value, device=device, dtype=dtype)


def _element(name: str, lattice: dict, device, dtype):
    import cheetah_b200 as cb

    class_name, params = lattice["elements"][name]
    cls = getattr(cb, class_name)
    converted = {
        key: (
            _element(value, lattice, device, dtype)
            if isinstance(value, str) and value in lattice["elements"]
            else _feature(value, device, dtype)
        )
        for key, value in params.items()
    }
    for key in ("resolution", "grid_shape"):
        if key in converted and isinstance(converted[key], list):
            converted[key] = tuple(converted[key])
    return cls(name=name, **converted)


def _segment(name: str, lattice: dict, device, dtype):
    import cheetah_b200 as cb

    elements = [
        _segment(child, lattice, device, dtype)
        if child in lattice["lattices"] else _element(child, lattice, device, dtype)
        for child in lattice["lattices"][name]
    ]
    return cb.Segment(elements=elements, name=name)


def load_segment(filepath, device=None, dtype=None):
    """``Segment.from_lattice_json`` (segment.py:370-384 -> latticejson.py:241-260)."""
    dtype = dtype if dtype is not None else torch.get_default_dtype()
    with open(filepath) as f:
        lattice = json.load(f)
    return _segment(lattice["root"], lattice, device, dtype)
