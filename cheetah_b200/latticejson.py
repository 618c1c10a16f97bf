"""Reader and writer for the reference's LatticeJSON files (cheetah/latticejson.py:9-260):

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


def _plain(value):
    return value.tolist() if isinstance(value, torch.Tensor) else value


def _describe(element, elements: dict) -> tuple[str, dict]:
    """Class name and JSON-ready parameters of one element; elements held as parameters
    (``Superimposed``) are described recursively and referenced by name
    (latticejson.py:26-60)."""
    from .elements import Element

    params = {}
    for feature in element.defining_features:
        if feature == "name":
            continue
        value = getattr(element, feature)
        if isinstance(value, Element):
            elements[value.name] = list(_describe(value, elements))
            params[feature] = value.name
        elif isinstance(value, tuple):
            params[feature] = list(value)
        else:
            params[feature] = _plain(value)
    params["metadata"] = element.metadata
    return element.__class__.__name__, params


def _describe_segment(segment, elements: dict, lattices: dict) -> None:
    from .elements import Segment

    cell = []
    for element in segment.elements:
        if isinstance(element, Segment):
            _describe_segment(element, elements, lattices)
        else:
            elements[element.name] = list(_describe(element, elements))
        cell.append(element.name)
    lattices[segment.name] = cell


def save_segment(segment, filepath, title: str | None = None,
                 info: str = "This is a placeholder lattice description") -> None:
    """``Segment.to_lattice_json`` (segment.py:386-396 -> latticejson.py:96-139): the same
    document structure and keys as the reference writes, so either package reads the file."""
    elements, lattices = {}, {}
    _describe_segment(segment, elements, lattices)
    document = {
        "version": "cheetah-0.8",
        "title": title if title is not None else (segment.name or "Unnamed Lattice"),
        "info": info,
        "root": segment.name if segment.name is not None else "cell",
        "elements": elements,
        "lattices": lattices,
    }
    with open(filepath, "w") as f:
        json.dump(document, f, indent=1)
