"""Import the UNMODIFIED reference package (``cheetah``) installed under ``oracle/_ref``.

TEST / BASELINE INFRASTRUCTURE ONLY -- nothing in ``cheetah_b200/`` imports this module.
``oracle/build_ref.py`` pip-installs /root/reference into ``oracle/_ref`` in the build container;
the directory travels to the GPU box.  ``oracle/refshim`` provides the inert ``matplotlib``
stand-in the reference imports at import time (matplotlib is not in the image).
"""

from __future__ import annotations

import sys
import warnings
from pathlib import Path

ORACLE = Path(__file__).resolve().parent
_module = None


def available() -> bool:
    return (ORACLE / "_ref" / "cheetah" / "__init__.py").exists()


def load():
    """The reference's ``cheetah`` module (None when ``oracle/_ref`` is absent)."""
    global _module
    if _module is not None:
        return _module
    if not available():
        return None
    for path in (ORACLE / "refshim", ORACLE / "_ref"):
        if str(path) not in sys.path:
            sys.path.insert(0, str(path))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import cheetah  # noqa: PLC0415  (the reference)
    assert Path(cheetah.__file__).resolve().is_relative_to(ORACLE / "_ref"), cheetah.__file__
    _module = cheetah
    return cheetah


def segment(description: list, device=None, dtype=None):
    """``cheetah.Segment`` of a plain-dict lattice description (``oracle/lattice_io.py``)."""
    from . import lattice_io

    cheetah = load()
    segment = cheetah.Segment(elements=lattice_io.build(description, cheetah, device, dtype))
    # defaults the description does not list (tilts, misalignments, ...) are created on the CPU
    return segment if device is None else segment.to(device)


def particle_beam(particles, energy, device=None, dtype=None, particle_charges=None,
                  survival_probabilities=None):
    """``cheetah.ParticleBeam`` of electrons from a (..., N, 7) tensor."""
    import torch

    cheetah = load()
    kwargs = {}
    if particle_charges is not None:
        kwargs["particle_charges"] = particle_charges.to(device=device, dtype=dtype)
    if survival_probabilities is not None:
        kwargs["survival_probabilities"] = survival_probabilities.to(device=device, dtype=dtype)
    return cheetah.ParticleBeam(
        particles=particles.to(device=device, dtype=dtype),
        energy=torch.as_tensor(energy, dtype=dtype, device=device),
        device=device, dtype=dtype, **kwargs,
    )
