"""Lattice elements: the host-side mirror of ``cheetah/accelerator/*.py`` for the hot path.

The classes keep the reference's names, constructor arguments, parameter attribute names
(tensors registered as ``nn.Module`` buffers / parameters), ``is_skippable`` /
``is_active`` rules, ``tracking_method`` hook and error behaviour, so user code and tests
written against ``cheetah`` read the same.  They hold no physics: ``track`` and
``first_order_transfer_map`` lower the element(s) to a lattice program and run the CUDA
kernels (``cheetah_b200/tracking.py``).  There is no CPU path.

Reference for the shared behaviour: cheetah/accelerator/element.py:17-312 (base class,
tracking-method setter with ``PhysicsWarning`` at :239-259), segment.py:27-170 and
:525-574 (container, flattening, skippable-run grouping).
"""

from __future__ import annotations

import itertools
import warnings
from typing import Any, Iterable

import torch
from torch import nn

from .beam import Beam
from .species import Species


class PhysicsWarning(UserWarning):
    """Soft failure of a physics feature (mirror of cheetah/utils/warnings.py)."""


_name_counter = itertools.count()

# Bumped whenever any element attribute is (re)assigned or moved between devices; cached
# lattice lowerings compare against it (the role of cheetah/utils/cache.py:29-40).
_lattice_epoch = 0


def lattice_epoch() -> int:
    return _lattice_epoch


def _bump_epoch() -> None:
    global _lattice_epoch
    _lattice_epoch += 1


def lattice_signature(elements) -> tuple:
    """Identity and per-element edit counter of every (flattened) element: what a cached lowering
    of ``elements`` is valid for.  The global epoch is only the cheap first test; this tells a
    real edit of THIS lattice from attribute traffic on unrelated elements (clones, other
    segments) and sees structural edits (removed, inserted or reordered elements)."""
    from .lowering import flatten

    return tuple((id(e), e.__dict__.get("_epoch", 0)) for e in flatten(elements))


class ElementList(nn.ModuleList):
    """``nn.ModuleList`` whose structural edits (``del seg.elements[1]``, ``insert``, ``append``,
    item assignment ...) bump the lattice epoch, so that cached lowerings are re-validated."""

    def __setitem__(self, idx, module):
        _bump_epoch()
        return super().__setitem__(idx, module)

    def __delitem__(self, idx):
        _bump_epoch()
        return super().__delitem__(idx)

    def __iadd__(self, modules):
        _bump_epoch()
        return super().__iadd__(modules)

    def insert(self, index, module):
        _bump_epoch()
        return super().insert(index, module)

    def append(self, module):
        _bump_epoch()
        return super().append(module)

    def extend(self, modules):
        _bump_epoch()
        return super().extend(modules)

    def pop(self, key):
        _bump_epoch()
        return super().pop(key)


class Element(nn.Module):
    """Base class of all lattice elements."""

    # name -> default value of the tensor parameters a subclass registers
    tensor_fields: dict[str, Any] = {}
    supported_tracking_methods: list[str] = []

    def __init__(
        self,
        name: str | None = None,
        sanitize_name: bool | None = None,
        metadata: dict | None = None,
        device: torch.device | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        self.name = name if name is not None else f"unnamed_element_{next(_name_counter)}"
        self.metadata = metadata if metadata is not None else {}
        if not isinstance(getattr(type(self), "length", None), property):
            self.register_buffer("length", torch.tensor(0.0, device=device, dtype=dtype))
        if not self.supported_tracking_methods:
            self.supported_tracking_methods = [self.__class__.__name__.lower()]
        self._tracking_method = self.supported_tracking_methods[0]

    # ---- parameter registration ---------------------------------------------------------
    def register_buffer_or_parameter(self, name: str, value: torch.Tensor) -> None:
        if isinstance(value, nn.Parameter):
            self.register_parameter(name, value)
        else:
            self.register_buffer(name, value)

    def _register_fields(self, given: dict, factory_kwargs: dict) -> None:
        for field, default in self.tensor_fields.items():
            value = given.get(field)
            if value is None:
                value = torch.tensor(default, **factory_kwargs)
            if field == "length":
                self.length = value
            else:
                self.register_buffer_or_parameter(field, value)

    def __setattr__(self, name: str, value: Any) -> None:
        _bump_epoch()
        object.__setattr__(self, "_epoch", self.__dict__.get("_epoch", 0) + 1)
        super().__setattr__(name, value)

    def _apply(self, fn, *args, **kwargs):
        _bump_epoch()
        object.__setattr__(self, "_epoch", self.__dict__.get("_epoch", 0) + 1)
        return super()._apply(fn, *args, **kwargs)

    # ---- tracking-method hook (element.py:231-259) ---------------------------------------
    @property
    def tracking_method(self) -> str:
        return self._tracking_method

    @tracking_method.setter
    def tracking_method(self, tracking_method: str) -> None:
        if tracking_method in self.supported_tracking_methods:
            self._tracking_method = tracking_method
        else:
            warnings.warn(
                f"Invalid tracking method '{tracking_method}' for element {self.name} of type "
                f"{self.__class__.__name__}, supported methods are "
                f"{self.supported_tracking_methods}. Keeping the previous tracking method "
                f"{self._tracking_method}.",
                PhysicsWarning,
                stacklevel=2,
            )

    @property
    def is_skippable(self) -> bool:
        return True

    # ---- the accelerated path -------------------------------------------------------------
    def first_order_transfer_map(self, energy: torch.Tensor, species: Species) -> torch.Tensor:
        from . import tracking

        return tracking.first_order_transfer_map([self], energy, species)

    def second_order_transfer_map(self, energy: torch.Tensor, species: Species) -> torch.Tensor:
        """Dense ``T_ijk`` with the first-order map in ``T[:, 6, :]`` (element.py:134-147) for the
        elements that support ``tracking_method="second_order"``."""
        from . import tracking

        return tracking.second_order_transfer_map(self, energy, species)

    def transfer_map(self, energy: torch.Tensor, species: Species) -> torch.Tensor:
        """Deprecated alias of ``first_order_transfer_map`` (element.py:67-102)."""
        warnings.warn(
            "The `transfer_map` method is deprecated and will be removed in a future version. "
            "Use `first_order_transfer_map` instead.", DeprecationWarning, stacklevel=2,
        )
        return self.first_order_transfer_map(energy, species)

    @property
    def defining_features(self) -> list[str]:
        """Names of the attributes that define the element (element.py:300-313)."""
        features = ["name", *getattr(self, "tensor_fields", {}), *getattr(self, "plain_fields", {})]
        if len(self.supported_tracking_methods) > 1:
            features.append("tracking_method")
        return features

    @property
    def defining_tensors(self) -> list[str]:
        return [f for f in self.defining_features if isinstance(getattr(self, f), torch.Tensor)]

    def track(self, incoming: Beam) -> Beam:
        from . import tracking

        return tracking.track([self], incoming)

    def forward(self, incoming: Beam) -> Beam:
        return self.track(incoming)

    def merge(self, other: "Element") -> "Element | None":
        """Merged element if ``self`` followed by ``other`` can be expressed as one element of the
        same type, else ``None`` (element.py:349-358; implemented by Drift, Quadrupole, Sextupole,
        Solenoid and Segment)."""
        return None

    def _merged(self, other: "Element", **fields) -> "Element":
        import os

        prefix = os.path.commonprefix([self.name, other.name])  # utils/names.py:17-38
        extras = {}
        if len(self.supported_tracking_methods) > 1:
            extras["tracking_method"] = self.tracking_method
        return self.__class__(
            length=self.length + other.length, **fields, **extras,
            name=prefix if prefix else f"{self.name}_{other.name}", sanitize_name=False,
            metadata={**other.metadata, **self.metadata},
            dtype=self.length.dtype, device=self.length.device,
        )

    def clone(self) -> "Element":
        """Copy of the element that does not share memory with it (element.py:323-336)."""
        import copy

        fields = {key: getattr(self, key).clone() for key in getattr(self, "tensor_fields", {})}
        extras = {key: copy.deepcopy(getattr(self, key))
                  for key in getattr(self, "plain_fields", {})}
        extras = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in extras.items()}
        if self.supported_tracking_methods and hasattr(self, "_tracking_method") and \
                len(self.supported_tracking_methods) > 1:
            extras["tracking_method"] = self.tracking_method
        return self.__class__(**fields, **extras, name=self.name, sanitize_name=False,
                              metadata=copy.deepcopy(self.metadata))

    def split(self, resolution: torch.Tensor) -> list["Element"]:
        """Slices no longer than ``resolution``; elements that cannot be split return
        themselves (element.py:338-347)."""
        return [self]

    def _split_evenly(self, resolution: torch.Tensor, **fields) -> list["Element"]:
        """``ceil(max|length| / resolution)`` equal slices with the other fields unchanged
        (drift.py:160-172, quadrupole.py:261-278, solenoid.py:126-143)."""
        num_splits = int((self.length.abs().max() / resolution).ceil().int())
        extras = {key: getattr(self, key) for key in getattr(self, "plain_fields", {})}
        if self.supported_tracking_methods and hasattr(self, "_tracking_method") and \
                len(self.supported_tracking_methods) > 1:
            extras["tracking_method"] = self.tracking_method
        return [
            self.__class__(
                length=self.length / num_splits, **fields, **extras,
                name=f"{self.name}_split_{i}", sanitize_name=False, metadata=self.metadata,
                dtype=self.length.dtype, device=self.length.device,
            )
            for i in range(num_splits)
        ]

    def __repr__(self) -> str:
        fields = ", ".join(f"{k}={getattr(self, k)!r}" for k in self.tensor_fields)
        return f"{self.__class__.__name__}({fields}, name={self.name!r})"


def _element_init(cls_fields: Iterable[str]):
    """Build the keyword constructor shared by the simple magnet classes."""

    def __init__(self, *args, name=None, sanitize_name=None, metadata=None, device=None,
                 dtype=None, **kwargs):
        fields = list(cls_fields)
        if len(args) > len(fields):
            raise TypeError(f"{type(self).__name__} takes at most {len(fields)} positional arguments")
        given = dict(zip(fields, args))
        tracking_method = kwargs.pop("tracking_method", None)
        extras = {k: kwargs.pop(k) for k in list(kwargs) if k in self.plain_fields}
        for key, value in kwargs.items():
            if key not in fields:
                raise TypeError(f"{type(self).__name__} got an unexpected keyword argument {key!r}")
            given[key] = value
        if device is None:  # defaults follow the tensors that were given
            device = next(
                (v.device for v in given.values() if isinstance(v, torch.Tensor)), None
            )
        Element.__init__(self, name=name, sanitize_name=sanitize_name, metadata=metadata,
                         device=device, dtype=dtype)
        self._register_fields(given, {"device": device, "dtype": dtype})
        for key, default in self.plain_fields.items():
            setattr(self, key, extras.get(key, default))
        if tracking_method is not None:
            self.tracking_method = tracking_method

    return __init__


class _SimpleElement(Element):
    plain_fields: dict[str, Any] = {}

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        if "__init__" not in cls.__dict__:
            cls.__init__ = _element_init(cls.tensor_fields)


class Drift(_SimpleElement):
    """Drift section (cheetah/accelerator/drift.py)."""

    tensor_fields = {"length": 0.0}
    supported_tracking_methods = ["linear", "second_order", "drift_kick_drift"]

    @property
    def is_skippable(self) -> bool:
        return self.tracking_method == "linear"

    def split(self, resolution: torch.Tensor) -> list[Element]:
        return self._split_evenly(resolution)

    def merge(self, other: "Drift") -> "Drift | None":
        if self.tracking_method != other.tracking_method:  # drift.py:175-187
            return None
        return self._merged(other)


class Quadrupole(_SimpleElement):
    """Quadrupole magnet (cheetah/accelerator/quadrupole.py)."""

    tensor_fields = {"length": 0.0, "k1": 0.0, "misalignment": (0.0, 0.0), "tilt": 0.0}
    plain_fields = {"num_steps": 1}
    supported_tracking_methods = ["linear", "second_order", "drift_kick_drift"]

    @property
    def is_skippable(self) -> bool:
        return self.tracking_method == "linear"

    @property
    def is_active(self) -> bool:
        return bool((self.k1 != 0).any())

    def split(self, resolution: torch.Tensor) -> list[Element]:
        return self._split_evenly(resolution, k1=self.k1, misalignment=self.misalignment,
                                  tilt=self.tilt)

    def merge(self, other: "Quadrupole") -> "Quadrupole | None":
        """Length-weighted k1, summed ``num_steps`` (quadrupole.py:280-301)."""
        if not (self.tracking_method == other.tracking_method
                and self.misalignment.equal(other.misalignment) and self.tilt.equal(other.tilt)):
            return None
        return self._merged(
            other,
            k1=(self.k1 * self.length + other.k1 * other.length) / (self.length + other.length),
            misalignment=self.misalignment, tilt=self.tilt,
            num_steps=self.num_steps + other.num_steps,
        )


class Sextupole(_SimpleElement):
    """Sextupole magnet; linear tracking is a drift (cheetah/accelerator/sextupole.py:84-88)."""

    tensor_fields = {"length": 0.0, "k2": 0.0, "misalignment": (0.0, 0.0), "tilt": 0.0}
    supported_tracking_methods = ["second_order", "linear"]

    @property
    def is_skippable(self) -> bool:
        return self.tracking_method == "linear"

    def merge(self, other: "Sextupole") -> "Sextupole | None":
        if not (self.tracking_method == other.tracking_method and self.k2.equal(other.k2)
                and self.misalignment.equal(other.misalignment) and self.tilt.equal(other.tilt)):
            return None  # sextupole.py:133-152
        return self._merged(other, k2=self.k2, misalignment=self.misalignment, tilt=self.tilt)


class Dipole(_SimpleElement):
    """Sector bend with pole-face edges (cheetah/accelerator/dipole.py)."""

    tensor_fields = {
        "length": 0.0, "angle": 0.0, "k1": 0.0, "dipole_e1": 0.0, "dipole_e2": 0.0,
        "tilt": 0.0, "gap": 0.0, "gap_exit": None, "fringe_integral": 0.0,
        "fringe_integral_exit": None,
    }
    plain_fields = {"fringe_at": "both", "fringe_type": "linear_edge"}
    supported_tracking_methods = ["linear", "second_order", "drift_kick_drift"]

    def _register_fields(self, given: dict, factory_kwargs: dict) -> None:
        # exit values default to the entrance ones (dipole.py:111-124)
        given = dict(given)
        for field, default in self.tensor_fields.items():
            if given.get(field) is None and default is not None:
                given[field] = torch.tensor(default, **factory_kwargs)
        if given.get("gap_exit") is None:
            given["gap_exit"] = given["gap"]
        if given.get("fringe_integral_exit") is None:
            given["fringe_integral_exit"] = given["fringe_integral"]
        for field in self.tensor_fields:
            if field == "length":
                self.length = given[field]
            else:
                self.register_buffer_or_parameter(field, given[field])

    @property
    def hx(self) -> torch.Tensor:
        return self.angle / self.length

    @property
    def is_skippable(self) -> bool:
        return self.tracking_method == "linear"

    @property
    def is_active(self) -> bool:
        return bool((self.angle != 0).any())


class RBend(Dipole):
    """Rectangular bend: a Dipole whose edge angles are offset by angle/2 (rbend.py:83-101)."""

    def __init__(self, length, angle=None, k1=None, rbend_e1=None, rbend_e2=None, tilt=None,
                 gap=None, gap_exit=None, fringe_integral=None, fringe_integral_exit=None,
                 fringe_at="both", fringe_type="linear_edge", tracking_method="linear",
                 name=None, sanitize_name=None, metadata=None, device=None, dtype=None):
        if device is None:
            device = length.device
        factory_kwargs = {"device": device, "dtype": dtype}
        angle = angle if angle is not None else torch.tensor(0.0, **factory_kwargs)
        rbend_e1 = rbend_e1 if rbend_e1 is not None else torch.tensor(0.0, **factory_kwargs)
        rbend_e2 = rbend_e2 if rbend_e2 is not None else torch.tensor(0.0, **factory_kwargs)
        Dipole.__init__(
            self, length=length, angle=angle, k1=k1, dipole_e1=rbend_e1 + angle / 2,
            dipole_e2=rbend_e2 + angle / 2, tilt=tilt, gap=gap, gap_exit=gap_exit,
            fringe_integral=fringe_integral, fringe_integral_exit=fringe_integral_exit,
            fringe_at=fringe_at, fringe_type=fringe_type, tracking_method=tracking_method,
            name=name, sanitize_name=sanitize_name, metadata=metadata, **factory_kwargs,
        )

    def clone(self) -> "RBend":
        import copy

        fields = {key: getattr(self, key).clone() for key in self.tensor_fields
                  if key not in ("dipole_e1", "dipole_e2")}
        extras = {key: copy.deepcopy(getattr(self, key)) for key in self.plain_fields}
        return self.__class__(**fields, **extras, rbend_e1=self.rbend_e1, rbend_e2=self.rbend_e2,
                              tracking_method=self.tracking_method, name=self.name,
                              sanitize_name=False, metadata=copy.deepcopy(self.metadata))

    @property
    def rbend_e1(self) -> torch.Tensor:
        return self.dipole_e1 - self.angle / 2

    @rbend_e1.setter
    def rbend_e1(self, value: torch.Tensor) -> None:
        self.dipole_e1 = value + self.angle / 2

    @property
    def rbend_e2(self) -> torch.Tensor:
        return self.dipole_e2 - self.angle / 2

    @rbend_e2.setter
    def rbend_e2(self, value: torch.Tensor) -> None:
        self.dipole_e2 = value + self.angle / 2


class HorizontalCorrector(_SimpleElement):
    tensor_fields = {"length": 0.0, "angle": 0.0}
    supported_tracking_methods = ["linear"]

    @property
    def is_active(self) -> bool:
        return bool((self.angle != 0).any())


class VerticalCorrector(_SimpleElement):
    tensor_fields = {"length": 0.0, "angle": 0.0}
    supported_tracking_methods = ["linear"]

    @property
    def is_active(self) -> bool:
        return bool((self.angle != 0).any())


class CombinedCorrector(_SimpleElement):
    tensor_fields = {"length": 0.0, "horizontal_angle": 0.0, "vertical_angle": 0.0}
    supported_tracking_methods = ["linear"]


class Solenoid(_SimpleElement):
    tensor_fields = {"length": 0.0, "k": 0.0, "misalignment": (0.0, 0.0)}
    supported_tracking_methods = ["linear"]

    @property
    def is_active(self) -> bool:
        return bool((self.k != 0).any())

    def merge(self, other: "Solenoid") -> "Solenoid | None":
        if not self.misalignment.equal(other.misalignment):  # solenoid.py:143-157
            return None
        return self._merged(
            other,
            k=(self.k * self.length + other.k * other.length) / (self.length + other.length),
            misalignment=self.misalignment,
        )

    def split(self, resolution: torch.Tensor) -> list[Element]:
        return self._split_evenly(resolution, k=self.k, misalignment=self.misalignment)


class Undulator(_SimpleElement):
    tensor_fields = {"length": 0.0, "period": 0.0, "kx": 0.0, "ky": 0.0}
    supported_tracking_methods = ["linear"]


class Cavity(_SimpleElement):
    """RF cavity; with voltage == 0 it is a skippable linear map (cavity.py:86-92)."""

    tensor_fields = {"length": 0.0, "voltage": 0.0, "phase": 0.0, "frequency": 0.0}
    plain_fields = {"cavity_type": "standing_wave"}

    @property
    def is_active(self) -> bool:
        return bool((self.voltage != 0).any())

    @property
    def is_skippable(self) -> bool:
        return not self.is_active


class TransverseDeflectingCavity(_SimpleElement):
    """Transverse deflecting RF cavity, tracked per particle with the Bmad-X drift-kick-drift
    map (cheetah/accelerator/transverse_deflecting_cavity.py)."""

    tensor_fields = {
        "length": 0.0, "voltage": 0.0, "phase": 0.0, "frequency": 0.0,
        "misalignment": (0.0, 0.0), "tilt": 0.0,
    }
    plain_fields = {"num_steps": 1}
    supported_tracking_methods = ["drift_kick_drift"]

    @property
    def is_active(self) -> bool:
        return bool((self.voltage != 0).any())

    @property
    def is_skippable(self) -> bool:
        return False


class Marker(_SimpleElement):
    tensor_fields = {}


class BPM(_SimpleElement):
    """Beam position monitor (cheetah/accelerator/bpm.py): when active it records
    ``reading = (mu_x, mu_y) - misalignment`` of the passing beam."""

    tensor_fields = {"misalignment": (0.0, 0.0)}
    plain_fields = {"is_active": False}

    @property
    def is_skippable(self) -> bool:
        return not self.is_active

    @property
    def reading(self) -> torch.Tensor:
        reading = self.__dict__.get("_reading")
        if reading is None:  # bpm.py:58-63
            reading = torch.full_like(self.misalignment, float("nan"))
        return reading

    @reading.setter
    def reading(self, value: torch.Tensor) -> None:
        object.__setattr__(self, "_reading", value)

    def track(self, incoming: Beam) -> Beam:
        from . import diagnostics

        return diagnostics.track_bpm(self, incoming)


class Screen(_SimpleElement):
    """Diagnostic screen (cheetah/accelerator/screen.py): when active it remembers the passing
    beam and renders ``reading`` (height, width) lazily with ``ch_screen_image``
    (cloud-in-cell, histogram), ``ch_screen_kde`` or, for a ParameterBeam, ``ch_screen_gaussian``."""

    tensor_fields = {"pixel_size": (1e-3, 1e-3), "misalignment": (0.0, 0.0)}
    plain_fields = {
        "resolution": (1024, 1024), "binning": 1, "method": "cloud-in-cell",
        "kde_bandwidth": None, "is_blocking": False, "is_active": False,
    }

    def __setattr__(self, name: str, value: Any) -> None:
        if name == "resolution":
            assert isinstance(value, (tuple, list)) and len(value) == 2, (
                "Invalid resolution. Must be a tuple of 2 integers."
            )
        if name == "method":
            assert value in ["histogram", "kde", "cloud-in-cell"], (
                f"Invalid method {value}. Must be 'histogram', 'kde', or 'cloud-in-cell'."
            )
        if name in ("resolution", "binning", "method", "pixel_size", "misalignment",
                    "kde_bandwidth"):
            object.__setattr__(self, "_cached_reading", None)
        super().__setattr__(name, value)

    @property
    def is_skippable(self) -> bool:
        return not self.is_active

    @property
    def effective_resolution(self) -> tuple[int, int]:
        return (self.resolution[0] // self.binning, self.resolution[1] // self.binning)

    @property
    def effective_pixel_size(self) -> torch.Tensor:
        return self.pixel_size * self.binning

    @property
    def extent(self) -> torch.Tensor:
        return torch.stack([
            -self.resolution[0] * self.pixel_size[0] / 2, self.resolution[0] * self.pixel_size[0] / 2,
            -self.resolution[1] * self.pixel_size[1] / 2, self.resolution[1] * self.pixel_size[1] / 2,
        ])

    @property
    def pixel_bin_edges(self) -> tuple[torch.Tensor, torch.Tensor]:
        return tuple(
            torch.linspace(
                -self.resolution[i] * self.pixel_size[i] / 2,
                self.resolution[i] * self.pixel_size[i] / 2,
                int(self.effective_resolution[i]) + 1,
                device=self.pixel_size.device, dtype=self.pixel_size.dtype,
            )
            for i in range(2)
        )

    @property
    def pixel_bin_centers(self) -> tuple[torch.Tensor, torch.Tensor]:
        edges = self.pixel_bin_edges
        return ((edges[0][1:] + edges[0][:-1]) / 2, (edges[1][1:] + edges[1][:-1]) / 2)

    def get_read_beam(self):
        return self.__dict__.get("_read_beam")

    def set_read_beam(self, value) -> None:
        object.__setattr__(self, "_read_beam", value)
        object.__setattr__(self, "_cached_reading", None)

    @property
    def reading(self) -> torch.Tensor:
        """Image of the screen, ``(..., height, width)`` (screen.py:241-344)."""
        cached = self.__dict__.get("_cached_reading")
        if cached is not None:
            return cached
        read_beam = self.get_read_beam()
        if read_beam is None:
            image = self.misalignment.new_zeros(
                (int(self.effective_resolution[1]), int(self.effective_resolution[0]))
            )
        else:
            from . import diagnostics

            image = diagnostics.screen_image(self, read_beam)
        object.__setattr__(self, "_cached_reading", image)
        return image

    def track(self, incoming: Beam) -> Beam:
        from . import diagnostics

        return diagnostics.track_screen(self, incoming)


class Aperture(_SimpleElement):
    """Physical aperture (cheetah/accelerator/aperture.py)."""

    tensor_fields = {"x_max": float("inf"), "y_max": float("inf")}
    plain_fields = {"shape": "rectangular", "is_active": True}

    @property
    def is_skippable(self) -> bool:
        return not self.is_active


class CustomTransferMap(Element):
    """Element defined by a user-supplied 7x7 map (custom_transfer_map.py:32-58)."""

    supported_tracking_methods = ["linear"]
    tensor_fields = {"predefined_transfer_map": None}

    def __init__(self, predefined_transfer_map: torch.Tensor, length: torch.Tensor | None = None,
                 name=None, sanitize_name=None, metadata=None, device=None, dtype=None) -> None:
        super().__init__(name=name, sanitize_name=sanitize_name, metadata=metadata,
                         device=device, dtype=dtype)
        if length is not None:
            self.length = length
        assert predefined_transfer_map.shape[-2:] == (7, 7)
        assert (predefined_transfer_map[..., -1, :-1] == 0.0).all() and (
            predefined_transfer_map[..., -1, -1] == 1.0
        ).all(), "The seventh row of the transfer map must be [0, 0, 0, 0, 0, 0, 1]."
        self.register_buffer_or_parameter("predefined_transfer_map", predefined_transfer_map)

    @property
    def defining_features(self) -> list[str]:
        return ["name", "predefined_transfer_map", "length"]

    def clone(self) -> "CustomTransferMap":
        import copy

        return self.__class__(self.predefined_transfer_map.clone(), length=self.length.clone(),
                              name=self.name, sanitize_name=False,
                              metadata=copy.deepcopy(self.metadata))

    @classmethod
    def from_merging_elements(cls, elements: list, incoming_beam: Beam) -> "CustomTransferMap":
        """One map for a run of skippable elements (custom_transfer_map.py:60-109): the product
        is formed on the device in fp64 by ``ch_compose_maps`` and rounded once."""
        assert all(element.is_skippable for element in elements), (
            "Combining the elements in a Segment that is not skippable will result in"
            " incorrect tracking results."
        )
        from . import tracking

        tm = tracking.first_order_transfer_map(
            list(elements), incoming_beam.energy, incoming_beam.species
        )
        length = sum(element.length for element in elements)
        name = "combined_" + "_".join(element.name for element in elements)
        return cls(tm, length=length, name=name, sanitize_name=False)


class SpaceChargeKick(_SimpleElement):
    """IGF space-charge kick (cheetah/accelerator/space_charge_kick.py)."""

    tensor_fields = {
        "effect_length": 0.0, "grid_extent_x": 3.0, "grid_extent_y": 3.0, "grid_extent_tau": 3.0,
    }
    plain_fields = {"grid_shape": (32, 32, 32)}

    @property
    def is_skippable(self) -> bool:
        return False


class Superimposed(Element):
    """One zero-length element placed at the centre of another
    (cheetah/accelerator/superimposed.py): lowered as [first half, element, second half]."""

    def __init__(self, base_element: Element, superimposed_element: Element,
                 name: str | None = None, sanitize_name: bool | None = None,
                 metadata: dict | None = None, device=None, dtype=None) -> None:
        super().__init__(name=name, sanitize_name=sanitize_name, metadata=metadata,
                         device=device, dtype=dtype)
        assert bool((superimposed_element.length == 0.0).all()), (
            "The superimposed element must have zero length."
        )
        self.base_element = base_element
        self.superimposed_element = superimposed_element
        halves = base_element.split(base_element.length / 2.0)
        assert len(halves) == 2, f"{type(base_element).__name__} cannot be split in two"
        self._segment = Segment(
            elements=[halves[0], superimposed_element, halves[1]], name=f"{self.name}_segment",
            sanitize_name=sanitize_name,
        )

    def flattened(self) -> "Segment":
        return self._segment.flattened()

    @property
    def defining_features(self) -> list[str]:
        return ["name", "base_element", "superimposed_element"]

    def clone(self) -> "Superimposed":
        import copy

        return self.__class__(self.base_element.clone(), self.superimposed_element.clone(),
                              name=self.name, sanitize_name=False,
                              metadata=copy.deepcopy(self.metadata))

    @property
    def is_skippable(self) -> bool:
        return self._segment.is_skippable

    @property
    def length(self) -> torch.Tensor:
        return self._segment.length

    def first_order_transfer_map(self, energy: torch.Tensor, species: Species) -> torch.Tensor:
        return self._segment.first_order_transfer_map(energy, species)

    def track(self, incoming: Beam) -> Beam:
        return self._segment.track(incoming)


class Segment(Element):
    """Ordered list of elements (cheetah/accelerator/segment.py)."""

    def __init__(self, elements: list[Element], name: str | None = None,
                 sanitize_name: bool | None = None, metadata: dict | None = None) -> None:
        super().__init__(name=name, sanitize_name=sanitize_name, metadata=metadata)
        self.elements = ElementList(elements)
        for element in elements:  # `segment.<element name>` access (segment.py:60-70)
            if element.name.isidentifier() and not hasattr(self, element.name):
                object.__setattr__(self, "_alias_" + element.name, element)
        self._plan_cache = None

    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            alias = self.__dict__.get("_alias_" + name)
            if alias is not None:
                return alias
            raise

    @property
    def length(self) -> torch.Tensor:
        total = None
        for element in self.elements:
            total = element.length if total is None else total + element.length
        return total

    @property
    def is_skippable(self) -> bool:
        return all(element.is_skippable for element in self.elements)

    def flattened(self) -> "Segment":
        """Flat copy of the element list (nested segments expanded; segment.py:143-157)."""
        flat = []
        for element in self.elements:
            if isinstance(element, Segment):
                flat.extend(element.flattened().elements)
            else:
                flat.append(element)
        return Segment(elements=flat, name=self.name, sanitize_name=False)

    @classmethod
    def from_lattice_json(cls, filepath, name: str | None = None, device=None, dtype=None
                          ) -> "Segment":
        """Load the reference's LatticeJSON format (segment.py:370-384)."""
        from . import latticejson

        segment = latticejson.load_segment(filepath, device=device, dtype=dtype)
        if name is not None:
            segment.name = name
        return segment

    def to_lattice_json(self, filepath, title: str | None = None,
                        info: str = "This is a placeholder lattice description") -> None:
        """Save in the reference's LatticeJSON format (segment.py:386-396)."""
        from . import latticejson

        latticejson.save_segment(self, filepath, title=title, info=info)

    # ---- container helpers (segment.py:73-229, :576-656) ------------------------------------
    @property
    def element_names(self) -> list[str]:
        return [element.name for element in self.elements]

    def element_index(self, element_name: str) -> int:
        try:
            return self.element_names.index(element_name)
        except ValueError:
            raise ValueError(f"Element '{element_name}' not found in segment.")

    def subcell(self, start: str | None = None, end: str | None = None,
                include_start: bool = True, include_end: bool = True) -> "Segment":
        """Elements from ``start`` to ``end`` (segment.py:94-141)."""
        names = self.element_names
        if start is not None and start not in names:
            raise ValueError(f"Element {start} is not part of the segment.")
        if end is not None and end not in names:
            raise ValueError(f"Element {end} is not part of the segment.")
        subcell = []
        is_in_subcell = start is None
        for element in self.elements:
            if element.name == start:
                is_in_subcell = True
                if include_start:
                    subcell.append(element)
                continue
            if element.name == end:
                if include_end and is_in_subcell:
                    subcell.append(element)
                break
            if is_in_subcell:
                subcell.append(element)
        return self.__class__(subcell)

    def reversed(self) -> "Segment":
        elements = [e.reversed() if isinstance(e, Segment) else e for e in self.elements][::-1]
        return self.__class__(elements=elements, name=f"{self.name}_reversed", sanitize_name=False)

    def partition_at(self, element_name: str, mode: str = "both") -> tuple:
        """Split around a named element (segment.py:599-629)."""
        index = self.element_index(element_name)
        elements = list(self.elements)
        pre = self.__class__(elements[: index + 1] if mode == "after" else elements[:index])
        post = self.__class__(elements[index:] if mode == "before" else elements[index + 1 :])
        return (pre, elements[index], post) if mode == "both" else (pre, post)

    def transfer_maps_merged(self, incoming_beam: Beam, except_for: list[str] | None = None
                             ) -> "Segment":
        """Runs of skippable elements replaced by one ``CustomTransferMap`` each
        (segment.py:179-229).  NOTE: ``Segment.track`` already merges every skippable run on the
        device for each call (``ch_compose_maps``), so this only saves the composition launch."""
        except_for = except_for or []
        merged, run = [], []
        beam = incoming_beam

        def flush() -> None:
            nonlocal beam, run
            if len(run) == 1:
                merged.append(run[0])
                beam = run[0].track(beam)
            elif len(run) > 1:
                merged.append(CustomTransferMap.from_merging_elements(run, incoming_beam=beam))
                beam = merged[-1].track(beam)
            run = []

        for element in self.elements:
            if element.is_skippable and element.name not in except_for:
                run.append(element)
            else:
                flush()
                merged.append(element)
                beam = element.track(beam)
        if run:
            merged.append(CustomTransferMap.from_merging_elements(run, incoming_beam=beam))
        return self.__class__(elements=merged, name=self.name, sanitize_name=False)

    def split(self, resolution: torch.Tensor) -> list[Element]:
        return [part for element in self.elements for part in element.split(resolution)]

    def clone(self) -> "Segment":
        """Deep copy: every element cloned (segment.py, element.py:323-336)."""
        import copy

        return self.__class__([element.clone() for element in self.elements], name=self.name,
                              sanitize_name=False, metadata=copy.deepcopy(self.metadata))

    # ---- the reference's lattice simplifications (segment.py:231-330).  The composer already
    # folds markers, inactive monitors and zero-strength magnets into one map, so these do not
    # change the cost of a track() here; they are kept because user code calls them.
    def merge(self, other: "Segment") -> "Segment":
        import os

        prefix = os.path.commonprefix([self.name, other.name])  # segment.py:591-597
        return self.__class__(elements=list(self.elements) + list(other.elements),
                              name=prefix if prefix else f"{self.name}_{other.name}",
                              sanitize_name=False, metadata={**other.metadata, **self.metadata})

    def with_consecutive_elements_merged(self, except_for: list[str] | None = None) -> "Segment":
        """Consecutive mergeable elements of the same type combined into one
        (segment.py:326-367).  The fused composer makes this unnecessary for speed here."""
        import copy

        except_for = except_for or []
        merged, current = [], self.elements[0]
        for following in list(self.elements)[1:]:
            if current.name not in except_for:
                if type(current) is Segment:
                    current = current.with_consecutive_elements_merged(except_for=except_for)
                elif type(current) is type(following) and following.name not in except_for:
                    combined = current.merge(following)
                    if combined is not None:
                        current = combined
                        continue
            merged.append(current)
            current = following
        merged.append(current)
        return self.__class__(elements=merged, name=self.name, sanitize_name=False,
                              metadata=copy.deepcopy(self.metadata))

    def _filtered(self, keep) -> "Segment":
        return self.__class__(elements=[e for e in self.elements if keep(e)], name=self.name,
                              sanitize_name=False)

    def without_inactive_markers(self, except_for: list[str] | None = None) -> "Segment":
        except_for = except_for or []
        return self._filtered(lambda e: not isinstance(e, Marker) or e.name in except_for)

    def without_inactive_zero_length_elements(self, except_for: list[str] | None = None
                                              ) -> "Segment":
        except_for = except_for or []
        return self._filtered(
            lambda e: bool((e.length != 0.0).any()) or bool(getattr(e, "is_active", False))
            or e.name in except_for
        )

    def inactive_elements_as_drifts(self, except_for: list[str] | None = None) -> "Segment":
        except_for = except_for or []

        def converted(element):
            if bool(getattr(element, "is_active", False)) or bool((element.length == 0.0).all()) \
                    or element.name in except_for:
                return element
            return Drift(element.length, name=element.name, sanitize_name=False)

        return self.__class__(elements=[converted(e) for e in self.elements],
                              name=f"{self.name}_inactive_as_drifts", sanitize_name=False)

    def set_attrs_on_every_element(self, filter_type=None, is_recursive: bool = True,
                                   **kwargs) -> None:
        """Set attributes on every element (of ``filter_type``), descending into nested
        segments (segment.py:605-629)."""
        for element in self.elements:
            if filter_type is None or isinstance(element, filter_type):
                for key, value in kwargs.items():
                    setattr(element, key, value)
            elif is_recursive and isinstance(element, Segment):
                element.set_attrs_on_every_element(filter_type, is_recursive=True, **kwargs)

    def get_beam_attrs_along_segment(self, attr_names, incoming: Beam, resolution=None):
        """Beam attributes at the end of every element (or slice), stacked along a new
        dimension in front of the attribute's own ones (segment.py:658-701)."""
        names = attr_names if isinstance(attr_names, tuple) else (attr_names,)
        inner = {"particles": 2, "particle_charges": 1, "survival_probabilities": 1, "x": 1,
                 "px": 1, "y": 1, "py": 1, "tau": 1, "p": 1, "mu": 1, "cov": 2, "energies": 1,
                 "momenta": 1}
        beams = list(self.beam_along_segment_generator(incoming, resolution=resolution))
        results = tuple(
            torch.stack(torch.broadcast_tensors(*[getattr(beam, name) for beam in beams]),
                        dim=-(inner.get(name, 0) + 1))
            for name in names
        )
        return results if isinstance(attr_names, tuple) else results[0]

    def beam_along_segment_generator(self, incoming: Beam, resolution=None):
        """Beams at the end of every element, or of every slice no longer than ``resolution``
        (segment.py:631-656)."""
        if resolution is not None:
            yield from self.__class__(
                elements=self.split(torch.as_tensor(resolution)), name=f"{self.name}_split"
            ).beam_along_segment_generator(incoming)
            return
        yield incoming
        for element in self.elements:
            outgoing = element.track(incoming)
            yield outgoing
            incoming = outgoing

    def first_order_transfer_map(self, energy: torch.Tensor, species: Species) -> torch.Tensor:
        if not self.is_skippable:
            return None  # segment.py:542-543
        from . import tracking

        return tracking.first_order_transfer_map(list(self.elements), energy, species)

    def track(self, incoming: Beam) -> Beam:
        from . import tracking

        return tracking.track(list(self.elements), incoming, cache_owner=self)

    def track_moments(self, incoming: Beam, keep_particles: bool = False,
                      covariance: bool = False):
        """Outgoing-beam moments from the fused kernel epilogue (tracking.track_moments);
        ``covariance=True`` adds the full 6x6 covariance matrix (``BeamMoments.cov``)."""
        from . import tracking

        return tracking.track_moments(list(self.elements), incoming, self, keep_particles,
                                      covariance)

    def __repr__(self) -> str:
        return f"Segment(elements={list(self.elements)!r}, name={self.name!r})"
