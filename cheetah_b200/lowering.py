"""Lower a list of element objects to the lattice program consumed by the CUDA library.

This replaces the per-call Python walk of ``Segment.track``
(cheetah/accelerator/segment.py:545-574): consecutive skippable elements and active
apertures form *linear sections* (one ``ch_compose_maps`` + one ``ch_apply_maps`` launch
each); non-linear elements (``SpaceChargeKick``) are *barriers* between sections.

Elements are duck-typed on their class name and the reference's attribute names, so the
same code lowers ``cheetah_b200`` elements and the reference's own ``cheetah`` elements
(INTEGRATION.md).  Parameter tensors are NOT copied: the program stores their device
pointers, so in-place updates of magnet settings are picked up without re-lowering.
"""

from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field

import torch

from . import _capi


@dataclass
class SlotSpec:
    tensor: torch.Tensor
    inner_size: int = 1  # trailing (non-vector) elements per setting
    inner_offset: int = 0
    is_length: bool = False

    @property
    def vector_shape(self) -> tuple:
        inner_dims = 0 if self.inner_size == 1 else (2 if self.inner_size == 49 else 1)
        return tuple(self.tensor.shape[: self.tensor.dim() - inner_dims])


@dataclass
class Op:
    opcode: int
    flags: int
    slots: list
    element: object


@dataclass
class LinearSection:
    op_begin: int
    op_end: int
    n_apertures: int = 0
    elliptical_mask: int = 0
    lattice_shape: tuple = ()          # broadcast vector shape of all parameters
    length_shape: tuple = ()           # ... of the length parameters only
    survival_shape: tuple = ()         # ... of everything up to and including the last aperture
    map_shape: tuple = ()              # ... of the parameters of the non-aperture ops only
    n_map_ops: int = 0                 # non-aperture ops (identity elements included)
    maps_before_last_aperture: bool = False
    has_maps: bool = False             # any non-identity op
    apertures: list = field(default_factory=list)
    cavity: object = None              # active Cavity closing the section (element, gain flag)


@dataclass
class NonlinearRun:
    """Consecutive drift_kick_drift / second_order elements (identity ops in between are
    no-ops): one ``ch_nonlinear_constants`` + one ``ch_track_nonlinear`` launch."""

    op_begin: int
    op_end: int
    lattice_shape: tuple = ()
    length_shape: tuple = ()
    survival_shape: tuple = ()
    methods: tuple = ()                # tracking methods present (for error messages)


@dataclass
class Barrier:
    element: object
    kind: str  # "space_charge", "bpm", "screen" or "unsupported"


class LatticeProgram:
    """Owns the device-side program (``ch_program``) and keeps its tensors alive."""

    def __init__(self, ops: list, stages: list, device: torch.device, watched: list) -> None:
        self.ops = ops
        self.stages = stages
        self.device = device
        self.watched = watched  # [(tensor, version)] whose VALUES shaped the lowering
        self.handle = None  # created on first use, so lowering itself needs no GPU
        self._keepalive = []

    @property
    def native(self):
        """The ``ch_program*`` handle (uploads the tables on first access)."""
        if self.handle is None:
            self._upload()
        return self.handle

    def _upload(self) -> None:
        lib = _capi.lib()
        n_ops = len(self.ops)
        opcodes = (ctypes.c_int32 * max(n_ops, 1))()
        flags = (ctypes.c_int32 * max(n_ops, 1))()
        slot_begin = (ctypes.c_int32 * (n_ops + 1))()
        ptrs, strides, dtypes = [], [], []
        for i, op in enumerate(self.ops):
            opcodes[i] = op.opcode
            flags[i] = op.flags
            slot_begin[i] = len(ptrs)
            for tensor, stride, offset in op.resolved:
                self._keepalive.append(tensor)
                ptrs.append(tensor.data_ptr() + offset * tensor.element_size())
                strides.append(stride)
                dtypes.append(_capi.dtype_code(tensor.dtype))
        slot_begin[n_ops] = len(ptrs)
        n_slots = len(ptrs)
        c_ptrs = (ctypes.c_void_p * max(n_slots, 1))(*ptrs)
        c_strides = (ctypes.c_int64 * max(n_slots, 1))(*strides)
        c_dtypes = (ctypes.c_int32 * max(n_slots, 1))(*dtypes)
        handle = ctypes.c_void_p()
        with _capi.device_guard(self.device):
            _capi.check(
                lib.ch_program_create(
                    opcodes, flags, slot_begin, n_ops, c_ptrs, c_strides, c_dtypes, n_slots,
                    _capi.current_stream(self.device), ctypes.byref(handle),
                )
            )
        self.handle = handle

    def is_stale(self) -> bool:
        return any(t._version != v for t, v in self.watched)

    def __del__(self) -> None:
        handle, self.handle = self.handle, None
        try:
            if handle is not None and _capi is not None and _capi._lib is not None:
                _capi._lib.ch_program_destroy(handle)
        except Exception:  # interpreter shutdown
            pass


def _type_name(element) -> str:
    return type(element).__name__


def flatten(elements) -> list:
    """Expand nested ``Segment`` / ``Superimposed`` (segment.py:143-157, superimposed.py:72-73)."""
    flat = []
    for element in elements:
        kind = _type_name(element)
        if kind == "Segment":
            flat.extend(flatten(list(element.elements)))
        elif kind == "Superimposed":
            flat.extend(flatten(list(element._segment.elements)))
        else:
            flat.append(element)
    return flat


def _length(element) -> SlotSpec:
    return SlotSpec(element.length, is_length=True)


def _lower_linear(element) -> tuple[int, int, list]:
    """(opcode, flags, slots) of one skippable element -- SURVEY appendix A."""
    kind = _type_name(element)
    if kind in ("Marker", "BPM", "Screen", "Aperture"):
        return _capi.OP_IDENTITY, 0, []
    if kind in ("Drift", "Sextupole"):
        return _capi.OP_DRIFT, 0, [_length(element)]
    if kind == "HorizontalCorrector":
        return _capi.OP_CORRECTOR, 1, [_length(element), SlotSpec(element.angle)]
    if kind == "VerticalCorrector":
        return _capi.OP_CORRECTOR, 2, [_length(element), SlotSpec(element.angle)]
    if kind == "CombinedCorrector":
        return _capi.OP_CORRECTOR, 3, [
            _length(element), SlotSpec(element.horizontal_angle), SlotSpec(element.vertical_angle),
        ]
    if kind == "Quadrupole":
        return _capi.OP_QUADRUPOLE, 0, [
            _length(element), SlotSpec(element.k1), SlotSpec(element.tilt),
            SlotSpec(element.misalignment, 2, 0), SlotSpec(element.misalignment, 2, 1),
        ]
    if kind in ("Dipole", "RBend"):
        return _capi.OP_DIPOLE, 0, [
            _length(element), SlotSpec(element.angle), SlotSpec(element.k1),
            SlotSpec(element.dipole_e1), SlotSpec(element.dipole_e2),
            SlotSpec(element.fringe_integral), SlotSpec(element.fringe_integral_exit),
            SlotSpec(element.gap), SlotSpec(element.tilt),
        ]
    if kind == "Solenoid":
        return _capi.OP_SOLENOID, 0, [
            _length(element), SlotSpec(element.k),
            SlotSpec(element.misalignment, 2, 0), SlotSpec(element.misalignment, 2, 1),
        ]
    if kind == "Undulator":
        return _capi.OP_UNDULATOR, 0, [
            _length(element), SlotSpec(element.period), SlotSpec(element.kx),
            SlotSpec(element.ky),
        ]
    if kind == "Cavity":
        traveling = getattr(element, "cavity_type", "standing_wave") == "traveling_wave"
        return _capi.OP_CAVITY_OFF, int(traveling), [_length(element)]
    if kind == "CustomTransferMap":
        return _capi.OP_CUSTOM_MAP, 0, [
            SlotSpec(element.predefined_transfer_map, 49, 0), _length(element),
        ]
    raise NotImplementedError(
        f"cheetah_b200: element type {kind} has no linear lowering (outside the hot path)"
    )


_zeros: dict = {}


def _zero(device: torch.device) -> torch.Tensor:
    """A device scalar 0 for the unused slots of a fixed slot layout."""
    key = str(device)
    if key not in _zeros:
        _zeros[key] = torch.zeros((), dtype=torch.float32, device=device)
    return _zeros[key]


def nonlinear_method(element) -> str | None:
    """"drift_kick_drift" / "second_order" if ``element`` is tracked per particle by
    ``ch_track_nonlinear``, else None."""
    kind = _type_name(element)
    if kind == "TransverseDeflectingCavity":
        return "drift_kick_drift"
    if kind in ("Drift", "Quadrupole", "Dipole", "RBend", "Sextupole"):
        method = getattr(element, "tracking_method", "linear")
        if method in ("drift_kick_drift", "second_order") and not (
            kind == "Sextupole" and method == "drift_kick_drift"
        ):
            return method
    return None


def _lower_nonlinear(element, device: torch.device) -> tuple[int, int, list]:
    """(opcode, flags, slots) of one drift_kick_drift / second_order element."""
    kind = _type_name(element)
    method = nonlinear_method(element)
    zero = SlotSpec(_zero(device))
    misalignment = lambda: [  # noqa: E731
        SlotSpec(element.misalignment, 2, 0), SlotSpec(element.misalignment, 2, 1)
    ]
    if kind == "TransverseDeflectingCavity":
        # transverse_deflecting_cavity.py:122-209
        return _capi.OP_DKD_TDC, 0, [
            _length(element), SlotSpec(element.voltage), SlotSpec(element.phase),
            SlotSpec(element.frequency), SlotSpec(element.tilt), *misalignment(),
        ]
    if method == "drift_kick_drift":
        if kind == "Drift":  # drift.py:106-154
            return _capi.OP_DKD_DRIFT, 0, [_length(element)]
        if kind == "Quadrupole":  # quadrupole.py:168-251
            num_steps = int(element.num_steps)
            if num_steps < 1:
                raise ValueError(f"Quadrupole {element.name!r}: num_steps must be >= 1")
            return _capi.OP_DKD_QUADRUPOLE, num_steps, [
                _length(element), SlotSpec(element.k1), SlotSpec(element.tilt), *misalignment(),
            ]
        # Dipole / RBend, dipole.py:183-370
        fringe_at = getattr(element, "fringe_at", "both")
        flags = int(fringe_at in ("entrance", "both")) | (int(fringe_at in ("exit", "both")) << 1)
        return _capi.OP_DKD_DIPOLE, flags, [
            _length(element), SlotSpec(element.angle),
            SlotSpec(element.dipole_e1), SlotSpec(element.dipole_e2),
            SlotSpec(element.fringe_integral), SlotSpec(element.fringe_integral_exit),
            SlotSpec(element.gap), SlotSpec(element.gap_exit), SlotSpec(element.tilt),
        ]
    # second order: slots length, k1, k2, angle, e1, e2, fint, fint_exit, gap, tilt, mis_x, mis_y
    if kind == "Drift":  # drift.py:67-83
        return _capi.OP_SECOND_ORDER, 0, [_length(element)] + [zero] * 11
    if kind == "Quadrupole":  # quadrupole.py:112-143
        return _capi.OP_SECOND_ORDER, 0, [
            _length(element), SlotSpec(element.k1), *([zero] * 7), SlotSpec(element.tilt),
            *misalignment(),
        ]
    if kind == "Sextupole":  # sextupole.py:90-116
        return _capi.OP_SECOND_ORDER, 0, [
            _length(element), zero, SlotSpec(element.k2), *([zero] * 6), SlotSpec(element.tilt),
            *misalignment(),
        ]
    # Dipole / RBend, dipole.py:396-466
    return _capi.OP_SECOND_ORDER, 1, [
        _length(element), SlotSpec(element.k1), zero, SlotSpec(element.angle),
        SlotSpec(element.dipole_e1), SlotSpec(element.dipole_e2),
        SlotSpec(element.fringe_integral), SlotSpec(element.fringe_integral_exit),
        SlotSpec(element.gap), SlotSpec(element.tilt), zero, zero,
    ]


def _check_tensor(tensor: torch.Tensor, device: torch.device, what: str) -> None:
    if not isinstance(tensor, torch.Tensor):
        raise TypeError(f"{what} must be a tensor, got {type(tensor)}")
    if tensor.device != device:
        raise ValueError(
            f"{what} lives on {tensor.device} but the beam is on {device}; move the lattice with "
            "`segment.to(device)` first"
        )
    if tensor.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError(
            f"{what} requires grad: the CUDA path is forward-only; differentiable tracking is the "
            "reference implementation's job (SURVEY.md 2, row 12)"
        )


def lower(elements, device: torch.device, target_shape: tuple = ()) -> LatticeProgram:
    """Lower ``elements`` for tensors on ``device``.

    ``target_shape`` is the vector shape contributed from outside the lattice (the beam's
    energy); parameters whose vector shape is neither ``()`` nor the section's full
    broadcast shape are materialised (expanded copies) and watched for staleness.
    """
    ops: list[Op] = []
    stages: list = []
    watched: list = []
    section: LinearSection | None = None
    target_shape = tuple(target_shape)

    def open_section() -> LinearSection:
        nonlocal section
        if section is None:
            section = LinearSection(op_begin=len(ops), op_end=len(ops))
        return section

    def close_section() -> None:
        nonlocal section
        close_run()
        if section is None:
            return
        section.op_end = len(ops)
        _finish_section(section, ops, device, target_shape, watched)
        stages.append(section)
        section = None

    # a non-linear run [run_begin, run_end); identity ops appended after run_end are "pending":
    # they join the run if another non-linear element follows, else they open a linear section
    run_begin = run_end = -1
    run_methods: list = []

    def close_run() -> None:
        nonlocal run_begin, run_end, section, run_methods
        if run_begin < 0:
            return
        run = NonlinearRun(op_begin=run_begin, op_end=run_end, methods=tuple(run_methods))
        _finish_section(run, ops, device, target_shape, watched)
        stages.append(run)
        if run_end < len(ops):  # pending identity ops start the next linear section
            section = LinearSection(op_begin=run_end, op_end=len(ops))
        run_begin = run_end = -1
        run_methods = []

    def is_identity(element) -> bool:
        return element.is_skippable and _type_name(element) in ("Marker", "BPM", "Screen", "Aperture")

    for element in flatten(elements):
        kind = _type_name(element)
        if kind == "Cavity":
            watched.append((element.voltage, element.voltage._version))
        method = nonlinear_method(element)
        if method is not None:
            if run_begin < 0:
                close_section()
                run_begin = len(ops)
            elif len(ops) + 1 - run_begin > _capi.NL_MAX_OPS:
                close_run()
                if section is not None:  # pending identities became a linear section
                    close_section()
                run_begin = len(ops)
            opcode, flags, slots = _lower_nonlinear(element, device)
            ops.append(Op(opcode, flags, slots, element))
            run_end = len(ops)
            run_methods.append(method)
            continue
        if run_begin >= 0:
            if is_identity(element):
                ops.append(Op(_capi.OP_IDENTITY, 0, [], element))
                continue
            close_run()
        if element.is_skippable:
            opcode, flags, slots = _lower_linear(element)
            sec = open_section()
            if opcode != _capi.OP_IDENTITY:
                sec.has_maps = True
            ops.append(Op(opcode, flags, slots, element))
        elif kind == "Aperture":
            shape = element.shape
            assert shape in ("rectangular", "elliptical"), f"Unknown aperture shape {shape}"
            sec = open_section()
            if sec.n_apertures == _capi.MAX_APERTURES:
                close_section()
                sec = open_section()
            if shape == "elliptical":
                sec.elliptical_mask |= 1 << sec.n_apertures
            sec.n_apertures += 1
            sec.apertures.append(element)
            ops.append(
                Op(_capi.OP_APERTURE, int(shape == "elliptical"),
                   [SlotSpec(element.x_max), SlotSpec(element.y_max)], element)
            )
        elif kind == "Cavity":
            # active cavity: its R matrix joins the section's cumulative map, its non-linear
            # longitudinal tail is applied by the apply kernel; the beam energy changes, so
            # the section ends here (cavity.py:100-251)
            sec = open_section()
            sec.has_maps = True
            gain_flag = torch.zeros((), dtype=torch.float32, device=device)
            traveling = getattr(element, "cavity_type", "standing_wave") == "traveling_wave"
            ops.append(
                Op(_capi.OP_CAVITY, int(traveling),
                   [_length(element), SlotSpec(element.voltage), SlotSpec(element.phase),
                    SlotSpec(element.frequency), SlotSpec(gain_flag)], element)
            )
            sec.cavity = (element, gain_flag)
            close_section()
            # the outgoing energy is energy + voltage cos(phase) q (cavity.py:122): every later
            # section sees an energy that also carries the cavity's vector dimensions, and its
            # slot strides must be resolved against that wider batch
            target_shape = tuple(torch.broadcast_shapes(
                tuple(target_shape), tuple(element.voltage.shape), tuple(element.phase.shape)))
        elif kind == "SpaceChargeKick":
            close_section()
            stages.append(Barrier(element, "space_charge"))
        elif kind in ("BPM", "Screen"):  # active monitors record a reading between sections
            close_section()
            stages.append(Barrier(element, kind.lower()))
        else:
            close_section()
            stages.append(Barrier(element, "unsupported"))
    close_section()
    return LatticeProgram(ops, stages, device, watched)


def _finish_section(section: LinearSection, ops: list, device, target_shape, watched) -> None:
    """Resolve vector shapes and per-slot strides of one linear section."""
    shapes, length_shapes = [tuple(target_shape)], []
    survival_shape: tuple = ()
    running: tuple = ()
    map_shape: tuple = ()
    n_map_ops, maps_before_last_aperture = 0, False
    for op in ops[section.op_begin : section.op_end]:
        for slot in op.slots:
            _check_tensor(slot.tensor, device, f"parameter of element {op.element.name!r}")
            shapes.append(slot.vector_shape)
            running = torch.broadcast_shapes(running, slot.vector_shape)
            if op.opcode != _capi.OP_APERTURE:
                map_shape = torch.broadcast_shapes(map_shape, slot.vector_shape)
            if slot.is_length:
                length_shapes.append(slot.vector_shape)
        if op.opcode == _capi.OP_APERTURE:
            survival_shape = running
            maps_before_last_aperture = n_map_ops > 0
        else:
            n_map_ops += 1
    full = tuple(torch.broadcast_shapes(*shapes))
    section.lattice_shape = full
    section.length_shape = tuple(torch.broadcast_shapes(*length_shapes)) if length_shapes else ()
    section.survival_shape = tuple(survival_shape)
    if isinstance(section, LinearSection):
        section.map_shape = tuple(map_shape)
        section.n_map_ops = n_map_ops
        section.maps_before_last_aperture = maps_before_last_aperture

    for op in ops[section.op_begin : section.op_end]:
        resolved = []
        for slot in op.slots:
            tensor = slot.tensor
            if tensor.dtype not in (torch.float32, torch.float64):
                watched.append((tensor, tensor._version))
                tensor = tensor.to(torch.float64)
            vshape = slot.vector_shape
            if math.prod(vshape) == 1:
                stride = 0
                if not tensor.is_contiguous():
                    watched.append((slot.tensor, slot.tensor._version))
                    tensor = tensor.contiguous()
            elif vshape == full and tensor.is_contiguous():
                stride = slot.inner_size
            else:  # partial broadcast or strided view: expanded copy, watched for staleness
                watched.append((slot.tensor, slot.tensor._version))
                inner = tuple(tensor.shape[len(vshape):])
                tensor = tensor.expand(*full, *inner).contiguous()
                stride = slot.inner_size
            resolved.append((tensor, stride, slot.inner_offset))
        op.resolved = resolved
