"""ctypes binding of ``include/cheetah_b200.h``.

The CUDA library is the product: there is no Python or CPU fallback.  ``lib()`` raises if
``libcheetah_b200.so`` has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``python -m cheetah_b200.build``).
"""

from __future__ import annotations

import ctypes
from ctypes import c_double, POINTER, c_char_p, c_int32, c_int64, c_uint32, c_void_p
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libcheetah_b200.so"

CH_F32, CH_F64 = 0, 1
ABI_VERSION = 3  # CH_ABI_VERSION of include/cheetah_b200.h

OP_IDENTITY = 0
OP_DRIFT = 1
OP_CORRECTOR = 2
OP_QUADRUPOLE = 3
OP_DIPOLE = 4
OP_SOLENOID = 5
OP_UNDULATOR = 6
OP_CAVITY_OFF = 7
OP_CUSTOM_MAP = 8
OP_APERTURE = 9
OP_CAVITY = 10
OP_DKD_DRIFT = 11
OP_DKD_QUADRUPOLE = 12
OP_DKD_DIPOLE = 13
OP_DKD_TDC = 14
OP_SECOND_ORDER = 15
NONLINEAR_OPS = (OP_DKD_DRIFT, OP_DKD_QUADRUPOLE, OP_DKD_DIPOLE, OP_DKD_TDC, OP_SECOND_ORDER)
NL_HEADER = 12
NL_MAX_OPS = 64

RECORD_HEADER = 2
RECORD_MAP = 42
RECORD_APERTURE = 16
RECORD_CAVITY = 24
MAX_APERTURES = 32


def record_len(n_apertures: int, cavity: bool = False) -> int:
    return RECORD_HEADER + RECORD_MAP + RECORD_APERTURE * n_apertures + (RECORD_CAVITY if cavity else 0)


# name -> (restype, argtypes); must list every symbol declared in include/cheetah_b200.h
SIGNATURES = {
    "ch_abi_version": (c_int32, []),
    "ch_last_error": (c_char_p, []),
    "ch_kernel_launch_count": (c_int64, []),
    "ch_program_create": (
        c_int32,
        [
            POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), c_int32,
            POINTER(c_void_p), POINTER(c_int64), POINTER(c_int32), c_int32,
            c_void_p, POINTER(c_void_p),
        ],
    ),
    "ch_program_destroy": (c_int32, [c_void_p]),
    "ch_compose_maps": (
        c_int32,
        [
            c_void_p, c_int32, c_int32, c_int64,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p,
        ],
    ),
    "ch_apply_maps": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_int64, c_int32, c_uint32,
            c_int64, c_int64,
            c_void_p, c_void_p,
            c_int32, c_int32, c_void_p,
        ],
    ),
    "ch_apply_maps_compact": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_int64, c_int32, c_uint32,
            c_int64, c_int64,
            c_void_p, c_void_p, c_void_p,
            c_int32, c_void_p,
        ],
    ),
    "ch_apply_maps_moments": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_int64, c_int32, c_uint32,
            c_int64, c_int64,
            c_void_p, c_void_p, c_void_p,
            c_int32, c_int32, c_void_p,
        ],
    ),
    "ch_apply_maps_parameter": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_void_p, c_int64,
            c_void_p, c_int64, c_void_p,
            c_int32, c_int64, c_void_p, c_void_p, c_int32, c_void_p,
        ],
    ),
    "ch_apply_maps_covariance": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_int64, c_int32, c_uint32,
            c_int64, c_int64,
            c_void_p, c_void_p, c_void_p,
            c_int32, c_int32, c_void_p,
        ],
    ),
    "ch_nonlinear_constants_len": (c_int64, [c_void_p, c_int32, c_int32]),
    "ch_nonlinear_constants": (
        c_int32,
        [
            c_void_p, c_int32, c_int32, c_int64,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_track_nonlinear": (
        c_int32,
        [
            c_void_p, c_int32, c_int32,
            c_void_p, c_int64, c_void_p,
            c_void_p, c_int64, c_void_p,
            c_int64, c_int64, c_void_p,
            c_int32, c_void_p,
        ],
    ),
    "ch_screen_image": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
            c_int64, c_int64, c_int32, c_void_p, c_void_p,
        ],
    ),
    "ch_screen_kde": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_int32, c_void_p, c_int32, c_void_p,
            c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_screen_gaussian": (
        c_int32,
        [
            c_void_p, c_void_p, c_void_p, c_double, c_double, c_int32, c_double, c_double,
            c_int32, c_int32, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_beam_moments": (
        c_int32,
        [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p],
    ),
    "ch_sc_moments_and_params": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_moments_and_params_deterministic": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_grid_params": (
        c_int32,
        [
            c_void_p, c_int64,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_sc_deposit": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_sc_deposit_deterministic": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_cic_deposit_deterministic": (
        c_int32,
        [
            c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_cic_deposit3d": (
        c_int32,
        [
            c_void_p, c_void_p, c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_cic_deposit": (
        c_int32,
        [
            c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_sc_green_function": (
        c_int32,
        [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p],
    ),
    "ch_sc_green_spectrum": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
         c_void_p],
    ),
    "ch_sc_poisson_solve": (
        c_int32,
        [
            c_void_p, c_void_p, c_void_p, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_field": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p],
    ),
    "ch_sc_field_bricks": (
        c_int32,
        [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p],
    ),
    "ch_sc_gather_kick": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_gather_kick_fused": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_void_p,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
            c_int32, c_int32, c_int32,
            c_void_p, c_void_p,
        ],
    ),
    "ch_sc_solve": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
        ],
    ),
    "ch_sc_field_gather": (
        c_int32,
        [
            c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int64,
            c_int32, c_int32, c_int32, c_int32,
            c_void_p, c_int64, c_void_p, c_int64,
            c_void_p, c_void_p,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int32,
            c_void_p, c_int64, c_int32,
            c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32,
            c_int32, c_int32, c_int32,
            c_void_p, c_void_p, c_void_p,
        ],
    ),
}

MOMENTS = 20
MOMENTS_COV = 36
SC_STATS = 12
SC_PARAMS = 24
SC_FIELD_NODES = 0
SC_FIELD_BRICKS = 1

_lib = None


class BackendError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"cheetah_b200: CUDA library {LIB_PATH} is missing. Build it with "
                "`python -m cheetah_b200.build`; there is no CPU fallback."
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        handle.ch_abi_version.restype = c_int32
        if handle.ch_abi_version() != ABI_VERSION:
            raise RuntimeError(
                f"cheetah_b200: {LIB_PATH} has ABI version {handle.ch_abi_version()}, the Python "
                f"binding expects {ABI_VERSION}; rebuild it with `python -m cheetah_b200.build`"
            )
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        message = lib().ch_last_error().decode(errors="replace")
        raise BackendError(f"cheetah_b200 C-ABI error {status}: {message}")


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return CH_F32
    if dtype == torch.float64:
        return CH_F64
    raise TypeError(f"cheetah_b200 supports float32 and float64 tensors, got {dtype}")


def current_stream(device: torch.device) -> int:
    """Raw cudaStream_t of torch's current stream on ``device`` (fast path, no Stream object)."""
    index = device.index
    if index is None:
        index = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(index)


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device: torch.device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the context
    manager costs ~8 us per use, as much as a whole C-ABI call)."""
    index = device.index
    if index is None or index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def launch_count() -> int:
    return int(lib().ch_kernel_launch_count())


def ptr(tensor: torch.Tensor | None) -> int | None:
    return None if tensor is None else tensor.data_ptr()
