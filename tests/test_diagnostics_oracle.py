"""Pin the diagnostics oracle (Screen image, BPM reading) against outputs
of the unmodified reference (tests/golden/diagnostics.npz, oracle/make_golden.py)."""

import json

import pytest
import torch

from oracle import diagnostics_oracle as diag
from oracle import track_oracle as oracle

from . import golden_utils as gu

ARRAYS = gu.load_npz("diagnostics.npz")
with (gu.GOLDEN / "diagnostics.json").open() as f:
    SCREENS = {k.split(".", 1)[1]: v for k, v in json.load(f).items()}

TAGS = [("f64", torch.float64), ("f32", torch.float32)]
PARTICLE_SCREENS = sorted(k for k in SCREENS if not k.startswith("gaussian"))
GAUSSIAN_SCREENS = sorted(k for k in SCREENS if k.startswith("gaussian"))


def segment_lattice(dtype) -> list:
    t = lambda v: torch.tensor(v, dtype=dtype)  # noqa: E731
    return [
        {"type": "HorizontalCorrector", "name": "h", "length": t(0.1), "angle": t([0.0, 1e-3, -2e-3])},
        {"type": "Drift", "name": "d", "length": t(0.5)},
    ]


@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("name", PARTICLE_SCREENS)
def test_screen_images(name, tag, dtype):
    beam = gu.beam_dict(ARRAYS, "incoming", dtype)
    image = diag.screen_reading(SCREENS[name], beam)
    expected = gu.tensor(ARRAYS[f"screen.{name}.{tag}"], dtype)
    assert image.shape == expected.shape
    # same ops in the same order on the same machine: equal up to scatter_add summation order
    assert torch.allclose(image, expected, rtol=1e-12 if dtype == torch.float64 else 1e-5,
                          atol=float(expected.max()) * (1e-14 if dtype == torch.float64 else 1e-6))
    assert torch.isclose(image.sum(), expected.sum(), rtol=1e-5)


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_bpm_and_vectorised_segment(tag, dtype):
    beam = gu.beam_dict(ARRAYS, "incoming", dtype)
    reading = diag.bpm_reading({"misalignment": (0.1, 0.2)}, beam)
    tol = 1e-12 if dtype == torch.float64 else 1e-5
    assert torch.allclose(reading, gu.tensor(ARRAYS[f"bpm.{tag}"], dtype), rtol=tol, atol=0)
    out = oracle.track(segment_lattice(dtype), beam)
    reading = diag.bpm_reading({"misalignment": (1e-4, 2e-4)}, out)
    expected = gu.tensor(ARRAYS[f"segment.bpm.{tag}"], dtype)
    assert reading.shape == expected.shape == (3, 2)
    assert torch.allclose(reading, expected, rtol=tol, atol=1e-9 if dtype == torch.float32 else 1e-15)
    image = diag.screen_reading(
        {"resolution": (96, 64), "pixel_size": (2.5e-5, 3e-5), "method": "cloud-in-cell"}, out)
    expected = gu.tensor(ARRAYS[f"segment.screen.{tag}"], dtype)
    assert image.shape == expected.shape == (3, 64, 96)
    assert torch.allclose(image, expected, rtol=1e-10 if dtype == torch.float64 else 2e-3,
                          atol=float(expected.max()) * (1e-12 if dtype == torch.float64 else 2e-3))


@pytest.mark.parametrize("tag,dtype", TAGS)
@pytest.mark.parametrize("name", GAUSSIAN_SCREENS)
def test_parameter_beam_images(name, tag, dtype):
    """Analytic bivariate-normal image of a ParameterBeam (screen.py:251-289)."""
    mu = gu.tensor(ARRAYS[f"parameter_beam.mu.{tag}"], dtype)
    cov = gu.tensor(ARRAYS[f"parameter_beam.cov.{tag}"], dtype)
    image = diag.parameter_beam_image(SCREENS[name], mu, cov)
    expected = gu.tensor(ARRAYS[f"screen.{name}.{tag}"], dtype)
    assert image.shape == expected.shape
    assert torch.allclose(image, expected, rtol=1e-9 if dtype == torch.float64 else 2e-4,
                          atol=float(expected.max()) * (1e-12 if dtype == torch.float64 else 1e-5))


@pytest.mark.parametrize("tag,dtype", TAGS)
def test_vectorised_kde_screen(tag, dtype):
    """method="kde" is the reference's vectorised Screen (tests/test_vectorized.py:305-352)."""
    beam = gu.beam_dict(ARRAYS, "incoming", dtype)
    out = oracle.track(segment_lattice(dtype), beam)
    image = diag.screen_reading(
        {"resolution": (96, 64), "pixel_size": (2.5e-5, 3e-5), "method": "kde"}, out)
    expected = gu.tensor(ARRAYS[f"segment.kde_screen.{tag}"], dtype)
    assert image.shape == expected.shape == (3, 64, 96)
    assert torch.allclose(image, expected, rtol=1e-9 if dtype == torch.float64 else 2e-3,
                          atol=float(expected.max()) * (1e-12 if dtype == torch.float64 else 2e-4))
    assert torch.allclose(image.sum(dim=(-2, -1)), torch.ones(3, dtype=dtype), atol=1e-4)
