"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d), shared by bench.py, tools/ and tests.

The ARES lattice comes from ``tests/golden/ares_lattice.json`` (the reference's
``docs/examples/ARESlatticeStage3v1_9.json`` converted by ``oracle/make_golden.py``: 195
elements).  A workload is returned as a plain-dict lattice description
(``cheetah_b200/lattice_description.py`` format) plus beam parameters, and can be instantiated either as ``cheetah_b200`` objects on a
CUDA device (the product) or kept as dicts for the CPU oracle (the baseline).

Config 3 recipe (frozen; SURVEY.md 8d asks to tune the ranges once so that the mean survival
over the settings is 30-70 %): generator seed 1; the 13 quadrupoles' k1 ~ U(-5, 5) 1/m^2 and
the 30 correctors' angle ~ U(-2e-5, 2e-5) rad, drawn element by element in lattice order;
apertures ARLISLHG1 and ARBCSLHB1 rectangular with x_max = y_max = 2 mm.  With the
``from_twiss`` beam (beta_x 3.14 m, beta_y 42 m, 1e8 eV) the mean survival is 42 %.  (The
U(-30, 30) / U(-1e-3, 1e-3) ranges first suggested by the survey lose every particle.)
"""

from __future__ import annotations

from pathlib import Path

import torch

from cheetah_b200 import lattice_description

GOLDEN = Path(__file__).resolve().parent / "tests" / "golden"
N_ELEMENTS_ARES = 195

CONFIG2_SETTINGS = {  # reference README.md:73-77
    "AREAMQZM1": ("k1", 8.2),
    "AREAMQZM2": ("k1", -14.3),
    "AREAMCVM1": ("angle", 9e-5),
    "AREAMQZM3": ("k1", 3.142),
    "AREAMCHM1": ("angle", -1e-4),
}


def _set(description: list, name: str, attr: str, value) -> None:
    for element in description:
        if element["name"] == name:
            element[attr] = value
            return
    raise KeyError(name)


def ares_config2(dtype=torch.float32) -> list:
    """ARES, single setting, five EA magnets powered (BASELINE configs[1])."""
    lattice = lattice_description.load(GOLDEN / "ares_lattice.json", dtype)
    for name, (attr, value) in CONFIG2_SETTINGS.items():
        _set(lattice, name, attr, torch.tensor(value, dtype=dtype))
    return lattice


def ares_config3(n_settings: int, dtype=torch.float32, begin: int = 0, end: int | None = None) -> list:
    """ARES with ``n_settings`` vectorised magnet settings (BASELINE configs[2]).

    ``[begin, end)`` selects a contiguous shard of the settings (multi-GPU / CPU samples): the
    random draw is always made for the full batch so that shards of different world sizes see
    the same settings.
    """
    end = n_settings if end is None else end
    lattice = lattice_description.load(GOLDEN / "ares_lattice.json", dtype)
    g = torch.Generator().manual_seed(1)
    for element in lattice:
        if element["type"] == "Quadrupole":
            full = (torch.rand(n_settings, generator=g) * 2 - 1) * 5.0
            element["k1"] = full[begin:end].to(dtype).contiguous()
        elif element["type"] in ("HorizontalCorrector", "VerticalCorrector"):
            full = (torch.rand(n_settings, generator=g) * 2 - 1) * 2e-5
            element["angle"] = full[begin:end].to(dtype).contiguous()
    for name in ("ARLISLHG1", "ARBCSLHB1"):
        _set(lattice, name, "x_max", torch.tensor(2e-3, dtype=dtype))
        _set(lattice, name, "y_max", torch.tensor(2e-3, dtype=dtype))
    return lattice


def ares_config3_dense(n_settings: int, dtype=torch.float32, begin: int = 0,
                       end: int | None = None) -> list:
    """Config 3 with x-y coupling: both gun solenoids powered (k = 0.5 / -0.4 1/m) and the
    quadrupole AREAMQZM2 tilted by 0.3 rad, so that none of the sparsity flags of the apply kernel
    holds and every setting takes its generic 72-multiply-add branch (solenoid.py:74-116,
    track_methods.py:345-382)."""
    lattice = ares_config3(n_settings, dtype, begin, end)
    _set(lattice, "ARLIMSOG1A", "k", torch.tensor(0.5, dtype=dtype))
    _set(lattice, "ARLIMSOG1B", "k", torch.tensor(-0.4, dtype=dtype))
    _set(lattice, "AREAMQZM2", "tilt", torch.tensor(0.3, dtype=dtype))
    return lattice


def ares_config3_tau_coupled(n_settings: int, dtype=torch.float32, begin: int = 0,
                             end: int | None = None) -> list:
    """`ares_config3_dense` followed by a short CustomTransferMap whose matrix couples tau into
    x and changes the delta row (what an off-crest RF structure or a dipole with RF does to a
    merged map): no sparsity flag holds, every setting takes the 72-multiply-add branch."""
    lattice = ares_config3_dense(n_settings, dtype, begin, end)
    matrix = torch.eye(7, dtype=dtype)
    matrix[0, 4], matrix[1, 4], matrix[2, 4] = 2e-3, -1e-3, 5e-4
    matrix[5, 4], matrix[5, 0], matrix[4, 5] = 3e-3, 1e-3, 0.2
    lattice.append({"type": "CustomTransferMap", "name": "tau_coupling",
                    "length": torch.tensor(0.1, dtype=dtype),
                    "predefined_transfer_map": matrix})
    return lattice


def ares_survey_ranges(n_settings: int, dtype=torch.float32, x_max: float = 0.5) -> list:
    """The ranges SURVEY 8d first proposed -- k1 ~ U(-30, 30) 1/m^2, corrector angles ~
    U(-1e-3, 1e-3) rad, seed 1 -- with the two apertures opened to `x_max` instead of narrowing
    the ranges: strongly over-focused settings with large map entries (a parity case, not the
    bench workload)."""
    lattice = lattice_description.load(GOLDEN / "ares_lattice.json", dtype)
    g = torch.Generator().manual_seed(1)
    for element in lattice:
        if element["type"] == "Quadrupole":
            element["k1"] = ((torch.rand(n_settings, generator=g) * 2 - 1) * 30.0).to(dtype)
        elif element["type"] in ("HorizontalCorrector", "VerticalCorrector"):
            element["angle"] = ((torch.rand(n_settings, generator=g) * 2 - 1) * 1e-3).to(dtype)
    for name in ("ARLISLHG1", "ARBCSLHB1"):
        _set(lattice, name, "x_max", torch.tensor(x_max, dtype=dtype))
        _set(lattice, name, "y_max", torch.tensor(x_max, dtype=dtype))
    return lattice


def twiss_beam_particles(num_particles: int, seed: int = 0) -> torch.Tensor:
    """(N, 7) float64 particles of the README beam (beta_x 3.14, beta_y 42, defaults of
    particle_beam.py:464-504), generated on the CPU so that every rank and the CPU baseline
    see identical inputs."""
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_twiss(
        num_particles=num_particles, beta_x=3.14, beta_y=42.0, dtype=torch.float64,
        generator=torch.Generator().manual_seed(seed),
    )
    return beam.particles


def product_segment(description: list, device, dtype):
    import cheetah_b200 as cb

    return cb.Segment(elements=lattice_description.build(description, device=device, dtype=dtype))


def product_beam(particles: torch.Tensor, device, dtype, energy: float = 1e8):
    import cheetah_b200 as cb

    beam = cb.ParticleBeam(
        particles=particles.to(device=device, dtype=dtype),
        energy=torch.tensor(energy, device=device, dtype=dtype),
        species=cb.Species("electron", device=device, dtype=dtype),
    )
    beam._unit_seventh = True
    return beam


def oracle_beam(particles: torch.Tensor, dtype, energy: float = 1e8) -> dict:
    from oracle import track_oracle as oracle

    return oracle.make_beam(particles.to(dtype), torch.tensor(energy, dtype=dtype))


def fodo_space_charge(cells: int = 50, grid: int = 64, dtype=torch.float32) -> list:
    """BASELINE configs[3]: `cells` FODO cells [QF, D, QD, D]; every 1 m drift D is realised as
    Drift(0.5) . SpaceChargeKick(effect_length=1, grid^3) . Drift(0.5) (the split-kick pattern of
    the reference's tests/test_space_charge_kick.py:56-66) -> 8 elements and 2 kicks per cell."""
    t = lambda v: torch.tensor(v, dtype=dtype)  # noqa: E731
    lattice = []
    for cell in range(cells):
        for sign, tag in ((1.0, "f"), (-1.0, "d")):
            lattice.append({"type": "Quadrupole", "name": f"q{tag}{cell}", "length": t(0.2),
                            "k1": t(4.2 * sign), "misalignment": torch.zeros(2, dtype=dtype),
                            "tilt": t(0.0), "tracking_method": "linear"})
            lattice.append({"type": "Drift", "name": f"d{tag}{cell}a", "length": t(0.5),
                            "tracking_method": "linear"})
            lattice.append({"type": "SpaceChargeKick", "name": f"sc{tag}{cell}",
                            "effect_length": t(1.0), "grid_extent_x": t(3.0),
                            "grid_extent_y": t(3.0), "grid_extent_tau": t(3.0),
                            "grid_shape": [grid, grid, grid]})
            lattice.append({"type": "Drift", "name": f"d{tag}{cell}b", "length": t(0.5),
                            "tracking_method": "linear"})
    return lattice


def parameters_beam_particles(num_particles: int, seed: int = 0) -> torch.Tensor:
    """(N, 7) float64 particles of ParticleBeam.from_parameters defaults (sigma_x = sigma_y =
    175 um, sigma_px = sigma_py = 4e-6, sigma_tau = 8 um, sigma_p = 2e-3; particle_beam.py:199-216)."""
    import cheetah_b200 as cb

    beam = cb.ParticleBeam.from_parameters(
        num_particles=num_particles, dtype=torch.float64,
        generator=torch.Generator().manual_seed(seed),
    )
    return beam.particles
