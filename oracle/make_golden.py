"""Generate the golden fixtures under ``tests/golden/`` from the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the only place where
``/root/reference`` exists):

    PYTHONPATH=oracle/refshim:/root/reference python oracle/make_golden.py

It (1) converts the reference's own golden pickles for the hot path
(``tests/resources/consistency_expected_outgoing/*.pkl``, pinned by the reference's
``tests/test_elements.py:356-431``) into ``tests/golden/consistency.npz`` +
``consistency.json`` and (2) runs the reference itself on ARES, aperture, cloud-in-cell
and space-charge cases, storing inputs and outputs.  Nothing at test / bench time reads
``/root/reference``; only the committed fixtures are used.
"""

from __future__ import annotations

import json
import pickle
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO / "oracle" / "refshim"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(REPO))

import cheetah  # noqa: E402  (the reference)

from oracle import lattice_io  # noqa: E402

OUT = REPO / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
warnings.simplefilter("ignore")

ROW_STRIDE = 4  # expected outputs are stored for every 4th particle to keep fixtures small


def np64(t: torch.Tensor) -> np.ndarray:
    return t.detach().to(torch.float64).cpu().numpy()


def beam_arrays(prefix: str, beam, rows=slice(None)) -> dict:
    return {
        f"{prefix}.particles": np64(beam.particles[..., rows, :]),
        f"{prefix}.energy": np64(beam.energy),
        f"{prefix}.particle_charges": np64(beam.particle_charges[..., rows]),
        f"{prefix}.survival_probabilities": np64(beam.survival_probabilities[..., rows]),
        f"{prefix}.s": np64(beam.s),
        f"{prefix}.mass_eV": np64(beam.species.mass_eV),
        f"{prefix}.num_elementary_charges": np64(beam.species.num_elementary_charges),
    }


# --------------------------------------------------------------------------------------
# 1. the reference's own golden pickles (tests/test_elements.py:356-431)
# --------------------------------------------------------------------------------------
CONSISTENCY_CASES = {
    # (class name, label): constructor kwargs -- tests/conftest.py:12-152, hot-path rows only
    ("Aperture", "inactive"): {"is_active": False},
    ("Aperture", "active"): {"is_active": True},
    ("BPM", "inactive"): {"is_active": False},
    ("Cavity", "default"): {"length": torch.tensor(1.0)},
    ("CombinedCorrector", "default"): {
        "length": torch.tensor(1.0),
        "horizontal_angle": torch.tensor([1.0, -2.0]),
        "vertical_angle": torch.tensor([1.0, -2.0]),
    },
    ("CustomTransferMap", "identity"): {"predefined_transfer_map": torch.eye(7)},
    ("Dipole", "linear"): {
        "length": torch.tensor(1.0),
        "angle": torch.tensor([1.0, -2.0]),
        "tilt": torch.tensor(0.42),
        "tracking_method": "linear",
    },
    ("Drift", "linear"): {"length": torch.tensor([1.0, -1.0]), "tracking_method": "linear"},
    ("HorizontalCorrector", "default"): {
        "length": torch.tensor(1.0),
        "angle": torch.tensor([1.0, -2.0]),
    },
    ("Marker", "default"): {},
    ("Quadrupole", "linear"): {
        "length": torch.tensor(1.0),
        "k1": torch.tensor([1.0, -2.0]),
        "tilt": torch.tensor(0.42),
        "misalignment": torch.tensor([0.01, -0.02]),
        "tracking_method": "linear",
    },
    ("RBend", "linear"): {
        "length": torch.tensor(1.0),
        "angle": torch.tensor([1.0, -2.0]),
        "tilt": torch.tensor(0.42),
        "tracking_method": "linear",
    },
    ("Screen", "default"): {},
    ("Sextupole", "linear"): {
        "length": torch.tensor(1.0),
        "k2": torch.tensor([1.0, -2.0]),
        "tilt": torch.tensor(0.42),
        "misalignment": torch.tensor([0.01, -0.02]),
        "tracking_method": "linear",
    },
    ("Solenoid", "default"): {
        "length": torch.tensor(1.0),
        "k": torch.tensor([1.0, -2.0]),
        "misalignment": torch.tensor([0.01, -0.02]),
    },
    ("SpaceChargeKick", "default"): {"effect_length": torch.tensor(1.0)},
    ("Undulator", "default"): {
        "length": torch.tensor(1.0),
        "period": torch.tensor(0.1),
        "kx": torch.tensor(1.3),
    },
    ("VerticalCorrector", "default"): {
        "length": torch.tensor(1.0),
        "angle": torch.tensor([1.0, -2.0]),
    },
}


def make_consistency() -> None:
    resources = REF / "tests" / "resources"
    with (resources / "ACHIP_EA1_2021.1351.001_subsampled_3000.pkl").open("rb") as f:
        incoming = pickle.load(f).to(torch.float64)
    arrays = beam_arrays("incoming", incoming)
    parameter_incoming = incoming.as_parameter_beam()
    arrays["incoming.mu"] = np64(parameter_incoming.mu)
    arrays["incoming.cov"] = np64(parameter_incoming.cov)
    arrays["incoming.total_charge"] = np64(parameter_incoming.total_charge)

    lattices = {}
    rows = slice(None, None, ROW_STRIDE)
    containers = {  # tests/conftest.py:100-102, :117-124
        ("Segment", "default"): lambda: cheetah.Segment(
            elements=[cheetah.Drift(length=torch.tensor(1.0))], name="default"),
        ("Superimposed", "default"): lambda: cheetah.Superimposed(
            base_element=cheetah.Quadrupole(length=torch.tensor(1.0), k1=torch.tensor(0.5)),
            superimposed_element=cheetah.BPM(), name="default"),
    }
    cases = dict(CONSISTENCY_CASES)
    cases.update({key: None for key in containers})
    for (cls_name, label), kwargs in cases.items():
        key = f"{cls_name}_{label}"
        if kwargs is None:
            element = containers[(cls_name, label)]().to(torch.float64)
        else:
            element = getattr(cheetah, cls_name)(name=label, **kwargs).to(torch.float64)
        lattices[key] = lattice_io._to_json([lattice_io.describe(element)])
        folder = resources / "consistency_expected_outgoing"
        with (folder / f"{cls_name}_ParticleBeam_{label}.pkl").open("rb") as f:
            expected = pickle.load(f)
        arrays.update(beam_arrays(f"{key}.expected", expected, rows))
        parameter_pickle = folder / f"{cls_name}_ParameterBeam_{label}.pkl"
        if parameter_pickle.exists():
            with parameter_pickle.open("rb") as f:
                expected_parameter = pickle.load(f)
            arrays[f"{key}.expected.mu"] = np64(expected_parameter.mu)
            arrays[f"{key}.expected.cov"] = np64(expected_parameter.cov)
        # Sanity: the reference in this container still reproduces its own pickle
        actual = element.track(incoming)
        assert torch.allclose(actual.particles, expected.particles), key
        assert torch.allclose(
            actual.survival_probabilities, expected.survival_probabilities
        ), key

    np.savez_compressed(OUT / "consistency.npz", **arrays)
    with (OUT / "consistency.json").open("w") as f:
        json.dump({"row_stride": ROW_STRIDE, "lattices": lattices}, f, separators=(",", ":"))
    print("consistency:", len(cases), "cases")


# --------------------------------------------------------------------------------------
# 2. ARES through the reference itself
# --------------------------------------------------------------------------------------
ARES_JSON = REF / "docs" / "examples" / "ARESlatticeStage3v1_9.json"
CONFIG2_SETTINGS = {  # README.md:73-77
    "AREAMQZM1": ("k1", 8.2),
    "AREAMQZM2": ("k1", -14.3),
    "AREAMCVM1": ("angle", 9e-5),
    "AREAMQZM3": ("k1", 3.142),
    "AREAMCHM1": ("angle", -1e-4),
}


def ares_beam(num_particles: int, dtype) -> "cheetah.ParticleBeam":
    torch.manual_seed(0)
    beam = cheetah.ParticleBeam.from_twiss(
        num_particles=num_particles,
        beta_x=torch.tensor(3.14),
        beta_y=torch.tensor(42.0),
        dtype=torch.float64,
    )
    return beam.to(dtype)


def vectorised_settings(segment, batch: int, seed: int = 1) -> dict:
    """Config-3 recipe (DESIGN.md "Workloads"): quads U(-5,5) 1/m^2, correctors
    U(-2e-5,2e-5) rad, drawn element by element from one seeded generator."""
    g = torch.Generator().manual_seed(seed)
    settings = {}
    for element in segment.elements:
        if isinstance(element, cheetah.Quadrupole):
            settings[element.name] = ("k1", (torch.rand(batch, generator=g) * 2 - 1) * 5.0)
        elif isinstance(element, (cheetah.HorizontalCorrector, cheetah.VerticalCorrector)):
            settings[element.name] = (
                "angle",
                (torch.rand(batch, generator=g) * 2 - 1) * 2e-5,
            )
    return settings


def make_ares() -> None:
    num_particles = 2048
    rows = slice(None, None, ROW_STRIDE)
    arrays = {}
    description = None
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        beam = ares_beam(num_particles, dtype)
        if tag == "f64":
            arrays.update(beam_arrays("incoming", beam))

        # case A: lattice file values (all magnets off) -- BASELINE config 1 shape
        segment = cheetah.Segment.from_lattice_json(str(ARES_JSON), dtype=dtype)
        if description is None:
            description = lattice_io.describe(segment)["elements"]
        out = segment.track(beam)
        arrays.update(beam_arrays(f"default.{tag}", out, rows))

        # case B: the five EA magnets of config 2
        for name, (attr, value) in CONFIG2_SETTINGS.items():
            setattr(getattr(segment, name), attr, torch.tensor(value, dtype=dtype))
        out = segment.track(beam)
        arrays.update(beam_arrays(f"config2.{tag}", out, rows))

        # case C: vectorised settings + finite apertures (config 3 recipe, small batch)
        segment = cheetah.Segment.from_lattice_json(str(ARES_JSON), dtype=dtype)
        batch = 6
        for name, (attr, value) in vectorised_settings(segment, batch).items():
            setattr(getattr(segment, name), attr, value.to(dtype))
        segment.ARLISLHG1.x_max = torch.tensor(2e-3, dtype=dtype)
        segment.ARLISLHG1.y_max = torch.tensor(2e-3, dtype=dtype)
        segment.ARBCSLHB1.x_max = torch.tensor(2e-3, dtype=dtype)
        segment.ARBCSLHB1.y_max = torch.tensor(2e-3, dtype=dtype)
        segment.ARBCSLHS1.shape = "elliptical"
        segment.ARBCSLHS1.x_max = torch.tensor(4e-3, dtype=dtype)
        segment.ARBCSLHS1.y_max = torch.tensor(3e-3, dtype=dtype)
        # fat beam (transverse x50) so that the apertures cut through the core
        fat = beam.clone()
        fat.particles[..., :4] *= 50.0
        out = segment.track(fat)
        arrays.update(beam_arrays(f"vectorised.{tag}", out, rows))
        print(
            "ares vectorised survival fractions",
            tag,
            out.survival_probabilities.mean(dim=-1),
        )

    lattice_io.dump(description, OUT / "ares_lattice.json")
    settings = vectorised_settings(
        cheetah.Segment.from_lattice_json(str(ARES_JSON), dtype=torch.float64), 6
    )
    for name, (attr, value) in settings.items():
        arrays[f"vectorised.settings.{name}.{attr}"] = np64(value)
    arrays["vectorised.transverse_scale"] = np.array(50.0)
    np.savez_compressed(OUT / "ares.npz", **arrays)
    print("ares: done,", len(description), "elements")


# --------------------------------------------------------------------------------------
# 3. apertures, cloud-in-cell and space charge through the reference itself
# --------------------------------------------------------------------------------------
def make_aperture() -> None:
    torch.manual_seed(3)
    particles = torch.randn(3, 1000, 7, dtype=torch.float64) * 1e-3
    particles[..., 6] = 1.0
    arrays = {"particles": np64(particles)}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        beam = cheetah.ParticleBeam(
            particles.to(dtype), energy=torch.tensor(1e8, dtype=dtype), dtype=dtype
        )
        for shape in ("rectangular", "elliptical"):
            aperture = cheetah.Aperture(
                x_max=torch.tensor([[1e-3], [5e-4]], dtype=dtype),
                y_max=torch.tensor(8e-4, dtype=dtype),
                shape=shape,
                dtype=dtype,
            )
            out = aperture.track(beam)
            arrays[f"{shape}.{tag}.survival"] = np64(out.survival_probabilities)
    np.savez_compressed(OUT / "aperture.npz", **arrays)
    print("aperture: done")


def make_cloud_in_cell() -> None:
    from cheetah.utils.cloud_in_cell import cloud_in_cell_charge_deposition

    torch.manual_seed(4)
    positions = torch.randn(2, 2000, 3, dtype=torch.float64)
    charges = torch.rand(2, 2000, dtype=torch.float64)
    extent = torch.tensor(
        [[[-2.0, 2.0], [-1.5, 2.5], [-3.0, 3.0]], [[-1.0, 1.0], [-2.0, 2.0], [-0.5, 0.5]]],
        dtype=torch.float64,
    )
    bins = (8, 6, 5)
    arrays = {
        "positions": np64(positions),
        "charges": np64(charges),
        "extent": np64(extent),
        "bins": np.array(bins),
    }
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        grid = cloud_in_cell_charge_deposition(
            positions.to(dtype), bins, extent.to(dtype), charges.to(dtype)
        )
        arrays[f"grid.{tag}"] = np64(grid)
        for dims in (1, 2):  # the 1-D and 2-D specialisations (cloud_in_cell.py:67-241)
            grid = cloud_in_cell_charge_deposition(
                positions[..., :dims].to(dtype), bins[:dims], extent[..., :dims, :].to(dtype),
                charges.to(dtype),
            )
            arrays[f"grid{dims}d.{tag}"] = np64(grid)
    np.savez_compressed(OUT / "cloud_in_cell.npz", **arrays)
    print("cloud_in_cell: done")


def make_space_charge() -> None:
    arrays = {}
    rows = slice(None, None, ROW_STRIDE)
    num_particles = 6000
    torch.manual_seed(5)
    base = cheetah.ParticleBeam.from_parameters(
        num_particles=num_particles,
        total_charge=torch.tensor(1e-9),
        energy=torch.tensor(1e8),
        dtype=torch.float64,
    )
    arrays.update(beam_arrays("incoming", base))

    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        # single kick, 16^3 grid (exposes every stage of the solver at small size)
        beam = base.to(dtype)
        kick = cheetah.SpaceChargeKick(
            effect_length=torch.tensor(0.7, dtype=dtype), grid_shape=(16, 16, 16), dtype=dtype
        )
        # intermediate stages, for unit-level parity of each kernel
        vector = cheetah.ParticleBeam(
            particles=beam.particles.unsqueeze(0),
            energy=beam.energy.unsqueeze(0),
            particle_charges=beam.particle_charges.unsqueeze(0),
            survival_probabilities=beam.survival_probabilities.unsqueeze(0),
            species=beam.species,
            dtype=dtype,
        )
        grid_dimensions = torch.stack(
            [3.0 * vector.sigma_x, 3.0 * vector.sigma_y, 3.0 * vector.sigma_tau], dim=-1
        )
        cell_size = 2 * grid_dimensions / torch.tensor(kick.grid_shape, dtype=dtype)
        xp = vector.to_xyz_pxpypz()
        arrays[f"kick16.{tag}.grid_dimensions"] = np64(grid_dimensions)
        arrays[f"kick16.{tag}.rho_padded"] = np64(
            kick._array_rho(vector, xp, cell_size, grid_dimensions)
        )
        arrays[f"kick16.{tag}.green"] = np64(kick._integrated_green_function(vector, cell_size))
        arrays[f"kick16.{tag}.potential"] = np64(
            kick._solve_poisson_equation(vector, xp, cell_size, grid_dimensions)
        )
        arrays[f"kick16.{tag}.forces"] = np64(
            kick._compute_forces(vector, xp, cell_size, grid_dimensions)[..., rows, :]
        )
        out = kick.track(beam)
        arrays.update(beam_arrays(f"kick16.{tag}", out, rows))

        # vectorised charges (2 beams), default 32^3 grid, low energy (strong kick)
        beam2 = cheetah.ParticleBeam(
            particles=base.particles.to(dtype),
            energy=torch.tensor(5e6, dtype=dtype),
            particle_charges=(
                base.particle_charges.to(dtype) * torch.tensor([[1.0], [3.0]], dtype=dtype)
            ),
            dtype=dtype,
        )
        kick32 = cheetah.SpaceChargeKick(effect_length=torch.tensor(0.2, dtype=dtype), dtype=dtype)
        out = kick32.track(beam2)
        arrays.update(beam_arrays(f"kick32vec.{tag}", out, rows))

        # FODO cell with the split-drift pattern (tests/test_space_charge_kick.py:56-66)
        # + an aperture so that survival weighting matters
        def drift_with_kick(length):
            return [
                cheetah.Drift(length=torch.tensor(length / 2, dtype=dtype), dtype=dtype),
                cheetah.SpaceChargeKick(
                    effect_length=torch.tensor(length, dtype=dtype),
                    grid_shape=(16, 16, 16),
                    dtype=dtype,
                ),
                cheetah.Drift(length=torch.tensor(length / 2, dtype=dtype), dtype=dtype),
            ]

        segment = cheetah.Segment(
            elements=[
                cheetah.Quadrupole(
                    length=torch.tensor(0.2, dtype=dtype),
                    k1=torch.tensor(4.2, dtype=dtype),
                    dtype=dtype,
                ),
                *drift_with_kick(1.0),
                cheetah.Aperture(
                    x_max=torch.tensor(1.5e-4, dtype=dtype),
                    y_max=torch.tensor(5e-4, dtype=dtype),
                    dtype=dtype,
                ),
                cheetah.Quadrupole(
                    length=torch.tensor(0.2, dtype=dtype),
                    k1=torch.tensor(-4.2, dtype=dtype),
                    dtype=dtype,
                ),
                *drift_with_kick(1.0),
            ]
        )
        if tag == "f64":
            lattice_io.dump(
                lattice_io.describe(segment)["elements"], OUT / "fodo_space_charge_lattice.json"
            )
        low_energy = cheetah.ParticleBeam(
            particles=base.particles.to(dtype),
            energy=torch.tensor(5e7, dtype=dtype),
            particle_charges=base.particle_charges.to(dtype) * 0.1,
            dtype=dtype,
        )
        out = segment.track(low_energy)
        arrays.update(beam_arrays(f"fodo.{tag}", out, rows))
        print("fodo survival", tag, out.survival_probabilities.mean().item())

    np.savez_compressed(OUT / "space_charge.npz", **arrays)
    print("space_charge: done")


def make_cavity() -> None:
    """Active cavities: standing / traveling wave, accelerating and
    decelerating, vectorised voltage, and inside a small segment (energy changes mid-lattice)."""
    arrays = {}
    rows = slice(None, None, ROW_STRIDE)
    torch.manual_seed(6)
    base = cheetah.ParticleBeam.from_parameters(
        num_particles=4000, energy=torch.tensor(6e6), sigma_tau=torch.tensor(3e-4),
        dtype=torch.float64,
    )
    arrays.update(beam_arrays("incoming", base))
    parameter_base = base.as_parameter_beam()
    arrays["incoming.mu"] = np64(parameter_base.mu)
    arrays["incoming.cov"] = np64(parameter_base.cov)
    cases = {
        "standing": dict(length=1.0377, voltage=2.0e7, phase=-20.0, frequency=1.3e9,
                         cavity_type="standing_wave"),
        "traveling": dict(length=4.139, voltage=3.0e7, phase=15.0, frequency=2.998e9,
                          cavity_type="traveling_wave"),
        "decelerating": dict(length=1.0, voltage=-2.0e6, phase=0.0, frequency=1.3e9,
                             cavity_type="standing_wave"),
        "vectorised": dict(length=1.0377, voltage=[1.0e7, 2.0e7, 3.0e7], phase=[-10.0, 0.0, 30.0],
                           frequency=1.3e9, cavity_type="standing_wave"),
    }
    descriptions = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        beam = base.to(dtype)
        for name, kw in cases.items():
            kwargs = {k: (v if isinstance(v, str) else torch.tensor(v, dtype=dtype))
                      for k, v in kw.items()}
            cavity = cheetah.Cavity(name=name, dtype=dtype, **kwargs)
            out = cavity.track(beam)
            arrays.update(beam_arrays(f"{name}.{tag}", out, rows))
            parameter_out = cavity.track(parameter_base.to(dtype))  # ParameterBeam branch
            arrays[f"{name}.{tag}.mu"] = np64(parameter_out.mu)
            arrays[f"{name}.{tag}.cov"] = np64(parameter_out.cov)
            arrays[f"{name}.{tag}.parameter_energy"] = np64(parameter_out.energy)
            if tag == "f64":
                descriptions[name] = lattice_io._to_json([lattice_io.describe(cavity)])
        t = lambda v: torch.tensor(v, dtype=dtype)  # noqa: E731
        segment = cheetah.Segment([
            cheetah.Drift(length=t(0.3), dtype=dtype),
            cheetah.Cavity(length=t(1.0377), voltage=t(1.5e7), phase=t(-5.0), frequency=t(1.3e9),
                           name="c1", dtype=dtype),
            cheetah.Quadrupole(length=t(0.2), k1=t(3.0), dtype=dtype),
            cheetah.Aperture(x_max=t(1.5e-4), y_max=t(2e-4), dtype=dtype),
            cheetah.Drift(length=t(0.5), dtype=dtype),
            cheetah.Cavity(length=t(1.0377), voltage=t(2.5e7), phase=t(10.0), frequency=t(1.3e9),
                           name="c2", dtype=dtype),
            cheetah.Drift(length=t(0.4), dtype=dtype),
        ])
        out = segment.track(beam)
        arrays.update(beam_arrays(f"segment.{tag}", out, rows))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # aperture on a ParameterBeam
            parameter_out = segment.track(parameter_base.to(dtype))
        arrays[f"segment.{tag}.mu"] = np64(parameter_out.mu)
        arrays[f"segment.{tag}.cov"] = np64(parameter_out.cov)
        arrays[f"segment.{tag}.parameter_energy"] = np64(parameter_out.energy)
        if tag == "f64":
            descriptions["segment"] = lattice_io._to_json(lattice_io.describe(segment)["elements"])
            print("cavity segment: energy", out.energy.item(), "survival",
                  out.survival_probabilities.mean().item())
    np.savez_compressed(OUT / "cavity.npz", **arrays)
    with (OUT / "cavity.json").open("w") as f:
        json.dump(descriptions, f, separators=(",", ":"))
    print("cavity: done")


# --------------------------------------------------------------------------------------
# 7. non-linear tracking methods: drift_kick_drift (Bmad-X) and second_order
# --------------------------------------------------------------------------------------
NONLINEAR_CONSISTENCY = {
    # the reference's own golden pickles, tests/conftest.py:26-131
    ("Dipole", "second_order"): {"length": 1.0, "angle": [1.0, -2.0], "tilt": 0.42},
    ("Dipole", "drift_kick_drift"): {"length": 1.0, "angle": [1.0, -2.0], "tilt": 0.42},
    ("Drift", "second_order"): {"length": [1.0, -1.0]},
    ("Drift", "drift_kick_drift"): {"length": [1.0, -1.0]},
    ("Quadrupole", "second_order"): {
        "length": 1.0, "k1": [1.0, -2.0], "tilt": 0.42, "misalignment": [0.01, -0.02]},
    ("Quadrupole", "drift_kick_drift"): {
        "length": 1.0, "k1": [1.0, -2.0], "tilt": 0.42, "misalignment": [0.01, -0.02]},
    ("RBend", "second_order"): {"length": 1.0, "angle": [1.0, -2.0], "tilt": 0.42},
    ("RBend", "drift_kick_drift"): {"length": 1.0, "angle": [1.0, -2.0], "tilt": 0.42},
    ("Sextupole", "second_order"): {
        "length": 1.0, "k2": [1.0, -2.0], "tilt": 0.42, "misalignment": [0.01, -0.02]},
    ("TransverseDeflectingCavity", "inactive"): {"length": 1.0, "voltage": 0.0},
    ("TransverseDeflectingCavity", "active"): {"length": 1.0, "voltage": 1e6},
}


def _tensor_kwargs(kw: dict, dtype) -> dict:
    return {
        k: (torch.tensor(v, dtype=dtype) if isinstance(v, (float, int, list)) and k != "num_steps"
            else v)
        for k, v in kw.items()
    }


def make_nonlinear() -> None:
    resources = REF / "tests" / "resources"
    rows = slice(None, None, ROW_STRIDE)
    arrays, lattices = {}, {}

    # (a) the reference's consistency pickles (incoming beam = consistency.npz "incoming")
    with (resources / "ACHIP_EA1_2021.1351.001_subsampled_3000.pkl").open("rb") as f:
        incoming = pickle.load(f).to(torch.float64)
    for (cls_name, label), kw in NONLINEAR_CONSISTENCY.items():
        key = f"consistency.{cls_name}_{label}"
        kwargs = _tensor_kwargs(kw, torch.float32)
        if cls_name != "TransverseDeflectingCavity":
            kwargs["tracking_method"] = label
        element = getattr(cheetah, cls_name)(name=label, **kwargs).to(torch.float64)
        lattices[key] = lattice_io._to_json([lattice_io.describe(element)])
        with (resources / "consistency_expected_outgoing" /
              f"{cls_name}_ParticleBeam_{label}.pkl").open("rb") as f:
            expected = pickle.load(f)
        arrays.update(beam_arrays(f"{key}.expected", expected, rows))
        actual = element.track(incoming)
        assert torch.allclose(actual.particles, expected.particles), key

    # (b) the Bmad-X fixtures (tests/test_drift.py:43-69, test_quadrupole.py:173-208,
    #     test_dipole.py:108-150, test_transverse_deflecting_cavity.py:10-43)
    bmadx_incoming = torch.load(resources / "bmadx" / "incoming.pt", weights_only=False)
    # particles are independent in these elements: every 4th one is stored, in and out
    arrays.update(beam_arrays("bmadx.incoming", bmadx_incoming, rows))
    angle = 20 * torch.pi / 180
    bmadx_cases = {
        "drift": ("Drift", {"length": 1.0, "tracking_method": "drift_kick_drift"}),
        "quadrupole": ("Quadrupole", {
            "length": 1.0, "k1": 10.0, "misalignment": [0.01, -0.02], "tilt": 0.5,
            "num_steps": 10, "tracking_method": "drift_kick_drift"}),
        "dipole": ("Dipole", {
            "length": 0.5, "angle": angle, "dipole_e1": angle / 2, "dipole_e2": angle - angle / 2,
            "tilt": 0.1, "fringe_integral": 0.5, "fringe_integral_exit": 0.5, "gap": 0.05,
            "gap_exit": 0.05, "fringe_at": "both", "fringe_type": "linear_edge",
            "tracking_method": "drift_kick_drift"}),
        "transverse_deflecting_cavity": ("TransverseDeflectingCavity", {
            "length": 1.0, "voltage": 1e7, "phase": 0.2, "frequency": 1e9}),
    }
    for name, (cls_name, kw) in bmadx_cases.items():
        element = getattr(cheetah, cls_name)(
            name=name, dtype=torch.float64, **_tensor_kwargs(kw, torch.float64))
        lattices[f"bmadx.{name}"] = lattice_io._to_json([lattice_io.describe(element)])
        outgoing = torch.load(resources / "bmadx" / f"outgoing_{name}.pt", weights_only=False)
        arrays[f"bmadx.{name}.expected.particles"] = np64(outgoing[0, rows])
        actual = element.track(bmadx_incoming)
        assert torch.allclose(actual.particles, outgoing, atol=1e-14, rtol=1e-14), name

    # (c) fresh reference outputs: vectorised settings, edge cases, other species, a mixed lattice
    torch.manual_seed(11)
    base = cheetah.ParticleBeam.from_parameters(
        num_particles=2000, energy=torch.tensor(5e7), sigma_x=torch.tensor(4e-4),
        sigma_y=torch.tensor(3e-4), sigma_px=torch.tensor(2e-4), sigma_py=torch.tensor(1e-4),
        sigma_tau=torch.tensor(5e-4), sigma_p=torch.tensor(2e-3),
        mu_x=torch.tensor(1e-4), mu_py=torch.tensor(-5e-5), dtype=torch.float64,
    )
    proton = cheetah.ParticleBeam.from_parameters(
        num_particles=2000, energy=torch.tensor(1.2e9), sigma_tau=torch.tensor(1e-3),
        sigma_p=torch.tensor(3e-3), species=cheetah.Species("proton"), dtype=torch.float64,
    )
    arrays.update(beam_arrays("fresh.incoming", base))
    arrays.update(beam_arrays("fresh.proton", proton))
    fresh = {
        "drift_dkd_vector": ("Drift", {"length": [0.3, 1.7, -0.4],
                                       "tracking_method": "drift_kick_drift"}, "incoming"),
        "drift_dkd_proton": ("Drift", {"length": 2.5, "tracking_method": "drift_kick_drift"},
                             "proton"),
        "quadrupole_dkd_steps": ("Quadrupole", {
            "length": 0.4, "k1": [4.2, -7.1, 0.0], "tilt": [0.0, 0.3, -0.2],
            "misalignment": [[0.0, 0.0], [1e-4, -2e-4], [3e-4, 0.0]], "num_steps": 5,
            "tracking_method": "drift_kick_drift"}, "incoming"),
        "quadrupole_dkd_proton": ("Quadrupole", {
            "length": 0.6, "k1": 2.0, "num_steps": 3, "tracking_method": "drift_kick_drift"},
            "proton"),
        "dipole_dkd_zero_angle": ("Dipole", {
            "length": 0.7, "angle": 0.0, "tracking_method": "drift_kick_drift"}, "incoming"),
        "dipole_dkd_entrance": ("Dipole", {
            "length": 0.8, "angle": [0.2, -0.35], "dipole_e1": 0.05, "dipole_e2": -0.03,
            "fringe_integral": 0.4, "gap": 0.03, "gap_exit": 0.05, "fringe_integral_exit": 0.3,
            "fringe_at": "entrance", "tracking_method": "drift_kick_drift"}, "incoming"),
        "dipole_dkd_neither": ("Dipole", {
            "length": 0.8, "angle": 0.25, "dipole_e1": 0.05, "tilt": 0.3,
            "fringe_at": "neither", "tracking_method": "drift_kick_drift"}, "incoming"),
        "rbend_dkd_exit": ("RBend", {
            "length": 0.5, "angle": 0.3, "rbend_e1": 0.02, "rbend_e2": -0.01,
            "fringe_integral": 0.5, "gap": 0.02, "fringe_at": "exit",
            "tracking_method": "drift_kick_drift"}, "incoming"),
        "tdc_vector": ("TransverseDeflectingCavity", {
            "length": 0.6, "voltage": [2e6, -5e6], "phase": [0.1, 0.35], "frequency": 2.856e9,
            "tilt": 0.2, "misalignment": [1e-4, -1e-4]}, "incoming"),
        "tdc_proton": ("TransverseDeflectingCavity", {
            "length": 0.6, "voltage": 3e7, "phase": 0.05, "frequency": 4e8}, "proton"),
        "drift_second_order_proton": ("Drift", {"length": 1.3, "tracking_method": "second_order"},
                                      "proton"),
        "quadrupole_second_order": ("Quadrupole", {
            "length": 0.3, "k1": [5.0, -12.0, 0.0], "tilt": [0.0, 0.785, 0.1],
            "tracking_method": "second_order"}, "incoming"),
        "sextupole_second_order": ("Sextupole", {
            "length": 0.25, "k2": [30.0, -80.0], "misalignment": [2e-4, 1e-4],
            "tracking_method": "second_order"}, "incoming"),
        "dipole_second_order_k1": ("Dipole", {
            "length": 0.9, "angle": [0.3, -0.1], "k1": [1.5, -0.8], "dipole_e1": 0.1,
            "dipole_e2": 0.05, "fringe_integral": 0.5, "gap": 0.03, "tilt": 0.15,
            "tracking_method": "second_order"}, "incoming"),
        "dipole_second_order_kx2_zero": ("Dipole", {
            "length": 0.5, "angle": 0.0, "k1": 0.0, "tracking_method": "second_order"},
            "incoming"),
    }
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        beams = {"incoming": base.to(dtype), "proton": proton.to(dtype)}
        for name, (cls_name, kw, which) in fresh.items():
            element = getattr(cheetah, cls_name)(
                name=name, dtype=dtype, **_tensor_kwargs(kw, dtype))
            out = element.track(beams[which])
            arrays.update(beam_arrays(f"fresh.{name}.{tag}", out, rows))
            if tag == "f64":
                lattices[f"fresh.{name}"] = {
                    "beam": which,
                    "lattice": lattice_io._to_json([lattice_io.describe(element)]),
                }
        # a mixed lattice: linear runs, markers, dkd and second-order elements, an aperture
        t = lambda v: torch.tensor(v, dtype=dtype)  # noqa: E731
        segment = cheetah.Segment([
            cheetah.Drift(length=t(0.4), tracking_method="drift_kick_drift", dtype=dtype),
            cheetah.Marker(name="m1"),
            cheetah.Quadrupole(length=t(0.2), k1=t([3.0, -3.0]), num_steps=4,
                               tracking_method="drift_kick_drift", dtype=dtype),
            cheetah.Drift(length=t(0.3), dtype=dtype),
            cheetah.HorizontalCorrector(length=t(0.1), angle=t(1e-4), dtype=dtype),
            cheetah.Dipole(length=t(0.5), angle=t(0.15), dipole_e1=t(0.07), dipole_e2=t(0.08),
                           fringe_integral=t(0.5), gap=t(0.02),
                           tracking_method="drift_kick_drift", dtype=dtype),
            cheetah.BPM(name="b1"),
            cheetah.Sextupole(length=t(0.15), k2=t(40.0), tracking_method="second_order",
                              dtype=dtype),
            cheetah.Aperture(x_max=t(8e-4), y_max=t(6e-4), dtype=dtype),
            cheetah.Quadrupole(length=t(0.2), k1=t(-2.5), tracking_method="second_order",
                               dtype=dtype),
            cheetah.Drift(length=t(0.6), tracking_method="second_order", dtype=dtype),
            cheetah.TransverseDeflectingCavity(length=t(0.3), voltage=t(1e6), phase=t(0.1),
                                               frequency=t(1.3e9), dtype=dtype),
            cheetah.Drift(length=t(0.25), dtype=dtype),
        ])
        out = segment.track(beams["incoming"])
        arrays.update(beam_arrays(f"fresh.segment.{tag}", out, rows))
        if tag == "f64":
            lattices["fresh.segment"] = {
                "beam": "incoming",
                "lattice": lattice_io._to_json(lattice_io.describe(segment)["elements"]),
            }
            print("nonlinear segment: survival", out.survival_probabilities.mean().item(),
                  "shape", tuple(out.particles.shape))
    np.savez_compressed(OUT / "nonlinear.npz", **arrays)
    with (OUT / "nonlinear.json").open("w") as f:
        json.dump({"row_stride": ROW_STRIDE, "lattices": lattices}, f, separators=(",", ":"))
    print("nonlinear:", len(NONLINEAR_CONSISTENCY), "+", len(bmadx_cases), "+", len(fresh) + 1,
          "cases")


# --------------------------------------------------------------------------------------
# 8. diagnostics: active Screen images and BPM readings
# --------------------------------------------------------------------------------------
def make_diagnostics() -> None:
    arrays, meta = {}, {}
    torch.manual_seed(21)
    base = cheetah.ParticleBeam.from_parameters(
        num_particles=6_000, sigma_x=torch.tensor(3e-4), sigma_y=torch.tensor(2e-4),
        mu_x=torch.tensor(1e-4), mu_y=torch.tensor(-5e-5), total_charge=torch.tensor(1e-10),
        dtype=torch.float64,
    )
    survival = (torch.rand(6_000, dtype=torch.float64) > 0.2).to(torch.float64) * torch.rand(
        6_000, dtype=torch.float64)
    base.survival_probabilities = survival
    arrays.update(beam_arrays("incoming", base))
    screens = {
        "cic": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], method="cloud-in-cell"),
        "cic_binned_misaligned": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], binning=2,
                                      misalignment=[2e-4, -1e-4], method="cloud-in-cell"),
        "cic_clipping": dict(resolution=(40, 30), pixel_size=[1e-5, 1e-5], method="cloud-in-cell"),
        "histogram": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], method="histogram"),
        "histogram_binned": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], binning=4,
                                 misalignment=[2e-4, -1e-4], method="histogram"),
        "kde": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], method="kde"),
        "kde_binned_bandwidth": dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5], binning=2,
                                     misalignment=[2e-4, -1e-4], method="kde",
                                     kde_bandwidth=8e-5),
        "kde_clipping": dict(resolution=(40, 30), pixel_size=[1e-5, 1e-5], method="kde"),
    }
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        beam = base.to(dtype)
        for name, kw in screens.items():
            kwargs = {k: (torch.tensor(v, dtype=dtype) if isinstance(v, (list, float)) else v)
                      for k, v in kw.items()}
            screen = cheetah.Screen(is_active=True, name=name, dtype=dtype, **kwargs)
            out = screen.track(beam)
            assert torch.equal(out.particles, beam.particles)
            arrays[f"screen.{name}.{tag}"] = np64(screen.reading)
            meta[f"screen.{name}"] = kw
        # vectorised beams (3 transverse offsets) on the cloud-in-cell screen
        segment = cheetah.Segment([
            cheetah.HorizontalCorrector(length=torch.tensor(0.1, dtype=dtype),
                                        angle=torch.tensor([0.0, 1e-3, -2e-3], dtype=dtype)),
            cheetah.Drift(length=torch.tensor(0.5, dtype=dtype)),
            cheetah.BPM(is_active=True, name="bpm", misalignment=torch.tensor([1e-4, 2e-4], dtype=dtype)),
            cheetah.Screen(is_active=True, name="screen", resolution=(96, 64),
                           pixel_size=torch.tensor([2.5e-5, 3e-5], dtype=dtype), dtype=dtype),
        ])
        out = segment.track(beam)
        kde_screen = cheetah.Screen(is_active=True, name="kde_screen", resolution=(96, 64),
                                    pixel_size=torch.tensor([2.5e-5, 3e-5], dtype=dtype),
                                    method="kde", dtype=dtype)
        kde_screen.track(out)  # vectorised read beam (3 settings)
        arrays[f"segment.kde_screen.{tag}"] = np64(kde_screen.reading)
        arrays[f"segment.screen.{tag}"] = np64(segment.screen.reading)
        arrays[f"segment.bpm.{tag}"] = np64(segment.bpm.reading)
        arrays[f"segment.outgoing_shape.{tag}"] = np.asarray(out.particles.shape)
        bpm = cheetah.BPM(is_active=True, misalignment=torch.tensor([0.1, 0.2], dtype=dtype))
        bpm.track(beam)
        arrays[f"bpm.{tag}"] = np64(bpm.reading)
        # ParameterBeam on a screen: analytic bivariate normal on the pixel grid
        parameter_beam = beam.as_parameter_beam()
        for name, kw in (("gaussian", dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5])),
                         ("gaussian_binned", dict(resolution=(96, 64), pixel_size=[2.5e-5, 3e-5],
                                                  binning=2, misalignment=[2e-4, -1e-4]))):
            kwargs = {k: (torch.tensor(v, dtype=dtype) if isinstance(v, list) else v)
                      for k, v in kw.items()}
            screen = cheetah.Screen(is_active=True, name=name, dtype=dtype, **kwargs)
            screen.track(parameter_beam)
            arrays[f"screen.{name}.{tag}"] = np64(screen.reading)
            meta[f"screen.{name}"] = kw
        arrays[f"parameter_beam.mu.{tag}"] = np64(parameter_beam.mu)
        arrays[f"parameter_beam.cov.{tag}"] = np64(parameter_beam.cov)
        blocking = cheetah.Screen(is_active=True, is_blocking=True, dtype=dtype)
        arrays[f"blocking.survival.{tag}"] = np64(blocking.track(beam).survival_probabilities)
    np.savez_compressed(OUT / "diagnostics.npz", **arrays)
    with (OUT / "diagnostics.json").open("w") as f:
        json.dump(meta, f, separators=(",", ":"))
    print("diagnostics:", len(screens), "screens, sum", float(arrays["screen.cic.f64"].sum()))


SCALAR_PROPERTIES = (
    [f"mu_{c}" for c in ("x", "px", "y", "py", "tau", "p")]
    + [f"sigma_{c}" for c in ("x", "px", "y", "py", "tau", "p")]
    + ["cov_xpx", "cov_ypy", "cov_taup", "cov_xp", "cov_pxp", "cov_yp", "cov_pyp", "cov_xy",
       "cov_xpy", "cov_xtau", "cov_pxy", "cov_pxpy", "cov_pxtau", "cov_ytau", "cov_pytau",
       "emittance_x", "emittance_y", "projected_emittance_x", "projected_emittance_y",
       "normalized_emittance_x", "normalized_emittance_y", "beta_x", "beta_y", "alpha_x",
       "alpha_y", "dispersion_x", "dispersion_px", "dispersion_y", "dispersion_py",
       "relativistic_gamma", "relativistic_beta", "p0c", "total_charge"]
)


def make_beam_properties() -> None:
    """Derived beam quantities of the reference (beam.py:262-557, particle_beam.py:1034-1346,
    :1699-1951, parameter_beam.py:62-760) for the host-side beam mirror."""
    arrays = {}
    torch.manual_seed(33)
    mixing = torch.eye(6, dtype=torch.float64) + 0.3 * torch.randn(6, 6, dtype=torch.float64)
    sigma = torch.tensor([3e-4, 4e-5, 2e-4, 3e-5, 1e-4, 2e-3], dtype=torch.float64)
    phase_space = (torch.randn(4000, 6, dtype=torch.float64) @ mixing.T) * sigma + sigma * 0.3
    particles = torch.cat([phase_space, torch.ones(4000, 1, dtype=torch.float64)], dim=-1)
    survival = (torch.rand(4000, dtype=torch.float64) > 0.15) * torch.rand(4000, dtype=torch.float64)
    charges = torch.rand(4000, dtype=torch.float64) * 1e-14
    beam = cheetah.ParticleBeam(particles, torch.tensor(6.3e7, dtype=torch.float64),
                                particle_charges=charges, survival_probabilities=survival,
                                dtype=torch.float64)
    arrays.update(beam_arrays("incoming", beam))
    for name in SCALAR_PROPERTIES:
        arrays[f"particle.{name}"] = np64(getattr(beam, name))
    arrays["particle.energies"] = np64(beam.energies)
    arrays["particle.momenta"] = np64(beam.momenta)
    arrays["particle.xyz_pxpypz"] = np64(beam.to_xyz_pxpypz())
    round_trip = cheetah.ParticleBeam.from_xyz_pxpypz(beam.to_xyz_pxpypz(), beam.energy,
                                                      dtype=torch.float64)
    arrays["particle.round_trip"] = np64(round_trip.particles)
    moved = beam.transformed_to(mu_x=torch.tensor(1e-3, dtype=torch.float64),
                                sigma_py=torch.tensor(5e-5, dtype=torch.float64),
                                total_charge=torch.tensor(3e-11, dtype=torch.float64))
    arrays["particle.transformed.particles"] = np64(moved.particles)
    arrays["particle.transformed.charges"] = np64(moved.particle_charges)
    parameter = beam.as_parameter_beam()
    arrays["parameter.mu"] = np64(parameter.mu)
    arrays["parameter.cov"] = np64(parameter.cov)
    for name in SCALAR_PROPERTIES:
        arrays[f"parameter.{name}"] = np64(getattr(parameter, name))
    twiss = cheetah.ParameterBeam.from_twiss(
        beta_x=torch.tensor([1.0, 2.5], dtype=torch.float64),
        alpha_x=torch.tensor(-0.7, dtype=torch.float64),
        emittance_x=torch.tensor(3e-9, dtype=torch.float64),
        beta_y=torch.tensor(4.0, dtype=torch.float64),
        alpha_y=torch.tensor([0.2, 0.0], dtype=torch.float64),
        emittance_y=torch.tensor(2e-9, dtype=torch.float64),
        sigma_tau=torch.tensor(1e-4, dtype=torch.float64),
        sigma_p=torch.tensor(1e-3, dtype=torch.float64),
        cov_taup=torch.tensor(2e-8, dtype=torch.float64),
        dispersion_x=torch.tensor(0.03, dtype=torch.float64),
        dispersion_py=torch.tensor(-0.02, dtype=torch.float64),
        energy=torch.tensor(1.2e8, dtype=torch.float64), dtype=torch.float64,
    )
    arrays["twiss.mu"] = np64(twiss.mu)
    arrays["twiss.cov"] = np64(twiss.cov)
    for name in SCALAR_PROPERTIES:
        arrays[f"twiss.{name}"] = np64(getattr(twiss, name))
    changed = twiss.transformed_to(sigma_x=torch.tensor(2e-4, dtype=torch.float64),
                                   mu_y=torch.tensor(1e-4, dtype=torch.float64))
    arrays["twiss.transformed.mu"] = np64(changed.mu)
    arrays["twiss.transformed.cov"] = np64(changed.cov)
    np.savez_compressed(OUT / "beam_properties.npz", **arrays)
    print("beam properties:", len(arrays), "arrays, emittance_x", float(arrays["particle.emittance_x"]))


def make_second_order_maps() -> None:
    """Dense second-order transfer maps of the reference (element.py:134-147 and the per-class
    assemblies) for Element.second_order_transfer_map."""
    arrays = {}
    t = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    energy = t([6.3e7, 1.2e8])
    species = cheetah.Species("electron", dtype=torch.float64)
    elements = {
        "drift": cheetah.Drift(length=t(0.7), dtype=torch.float64),
        "quadrupole": cheetah.Quadrupole(length=t(0.2), k1=t([4.2, -3.1]), tilt=t(0.3),
                                         misalignment=t([2e-4, -1e-4]), dtype=torch.float64),
        "sextupole": cheetah.Sextupole(length=t(0.15), k2=t(25.0), tilt=t(-0.2),
                                       misalignment=t([1e-4, 3e-4]), dtype=torch.float64),
        "dipole": cheetah.Dipole(length=t(0.5), angle=t(0.2), k1=t(0.4), dipole_e1=t(0.05),
                                 dipole_e2=t(0.08), tilt=t(0.1), fringe_integral=t(0.5),
                                 fringe_integral_exit=t(0.4), gap=t(0.03), dtype=torch.float64),
        "rbend": cheetah.RBend(length=t(0.4), angle=t(-0.15), fringe_integral=t(0.3), gap=t(0.02),
                               dtype=torch.float64),
    }
    for name, element in elements.items():
        arrays[name] = np64(element.second_order_transfer_map(energy, species))
    np.savez_compressed(OUT / "second_order_maps.npz", **arrays)
    print("second-order maps:", {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    if "--only-cic" in sys.argv:
        make_cloud_in_cell()
        sys.exit(0)
    if "--only-consistency" in sys.argv:
        make_consistency()
        sys.exit(0)
    if "--only-diagnostics" in sys.argv:
        make_diagnostics()
        sys.exit(0)
    if "--only-cavity" in sys.argv:
        make_cavity()
        sys.exit(0)
    if "--only-nonlinear" in sys.argv:
        make_nonlinear()
        sys.exit(0)
    if "--only-second-order-maps" in sys.argv:
        make_second_order_maps()
        sys.exit(0)
    if "--only-beam-properties" in sys.argv:
        make_beam_properties()
        sys.exit(0)
    make_consistency()
    make_ares()
    make_aperture()
    make_cloud_in_cell()
    make_space_charge()
    make_cavity()
    make_nonlinear()
    make_diagnostics()
    make_beam_properties()
    make_second_order_maps()
    for path in sorted(OUT.iterdir()):
        print(f"{path.name:40s} {path.stat().st_size / 1024:8.1f} KiB")
