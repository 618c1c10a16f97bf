"""Secondary benchmark: the per-particle non-linear tracking methods.
Same JSON schema as bench.py (which keeps the driver contract for the headline workload).

    python bench_nonlinear.py [--steps K] [--warmup W] [--particles N] [--settings B]

Cases (float32, one B200):
  * single elements with a PER-SETTING beam (B, N, 7): 28 B read + 28 B written per (particle,
    setting) -- these are HBM-bound and carry the roofline figure;
  * a 20-element drift_kick_drift FODO line and a 20-element second_order line as ONE fused run
    with a shared beam: 28 B written per (particle, setting) for all 20 elements -- bound by
    instruction issue, not by HBM (the reference makes 20 x ~100 elementwise passes).
`value` is the fused drift_kick_drift line in particle-steps/s (settings x particles x elements).
"""

from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import workloads  # noqa: E402


def fodo(method: str, cells: int, settings: int, dtype) -> list:
    """[Drift, Quadrupole(+k), Drift, Quadrupole(-k)] x cells with per-setting k1 (seed 5)."""
    g = torch.Generator().manual_seed(5)
    description = []
    for cell in range(cells):
        for sign in (1.0, -1.0):
            k1 = sign * (4.0 + torch.rand(settings, generator=g, dtype=torch.float64))
            description.append({"type": "Drift", "name": f"d{cell}{sign}",
                                "length": torch.tensor(0.5, dtype=dtype),
                                "tracking_method": method})
            quad = {"type": "Quadrupole", "name": f"q{cell}{sign}",
                    "length": torch.tensor(0.2, dtype=dtype), "k1": k1.to(dtype),
                    "tracking_method": method}
            if method == "drift_kick_drift":
                quad["num_steps"] = 5
            description.append(quad)
    return description


def main() -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--particles", type=int, default=1_000_000)
    p.add_argument("--settings", type=int, default=64)
    p.add_argument("--no-cpu-baseline", action="store_true")
    args = p.parse_args()

    import cheetah_b200 as cb
    from cheetah_b200 import _capi
    from cheetah_b200 import lattice_description
    from oracle import track_oracle as oracle

    device, dtype = torch.device("cuda", 0), torch.float32
    n, B = args.particles, args.settings
    particles = workloads.parameters_beam_particles(n).to(dtype)
    species = cb.Species("electron", device=device, dtype=dtype)
    energy = torch.tensor(1e8, device=device, dtype=dtype)
    shared = cb.ParticleBeam(particles.to(device), energy, species=species)
    shared._unit_seventh = True
    batched = cb.ParticleBeam(particles.to(device).expand(B, n, 7).contiguous(), energy,
                              species=species)
    batched._unit_seventh = True

    peak, peak_kind = 6450.0, "fallback"
    peaks = REPO / "MEASURED_PEAKS.json"
    if peaks.exists():
        peak, peak_kind = float(json.loads(peaks.read_text())["hbm_gbs"]), "measured"

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _capi.launch_count()
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.steps, (_capi.launch_count() - before) // args.steps

    t = lambda v: torch.tensor(v, device=device, dtype=dtype)  # noqa: E731
    singles = {
        "Drift drift_kick_drift": cb.Drift(length=t(1.0), tracking_method="drift_kick_drift"),
        "Quadrupole drift_kick_drift (5 steps)": cb.Quadrupole(
            length=t(0.2), k1=t(4.2), num_steps=5, tracking_method="drift_kick_drift"),
        "Dipole drift_kick_drift (fp64 body)": cb.Dipole(
            length=t(0.5), angle=t(0.2), dipole_e1=t(0.1), dipole_e2=t(0.1),
            fringe_integral=t(0.5), gap=t(0.03), tracking_method="drift_kick_drift"),
        "TransverseDeflectingCavity (fp64 kick)": cb.TransverseDeflectingCavity(
            length=t(0.5), voltage=t(1e6), phase=t(0.1), frequency=t(2.856e9)),
        "Quadrupole second_order": cb.Quadrupole(length=t(0.2), k1=t(4.2),
                                                 tracking_method="second_order"),
        "Sextupole second_order": cb.Sextupole(length=t(0.2), k2=t(30.0)),
    }
    kernels = []
    for name, element in singles.items():
        segment = cb.Segment([element])
        ms, launches = timed(lambda: segment.track(batched))
        nbytes = B * n * 56
        kernels.append({
            "case": f"{name}, per-setting beam ({B} x {n})", "ms": ms, "launches": launches,
            "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
            "particle_steps_per_s": B * n / (ms * 1e-3),
        })

    fused = {}
    for method in ("drift_kick_drift", "second_order"):
        description = fodo(method, 5, B, dtype)
        segment = cb.Segment(lattice_description.build(description, device=device, dtype=dtype))
        ms, launches = timed(lambda: segment.track(shared))
        fused[method] = {
            "case": f"20-element {method} FODO line, shared beam, {B} settings x {n} particles",
            "ms": ms, "launches": launches, "n_elements": len(description),
            "particle_steps_per_s": B * n * len(description) / (ms * 1e-3),
            "algorithmic_bytes": B * n * 28 + n * 28,
            "achieved_gbs": (B * n * 28 + n * 28) / (ms * 1e-3) / 1e9,
            "bound": "instruction issue (20 elements per HBM round trip)",
        }

    cpu_baseline = None
    if not args.no_cpu_baseline:
        description = fodo("drift_kick_drift", 5, 1, dtype)
        cpu_beam = oracle.make_beam(particles, torch.tensor(1e8, dtype=dtype))
        oracle.track(description, cpu_beam)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            oracle.track(description, cpu_beam)
        per_pass = (time.perf_counter() - t0) / reps
        cpu_baseline = {
            "value": n * len(description) / per_pass, "unit": "particle-steps/s",
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 of {B} settings x {n} particles x {len(description)} drift_kick_drift "
                      f"elements, float32, torch CPU, mean of {reps} passes ({per_pass * 1e3:.0f} ms)",
        }

    dominant = kernels[0]
    headline = fused["drift_kick_drift"]
    line = {
        "metric": "particle-steps/sec (Segment.track, ParticleBeam)",
        "value": headline["particle_steps_per_s"], "unit": "particle-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": headline["ms"],
        "higher_is_better": True, "scaling": "n/a", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": headline["case"] + " (tracking_method='drift_kick_drift')",
                   "particles": n, "settings": B, "n_elements": headline["n_elements"]},
        "roofline": {
            "kernel": "nonlinear_track_kernel<float,2,128> (" + dominant["case"] + ")",
            "bound": "hbm", "achieved": dominant["achieved_gbs"], "peak": peak,
            "peak_kind": peak_kind, "unit": "GB/s", "frac": dominant["frac_of_hbm_peak"],
            "traffic": None,
        },
        "single_elements": kernels,
        "fused_runs": fused,
        "cpu_baseline": cpu_baseline,
        "gpu_launches": headline["launches"] * args.steps,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
