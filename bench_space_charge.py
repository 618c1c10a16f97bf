"""Secondary benchmark: BASELINE configs[3] -- 50 FODO cells (400 elements, 100 space-charge
kicks on a 64^3 grid), 1e6 particles, one B200.  Same JSON schema as bench.py (bench.py keeps
the driver contract for the headline workload; this script documents the second kernel family).

    python bench_space_charge.py [--steps K] [--warmup W] [--cells 50] [--grid 64]

`value` is eager `Segment.track`; `graph_value` replays the same call from a CUDA graph
(cheetah_b200.GraphedTrack).  `kernels` lists, for every kernel of one kick, the mean duration
from CUDA events around back-to-back launches of that stage alone and the achieved bandwidth on
its algorithmic bytes against MEASURED_PEAKS.json hbm_gbs.
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


def main() -> None:
    p = argparse.ArgumentParser()
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--cells", type=int, default=50)
    p.add_argument("--grid", type=int, default=64)
    p.add_argument("--particles", type=int, default=1_000_000)
    p.add_argument("--beams", type=int, default=1,
                   help="independent beams tracked together (BASELINE configs[4]: 1024 beams "
                        "over 8 GPUs = 128 per GPU); total charge linspace(1e-11, 1e-9)")
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    args = p.parse_args()

    import cheetah_b200 as cb
    from cheetah_b200 import _capi, space_charge
    from oracle import track_oracle as oracle

    device, dtype = torch.device("cuda", 0), torch.float32
    n, grid = args.particles, args.grid
    description = workloads.fodo_space_charge(args.cells, grid, dtype)
    segment = workloads.product_segment(description, device, dtype)
    particles = workloads.parameters_beam_particles(n)
    charges = torch.full((n,), 1e-10 / n, dtype=dtype)
    B = args.beams
    if B == 1:
        beam_particles, beam_charges = particles.to(device=device, dtype=dtype), charges.to(device)
    else:  # per-beam particles (seed = beam index) and charges
        beam_particles = torch.empty((B, n, 7), device=device, dtype=dtype)
        for i in range(B):
            beam_particles[i] = workloads.parameters_beam_particles(n, seed=i).to(device=device, dtype=dtype)
        total = torch.linspace(1e-11, 1e-9, B, dtype=dtype)
        beam_charges = (total.unsqueeze(-1) / n).expand(B, n).contiguous().to(device)
    beam = cb.ParticleBeam(
        beam_particles, torch.tensor(1e8, device=device, dtype=dtype),
        particle_charges=beam_charges, species=cb.Species("electron", device=device, dtype=dtype),
    )
    beam._unit_seventh = True
    n_elements, kicks = len(description), 2 * args.cells

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        before = _capi.launch_count()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps, _capi.launch_count() - before

    ms, launches = timed(lambda: segment.track(beam), args.steps, args.warmup)
    graph_ms = None
    if not args.no_graph:
        graphed = cb.GraphedTrack(segment, beam)
        graph_ms, _ = timed(lambda: graphed.replay(), args.steps, 1)

    # ---- per-kernel timing of ONE kick, stage by stage -------------------------------------
    lib = _capi.lib()
    out, ws = space_charge.kick(
        beam.particles, beam.energy, beam.particle_charges, beam.survival_probabilities,
        beam.species.mass_eV, torch.tensor(1.0, device=device),
        tuple(torch.tensor(3.0, device=device) for _ in range(3)), (grid, grid, grid),
    )
    torch.cuda.synchronize()
    stream = _capi.current_stream(device)
    code = _capi.CH_F32
    pp, q, w = beam.particles, beam.particle_charges, beam.survival_probabilities
    ps, qs = (0, 0) if B == 1 else (n * 7, n)
    out = out.reshape(B, n, 7)
    one = torch.tensor(1.0, device=device)
    three = torch.tensor(3.0, device=device)
    cells3 = grid ** 3
    spectrum_bytes = (2 * grid) * (2 * grid) * (grid + 1) * 8
    stages = {
        "sc_moments_kernel (+grid params)": (
            lambda: lib.ch_sc_moments_and_params(
                pp.data_ptr(), ps, w.data_ptr(), 0, n, B, beam.energy.data_ptr(), 0, code,
                beam.species.mass_eV.data_ptr(), code, one.data_ptr(), 0, code,
                three.data_ptr(), 0, three.data_ptr(), 0, three.data_ptr(), 0, code,
                grid, grid, grid, code, ws.stats.data_ptr(), ws.params.data_ptr(), stream),
            B * n * 32, "hbm: 28 B row + 4 B survival per particle"),
        "sc_deposit_kernel": (
            lambda: lib.ch_sc_deposit(
                pp.data_ptr(), ps, q.data_ptr(), qs, w.data_ptr(), 0, ws.params.data_ptr(), n, B,
                grid, grid, grid, code, ws.rho_quad.data_ptr(), stream),
            B * n * 36, "hbm: 28 B row + charge + survival per particle (atomics stay in L2)"),
        "green lattice + 3 even FFT passes": (
            lambda: (lib.ch_sc_green_function(ws.params.data_ptr(), B, grid, grid, grid, code,
                                              ws.lattice.data_ptr(), None, stream),
                     lib.ch_sc_green_spectrum(ws.lattice.data_ptr(), B, grid, grid, grid, code,
                                              ws.green_scratch.data_ptr(),
                                              ws.green_spectrum.data_ptr(), stream)),
            B * ((grid + 1) ** 3 * (8 + 8) + 4 * cells3 * 4),
            "lattice write+read, 3 passes (L2-resident for one beam)"),
        "poisson: r2c z, y, fused x conv, inverse y, c2r z": (
            lambda: lib.ch_sc_poisson_solve(
                ws.rho_quad.data_ptr(), ws.green_spectrum.data_ptr(), ws.params.data_ptr(), B,
                grid, grid, grid, code, ws.rho_spectrum.data_ptr(), ws.phi.data_ptr(), stream),
            B * (int(spectrum_bytes * (0.25 + 0.5 + 0.5 + 1.0 + 0.5 + 0.5 + 0.25)) + cells3 * 12),
            "spectrum passes touch 1/4 .. 1 of the (2n)^2 (n+1) complex array (L2-resident for "
            "one beam)"),
        "sc_field_kernel": (
            lambda: lib.ch_sc_field(ws.phi.data_ptr(), ws.params.data_ptr(), B, grid, grid, grid,
                                    code, ws.field.data_ptr(), stream),
            B * cells3 * (4 + 32), "phi read + paired field write"),
        "sc_gather_kick_kernel": (
            lambda: lib.ch_sc_gather_kick(pp.data_ptr(), ps, ws.field.data_ptr(),
                                          ws.params.data_ptr(), n, B, grid, grid, grid, code,
                                          out.data_ptr(), None, stream),
            B * n * 56, "hbm: 28 B row read + 28 B row written per particle (gathers hit L2)"),
    }
    peak = 6450.0
    peaks = REPO / "MEASURED_PEAKS.json"
    if peaks.exists():
        peak = float(json.loads(peaks.read_text())["hbm_gbs"])
    kernels = []
    for name, (fn, nbytes, note) in stages.items():
        stage_ms, _ = timed(fn, 20 if B == 1 else 5, 3 if B == 1 else 1)
        kernels.append({
            "kernel": name, "us": stage_ms * 1e3, "algorithmic_bytes": nbytes,
            "achieved_gbs": nbytes / (stage_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": nbytes / (stage_ms * 1e-3) / 1e9 / peak, "bytes": note,
        })
    kick_us = sum(k["us"] for k in kernels)
    dominant = max(kernels, key=lambda k: k["us"])

    cpu_baseline = None
    if not args.no_cpu_baseline:
        torch.set_num_threads(torch.get_num_threads())
        el = {"type": "SpaceChargeKick", "effect_length": torch.tensor(1.0),
              "grid_shape": (grid, grid, grid)}
        cpu_beam = oracle.make_beam(particles.to(dtype), torch.tensor(1e8), particle_charges=charges)
        oracle.track_space_charge(el, cpu_beam)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            oracle.track_space_charge(el, cpu_beam)
        per_kick = (time.perf_counter() - t0) / reps
        # 100 kicks + linear runs: the kicks are > 99 % of the CPU time
        cpu_baseline = {
            "value": n * n_elements / (per_kick * kicks), "unit": "particle-steps/s",
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} space-charge kicks of {n} particles on {grid}^3 ({per_kick * 1e3:.0f} ms "
                      f"each), extrapolated to the {kicks} kicks of the lattice (linear runs neglected)",
        }

    line = {
        "metric": "particle-steps/sec (Segment.track, ParticleBeam)",
        "value": B * n * n_elements / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "n/a", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": f"{args.cells} FODO cells x [Quadrupole, Drift/2, SpaceChargeKick({grid}^3), "
                        f"Drift/2] x 2 = {n_elements} elements, {kicks} kicks, {n} particles, "
                        "total charge 1e-10 C, 1e8 eV -- BASELINE configs[3]"
                        + ("" if B == 1 else f"; {B} independent beams with per-beam particles and "
                           "charges (one GPU's share of BASELINE configs[4])"),
            "particles": n, "beams": B, "n_elements": n_elements, "kicks": kicks, "grid": grid,
        },
        "particle_kicks_per_s": B * n * kicks / (ms * 1e-3),
        "graph_value": None if graph_ms is None else B * n * n_elements / (graph_ms * 1e-3),
        "graph_ms_per_step": graph_ms,
        "gpu_launches": launches,
        "roofline": {
            "kernel": dominant["kernel"], "bound": "hbm", "achieved": dominant["achieved_gbs"],
            "peak": peak, "unit": "GB/s", "frac": dominant["frac_of_hbm_peak"], "traffic": None,
        },
        "kick_us_sum_of_stages": kick_us,
        "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
