"""Space-charge workloads of BASELINE.json: configs[3] (50 FODO cells = 400 elements, 100
SpaceChargeKicks on a 64^3 grid, 1e6 particles, one beam) and configs[4] (the same lattice with
1024 independent beams sharded over the GPUs).  Imported by bench.py (its `space_charge` section
and `--workload space_charge`); also a stand-alone script with the same JSON schema:

    python bench_space_charge.py [--steps K] [--warmup W] [--cells 50] [--grid 64] [--beams B]

`value` is eager `Segment.track`; `graph_value` replays the same call from a CUDA graph
(cheetah_b200.GraphedTrack).  `kernels` lists, for every kernel family of one kick, the mean
duration from CUDA events around back-to-back launches of that stage alone, its algorithmic
bytes and the achieved bandwidth against MEASURED_PEAKS.json hbm_gbs; `traffic`
is dram bytes read + written from the committed `ncu --set full` capture
(profiles/sc_traffic.json), scaled per beam.
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

SIGMAS = (175e-6, 4e-6, 175e-6, 4e-6, 8e-6, 2e-3)  # ParticleBeam.from_parameters defaults


def hbm_peak() -> tuple[float, str]:
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def make_beam(n: int, n_beams: int, device, dtype, first_beam: int = 0, total_beams: int | None = None):
    """Config 4 (n_beams == 1): the seeded from_parameters beam, total charge 1e-10 C.  Config 5:
    beams `first_beam .. first_beam + n_beams` of `total_beams` independent Gaussian beams with the
    same second moments, drawn on the device (generator seed = global beam index) and total
    charges linspace(1e-11, 1e-9, total_beams)."""
    import cheetah_b200 as cb

    if n_beams == 1 and total_beams in (None, 1):
        particles = workloads.parameters_beam_particles(n).to(device=device, dtype=dtype)
        charges = torch.full((n,), 1e-10 / n, dtype=dtype, device=device)
    else:
        total_beams = total_beams or n_beams
        particles = torch.empty((n_beams, n, 7), device=device, dtype=dtype)
        sigma = torch.tensor(SIGMAS, device=device, dtype=dtype)
        g = torch.Generator(device=device)
        for i in range(n_beams):
            g.manual_seed(first_beam + i)
            particles[i, :, :6] = torch.randn((n, 6), device=device, dtype=dtype, generator=g) * sigma
        particles[..., 6] = 1.0
        total = torch.linspace(1e-11, 1e-9, total_beams, dtype=torch.float64)
        total = total[first_beam : first_beam + n_beams].to(device=device, dtype=dtype)
        charges = (total.unsqueeze(-1) / n).expand(n_beams, n).contiguous()
    beam = cb.ParticleBeam(
        particles, torch.tensor(1e8, device=device, dtype=dtype), particle_charges=charges,
        species=cb.Species("electron", device=device, dtype=dtype),
    )
    beam._unit_seventh = True
    return beam


def timed(fn, steps: int, warmup: int):
    """(ms per call, C-ABI launches per call) with CUDA events on the current stream."""
    from cheetah_b200 import _capi

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
    return a.elapsed_time(b) / steps, (_capi.launch_count() - before) / steps


def stage_table(beam, grid: int, reps: int | None = None) -> list[dict]:
    """Per-stage timings of ONE kick on `beam` (every C-ABI stage launched back to back alone)."""
    from cheetah_b200 import _capi, space_charge

    lib = _capi.lib()
    device = beam.particles.device
    n = beam.particles.shape[-2]
    B = beam.particles.shape[0] if beam.particles.dim() == 3 else 1
    one = torch.tensor(1.0, device=device)
    three = torch.tensor(3.0, device=device)
    out, ws = space_charge.kick(
        beam.particles, beam.energy, beam.particle_charges, beam.survival_probabilities,
        beam.species.mass_eV, one, (three, three, three), (grid, grid, grid),
    )
    torch.cuda.synchronize()
    stream = _capi.current_stream(device)
    code = _capi.CH_F32
    pp, q, w = beam.particles, beam.particle_charges, beam.survival_probabilities
    ps, qs = (0, 0) if B == 1 else (n * 7, n)
    out = out.reshape(B, n, 7)
    cells3 = grid ** 3
    spectrum_bytes = (2 * grid) * (2 * grid) * (grid + 1) * 8
    stages = {
        "moments": (
            "sc_moments_kernel (+grid parameters; first kick of a lattice only)",
            lambda: lib.ch_sc_moments_and_params(
                pp.data_ptr(), ps, w.data_ptr(), 0, n, B, beam.energy.data_ptr(), 0, code,
                beam.species.mass_eV.data_ptr(), code, one.data_ptr(), 0, code,
                three.data_ptr(), 0, three.data_ptr(), 0, three.data_ptr(), 0, code,
                grid, grid, grid, code, ws.stats.data_ptr(), ws.params.data_ptr(), stream),
            B * n * 32, "28 B row + 4 B survival per particle"),
        "deposit": (
            "sc_deposit_kernel",
            lambda: lib.ch_sc_deposit(
                pp.data_ptr(), ps, q.data_ptr(), qs, w.data_ptr(), 0, ws.params.data_ptr(), n, B,
                grid, grid, grid, code, ws.rho_quad.data_ptr(), stream),
            B * n * 36, "28 B row + charge + survival per particle (reductions stay in L2)"),
        "green": (
            "sc_green_lattice_kernel + 3 x fft_even_pass_kernel",
            lambda: (lib.ch_sc_green_function(ws.params.data_ptr(), B, grid, grid, grid, code,
                                              ws.lattice.data_ptr(), None, stream),
                     lib.ch_sc_green_spectrum(ws.lattice.data_ptr(), ws.params.data_ptr(), B, grid, grid,
                                              grid, code,
                                              ws.green_scratch.data_ptr(),
                                              ws.green_spectrum.data_ptr(), stream)),
            B * ((grid + 1) ** 3 * (8 + 8) + 4 * cells3 * 4),
            "fp64 lattice write + read, three real-even passes"),
        "poisson": (
            "fft_r2c_z, fft_strided<0>, fft_strided<2> (x conv), fft_strided<1>, fft_c2r_z",
            lambda: lib.ch_sc_poisson_solve(
                ws.rho_quad.data_ptr(), ws.green_spectrum.data_ptr(), ws.params.data_ptr(), B,
                grid, grid, grid, code, ws.rho_spectrum.data_ptr(), ws.phi.data_ptr(), stream),
            B * (int(spectrum_bytes * (0.25 + 0.5 + 0.5 + 1.0 + 0.5 + 0.5 + 0.25)) + cells3 * 12),
            "the passes touch 1/4 .. 1 of the (2n)^2 (n+1) complex spectrum, read + write"),
    }
    if ws.bricks:
        null = (None, 0, None, 0, None, None, None, 0, code, None, code, None, 0, code,
                None, 0, None, 0, None, 0, code, grid, grid, grid)
        stages["field_gather"] = (
            "sc_field_brick_kernel + sc_gather_brick_kernel per group of "
            f"{ws.group_beams} beam(s) (ch_sc_field_gather, nothing fused)",
            lambda: lib.ch_sc_field_gather(
                pp.data_ptr(), ps, ws.phi.data_ptr(), ws.field.data_ptr(), ws.group_beams,
                ws.params.data_ptr(), n, B, grid, grid, grid, code, *null, out.data_ptr(), None,
                stream),
            B * (n * 56 + cells3 * 4),
            "phi read; 28 B row read + 28 B row written per particle (96-byte bricks stay in L2)")
    else:
        stages["field"] = (
            "sc_field_kernel",
            lambda: lib.ch_sc_field(ws.phi.data_ptr(), ws.params.data_ptr(), B, grid, grid, grid,
                                    code, ws.field.data_ptr(), stream),
            B * cells3 * (4 + 32), "phi read + z-paired field write")
        stages["gather"] = (
            "sc_gather_kick_kernel",
            lambda: lib.ch_sc_gather_kick(
                pp.data_ptr(), ps, ws.field.data_ptr(), _capi.SC_FIELD_NODES,
                ws.params.data_ptr(), n, B, grid, grid, grid, code, out.data_ptr(), None, stream),
            B * n * 56, "28 B row read + 28 B row written per particle (gathers hit L1/L2)")
    peak, _ = hbm_peak()
    traffic = {}
    traffic_path = REPO / "profiles" / "sc_traffic.json"
    if traffic_path.exists():
        traffic = json.loads(traffic_path.read_text())
    rows = []
    for key, (name, fn, nbytes, note) in stages.items():
        r = reps or (20 if B == 1 else 5)
        stage_ms, _ = timed(fn, r, 3 if B == 1 else 1)
        per_beam = traffic.get("dram_bytes_per_beam", {}).get(key)
        rows.append({
            "stage": key, "kernel": name, "us": stage_ms * 1e3, "algorithmic_bytes": nbytes,
            "achieved_gbs": nbytes / (stage_ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": nbytes / (stage_ms * 1e-3) / 1e9 / peak,
            "traffic": None if per_beam is None else per_beam * B,
            "bytes": note,
        })
    return rows


def cpu_kick_baseline(n: int, grid: int, n_elements: int, kicks: int, reps: int = 3) -> dict:
    """One kick of one beam on the host cores: the unmodified reference when oracle/_ref holds it,
    else the oracle port; extrapolated to the lattice (the kicks are > 99 % of its CPU time)."""
    from oracle import reference

    dtype = torch.float32
    particles = workloads.parameters_beam_particles(n).to(dtype)
    charges = torch.full((n,), 1e-10 / n, dtype=dtype)
    threads = torch.get_num_threads()
    if reference.available():
        cheetah = reference.load()
        element = cheetah.SpaceChargeKick(effect_length=torch.tensor(1.0),
                                          grid_shape=(grid, grid, grid))
        beam = reference.particle_beam(particles, 1e8, dtype=dtype, particle_charges=charges)
        run, kind = (lambda: element.track(beam)), "reference"
    else:
        from oracle import track_oracle as oracle

        el = {"type": "SpaceChargeKick", "effect_length": torch.tensor(1.0),
              "grid_shape": (grid, grid, grid)}
        beam = oracle.make_beam(particles, torch.tensor(1e8), particle_charges=charges)
        run, kind = (lambda: oracle.track_space_charge(el, beam)), "port"
    run()
    t0 = time.perf_counter()
    for _ in range(reps):
        run()
    per_kick = (time.perf_counter() - t0) / reps
    return {
        "value": n * n_elements / (per_kick * kicks), "unit": "particle-steps/s",
        "cores": threads, "kind": kind,
        "sample": f"{reps} SpaceChargeKick.track calls of {n} particles on {grid}^3 "
                  f"({per_kick * 1e3:.0f} ms each, torch CPU {threads} threads), extrapolated to the "
                  f"{kicks} kicks of the lattice (linear sections neglected)",
        "ms_per_kick": per_kick * 1e3,
    }


def parity_check(n: int = 100_000, grid: int = 64) -> dict:
    """One kick + the following half drift on the GPU against the fp64 oracle (same inputs)."""
    from oracle import lattice_io
    from oracle import track_oracle as oracle

    device, dtype = torch.device("cuda", torch.cuda.current_device()), torch.float32
    description = workloads.fodo_space_charge(1, grid, dtype)[:4]  # quad, drift, kick, drift
    segment = workloads.product_segment(description, device, dtype)
    particles = workloads.parameters_beam_particles(n)
    charges = torch.full((n,), 1e-10 / n, dtype=torch.float64)
    import cheetah_b200 as cb

    beam = cb.ParticleBeam(
        particles.to(device=device, dtype=dtype), torch.tensor(1e8, device=device, dtype=dtype),
        particle_charges=charges.to(device=device, dtype=dtype),
        species=cb.Species("electron", device=device, dtype=dtype),
    )
    out = segment.track(beam).particles.double().cpu()
    expected = oracle.track(
        lattice_io.cast(description, torch.float64),
        oracle.make_beam(particles, torch.tensor(1e8, dtype=torch.float64), particle_charges=charges),
    )["particles"]
    no_kick = oracle.track(
        lattice_io.cast([d for d in description if d["type"] != "SpaceChargeKick"], torch.float64),
        oracle.make_beam(particles, torch.tensor(1e8, dtype=torch.float64), particle_charges=charges),
    )["particles"]
    kick = (expected - no_kick)[:, [1, 3, 5]]
    err = ((out - expected)[:, [1, 3, 5]].abs().amax(dim=0) / kick.abs().amax(dim=0)).max()
    return {
        "what": f"quadrupole + drift + SpaceChargeKick({grid}^3) + drift, {n} particles, float32 on "
                "the GPU vs the float64 CPU oracle",
        "max_error_relative_to_kick": float(err), "tolerance": 3e-3, "ok": bool(err < 3e-3),
    }


def run(device, n: int, n_beams: int, cells: int, grid: int, steps: int, warmup: int,
        first_beam: int = 0, total_beams: int | None = None, graph: bool = True,
        stages: bool = True) -> dict:
    """Track the FODO space-charge lattice with `n_beams` beams on `device`; returns timings."""
    import cheetah_b200 as cb

    dtype = torch.float32
    description = workloads.fodo_space_charge(cells, grid, dtype)
    segment = workloads.product_segment(description, device, dtype)
    beam = make_beam(n, n_beams, device, dtype, first_beam, total_beams)
    n_elements, kicks = len(description), 2 * cells
    ms, launches = timed(lambda: segment.track(beam), steps, warmup)
    graph_ms = None
    if graph:
        graphed = cb.GraphedTrack(segment, beam)
        graph_ms, _ = timed(lambda: graphed.replay(), steps, 1)
        del graphed
    rows = stage_table(beam, grid) if stages else None
    return {
        "beam": beam, "segment": segment, "ms": ms, "graph_ms": graph_ms, "launches": launches,
        "n_elements": n_elements, "kicks": kicks, "stages": rows,
    }


def section(device, n: int, n_beams: int, cells: int, grid: int, steps: int, warmup: int,
            label: str, **kwargs) -> dict:
    """JSON-ready summary of `run` (used by bench.py's `space_charge` section)."""
    r = run(device, n, n_beams, cells, grid, steps, warmup, **kwargs)
    best_ms = min(r["ms"], r["graph_ms"]) if r["graph_ms"] else r["ms"]
    out = {
        "workload": label, "beams": n_beams, "particles": n, "n_elements": r["n_elements"],
        "kicks": r["kicks"], "grid": grid,
        "ms_per_step": r["ms"], "value": n_beams * n * r["n_elements"] / (r["ms"] * 1e-3),
        "graph_ms_per_step": r["graph_ms"],
        "graph_value": None if r["graph_ms"] is None
        else n_beams * n * r["n_elements"] / (r["graph_ms"] * 1e-3),
        "unit": "particle-steps/s",
        "particle_kicks_per_s": n_beams * n * r["kicks"] / (best_ms * 1e-3),
        "us_per_beam_kick": best_ms * 1e3 / (r["kicks"] * n_beams),
        # 60-76 B of particle traffic per particle and kick
        "particle_bytes_floor_frac": (n * 76 / (hbm_peak()[0] * 1e9))
        / (best_ms * 1e-3 / (r["kicks"] * n_beams)),
        "gpu_launches_per_step": r["launches"],
    }
    if r["stages"] is not None:
        out["kick_us_sum_of_stages"] = sum(k["us"] for k in r["stages"])
        out["stages"] = r["stages"]
    return out


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

    device = torch.device("cuda", 0)
    n, grid, B = args.particles, args.grid, args.beams
    r = run(device, n, B, args.cells, grid, args.steps, args.warmup, graph=not args.no_graph)
    ms, graph_ms, kernels = r["ms"], r["graph_ms"], r["stages"]
    n_elements, kicks = r["n_elements"], r["kicks"]
    dominant = max(kernels, key=lambda k: k["us"])
    peak, _ = hbm_peak()
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline = cpu_kick_baseline(n, grid, n_elements, kicks)
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
        "gpu_launches": r["launches"] * args.steps,
        "roofline": {
            "kernel": dominant["kernel"], "bound": "hbm", "achieved": dominant["achieved_gbs"],
            "peak": peak, "unit": "GB/s", "frac": dominant["frac_of_hbm_peak"],
            "traffic": dominant["traffic"],
        },
        "kick_us_sum_of_stages": sum(k["us"] for k in kernels),
        "kernels": kernels,
        "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
