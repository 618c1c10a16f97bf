"""bench.py -- particle-steps/s of Segment.track(ParticleBeam) on ARES (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one ``Segment.track`` call: the 195-element ARES lattice, 1e6 particles (one beam
shared by all settings), 4096 vectorised magnet settings sharded over the ranks (BASELINE
configs[2]; fits one B200: 131 GB of output).  particle-steps = settings x particles x 195.

Lines printed by rank 0 (one JSON object):
  value        device-timed whole-job throughput, inputs resident in HBM (CUDA events, K steps,
               barrier + synchronize on both sides, max over ranks)
  e2e          same metric through the host-buffer API (cheetah_b200.host.HostTracker): beam and
               settings uploaded from pinned host memory and ALL output bytes downloaded to host
               memory inside the timed region
  roofline     ch_apply_maps: algorithmic bytes per launch / mean launch time (CUDA events on
               the launching stream during the timed steps) against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle (a PyTorch-CPU restatement of the reference, oracle/) timed on this
               box's host cores on a bounded sample of the same workload (rank 0, N=1 only)
``--impl reference`` times that CPU oracle as the reference arm (the reference is pure Python
and cannot travel to the GPU box; the oracle is pinned to it by tests/test_oracle_golden.py).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import workloads  # noqa: E402

METRIC = "particle-steps/sec (Segment.track, ParticleBeam)"
UNIT = "particle-steps/s"
N_ELEMENTS = workloads.N_ELEMENTS_ARES


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--settings", type=int, default=4096, help="total vectorised settings")
    p.add_argument("--particles", type=int, default=1_000_000)
    p.add_argument("--cpu-sample-settings", type=int, default=8)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-observables", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


# ----------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._reader, daemon=True).start()
        except OSError:
            self.proc = None

    def _reader(self) -> None:
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                sm_max.append(float(parts[2]))
            except ValueError:
                continue
            for name, value in zip(names, parts[5:9]):
                if value.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ----------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline and --impl reference)
# ----------------------------------------------------------------------------------------
def time_cpu_oracle(n_settings_total: int, sample_settings: int, n_particles: int, steps: int,
                    warmup: int) -> dict:
    from oracle import track_oracle as oracle

    dtype = torch.float32
    sample = min(sample_settings, n_settings_total)
    lattice = workloads.ares_config3(n_settings_total, dtype, 0, sample)
    beam = workloads.oracle_beam(workloads.twiss_beam_particles(n_particles), dtype)
    threads = torch.get_num_threads()
    for _ in range(warmup):
        oracle.track(lattice, beam)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = oracle.track(lattice, beam)
        times.append(time.perf_counter() - t0)
    assert out["particles"].shape == (sample, n_particles, 7)
    mean = sum(times) / len(times)
    return {
        "value": sample * n_particles * N_ELEMENTS / mean,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": (
            f"{sample} of {n_settings_total} settings x {n_particles} particles x {N_ELEMENTS} "
            f"elements per pass, float32, torch CPU {threads} threads, mean of {steps} passes "
            f"after {warmup} warm-up ({mean * 1e3:.0f} ms/pass)"
        ),
        "ms_per_pass": mean * 1e3,
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
    torch.set_num_threads(os.cpu_count() or 1)
    # one pass over the 8-setting sample takes ~0.12 s on 16 cores: K and W are honoured as given
    # (bounded only against absurd values so that the arm always ends within minutes)
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 20))
    base = time_cpu_oracle(args.settings, args.cpu_sample_settings, args.particles, steps, warmup)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": base["ms_per_pass"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, per_rank=args.settings),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, per_rank: int) -> dict:
    return {
        "workload": (
            f"ARES Segment ({N_ELEMENTS} elements incl. 3 apertures), {args.particles} particles "
            f"(one from_twiss beam shared by all settings), {args.settings} vectorised magnet "
            "settings (13 quadrupoles + 30 correctors), linear transfer maps -- BASELINE configs[2]"
        ),
        "settings": args.settings,
        "settings_per_rank": per_rank,
        "particles": args.particles,
        "n_elements": N_ELEMENTS,
        "parallelism": f"settings sharded over {args.gpus} rank(s), no per-step collective",
        "l2": "inputs/outputs larger than L2: each step writes settings_per_rank x particles x 32 B",
    }


# ----------------------------------------------------------------------------------------
def emit(line: dict) -> None:
    """Write the one JSON line to the REAL stdout (fd saved before libraries could print)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries (NCCL prints its version banner) write to fd 1; keep stdout clean for the JSON line
# by pointing fd 1 at stderr for the rest of the process.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main() -> None:
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    import cheetah_b200  # noqa: F401  (fails loudly if the CUDA library is missing)
    from cheetah_b200 import _capi, sharding, tracking
    from cheetah_b200.host import HostTracker

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device"
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dtype = torch.float32
    begin, end = sharding.shard_bounds(args.settings, rank, world)
    per_rank = end - begin

    # ---- setup: rank 0's beam and lattice scalars are broadcast once (the only collective) ----
    particles = workloads.twiss_beam_particles(args.particles)
    beam = workloads.product_beam(particles, device, dtype)
    description = workloads.ares_config3(args.settings, dtype, begin, end)
    segment = workloads.product_segment(description, device, dtype)
    setup_bytes = sharding.broadcast_module(beam)
    free, total = torch.cuda.mem_get_info(device)
    need = per_rank * args.particles * 32
    assert need < free * 0.95, f"workload needs {need / 1e9:.0f} GB, {free / 1e9:.0f} GB free"

    # ---- device-resident timing ----------------------------------------------------------------
    out = None
    for _ in range(max(1, args.warmup)):
        del out
        out = segment.track(beam)
    survival_mean = float(out.survival_probabilities.mean())
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    tracking.apply_events = []
    launches_before = _capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        del out
        out = segment.track(beam)
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = _capi.launch_count() - launches_before
    apply_ms = [a.elapsed_time(b) for a, b in tracking.apply_events]
    tracking.apply_events = None
    clocks = sampler.stop()
    del out
    torch.cuda.empty_cache()

    times = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms = float(times[0])
    ms_per_step = elapsed_ms / args.steps
    value = args.settings * args.particles * N_ELEMENTS / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches) -----------------------------
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_kind = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured"
    else:
        peak, peak_kind = 6650.0, "fallback"
    algorithmic_bytes = per_rank * args.particles * 32 + args.particles * 32
    mean_apply_ms = sum(apply_ms) / len(apply_ms)
    achieved = algorithmic_bytes / (mean_apply_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    traffic_path = REPO / "profiles" / "apply_maps_traffic.json"
    if traffic_path.exists():
        # dram__bytes_read + dram__bytes_write of one `ncu --set full` capture of the same kernel
        # at 256 settings (a 4096-setting launch writes 131 GB, too much for ncu's replay
        # save/restore), scaled per (particle, setting) to this launch
        per_unit = json.loads(traffic_path.read_text())["dram_bytes_per_particle_setting"]
        traffic = per_unit * per_rank * args.particles
        traffic_note = "ncu dram bytes at 256 settings, scaled per (particle, setting)"
    roofline = {
        "kernel": "apply_maps_kernel<float, 4, 256, UNIT7=1, MOMENTS=0, WRITE=1, CAVITY=0> "
                  "(ch_apply_maps)",
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "peak_kind": peak_kind,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic,
        "traffic_note": traffic_note,
        "algorithmic_bytes_per_launch": algorithmic_bytes,
        "mean_launch_ms": mean_apply_ms,
        "launches_timed": len(apply_ms),
        "share_of_step": mean_apply_ms / (start.elapsed_time(stop) / args.steps),
    }

    # ---- fused observables: moments of the outgoing beam, no (B, N, 7) array in HBM ----------
    observables = None
    if not args.no_observables:
        for _ in range(2):
            segment.track_moments(beam)
        barrier()
        o_start, o_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o_start.record()
        for _ in range(args.steps):
            observed = segment.track_moments(beam)
        o_stop.record()
        barrier()
        t = torch.tensor([o_start.elapsed_time(o_stop)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        o_ms = float(t[0]) / args.steps
        observables = {
            "what": "Segment.track_moments: mu, sigma (6 each) and surviving-particle count per "
                    "setting from observe_maps_kernel (packed FFMA2 pairs); outgoing particles never "
                    "written",
            "value": args.settings * args.particles * N_ELEMENTS / (o_ms * 1e-3),
            "unit": UNIT,
            "ms_per_step": o_ms,
            # settings that lose (almost) every particle have no finite sigma, as in the reference
            "mean_sigma_x": float(
                observed.sigma[..., 0][observed.sigma[..., 0].isfinite()].mean()
            ),
        }

    # ---- end to end through the host-buffer API -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_description = workloads.ares_config3(args.settings, dtype, begin, end)
        import cheetah_b200 as cb
        from cheetah_b200 import lattice_description

        host_segment = cb.Segment(
            elements=lattice_description.build(host_description, dtype=dtype)
        )
        host_beam = cb.ParticleBeam(
            particles=particles.to(dtype), energy=torch.tensor(1e8, dtype=dtype),
            species=cb.Species("electron", dtype=dtype),
        )
        tracker = HostTracker(host_segment, args.particles, per_rank, device=device, dtype=dtype,
                              chunk_settings=64, ring=2)
        tracker.track(host_beam)  # warm-up (pins, lowers, first-touch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            tracker.track(host_beam)
            torch.cuda.synchronize()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        e2e_h2d, e2e_d2h = tracker.h2d_bytes, tracker.d2h_bytes  # of HostTracker.track
        e2e_observables = None
        if not args.no_observables:
            tracker.track_moments(host_beam)
            barrier()
            t0 = time.perf_counter()
            for _ in range(max(args.e2e_steps, 3)):
                moments_host = tracker.track_moments(host_beam)
            barrier()
            eo_s = (time.perf_counter() - t0) / max(args.e2e_steps, 3)
            t = torch.tensor([eo_s], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            eo_s = float(t[0])
            e2e_observables = {
                "value": args.settings * args.particles * N_ELEMENTS / eo_s,
                "unit": UNIT,
                "ms_per_step": eo_s * 1e3,
                "h2d_bytes_per_step": tracker.h2d_bytes,
                "d2h_bytes_per_step": tracker.d2h_bytes,
                "api": "cheetah_b200.host.HostTracker.track_moments (CPU tensors in, moments on "
                       "the host out)",
                "host_mu_shape": list(moments_host.mu.shape),
            }
            observables["e2e"] = e2e_observables
        e2e = {
            "value": args.settings * args.particles * N_ELEMENTS / e2e_s,
            "unit": UNIT,
            "h2d_bytes_per_step": e2e_h2d,
            "d2h_bytes_per_step": e2e_d2h,
            "ms_per_step": e2e_s * 1e3,
            "steps": args.e2e_steps,
            "api": "cheetah_b200.host.HostTracker.track (CPU tensors in, pinned host ring out; "
                   "bytes are per rank)",
            "d2h_gbs_per_rank": e2e_d2h / e2e_s / 1e9,
        }
        del tracker
        torch.cuda.empty_cache()

    # ---- BASELINE configs[1]: same lattice, one setting (README magnet values), B = 1 -----------
    config2 = None
    if rank == 0 and world == 1 and not args.no_observables:
        import cheetah_b200 as cb

        segment2 = workloads.product_segment(workloads.ares_config2(dtype), device, dtype)
        for _ in range(3):
            segment2.track(beam)
        torch.cuda.synchronize()
        c_start, c_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        c_start.record()
        for _ in range(reps):
            segment2.track(beam)
        c_stop.record()
        torch.cuda.synchronize()
        eager_ms = c_start.elapsed_time(c_stop) / reps
        graphed = cb.GraphedTrack(segment2, beam)
        for _ in range(3):
            graphed.replay()
        torch.cuda.synchronize()
        c_start.record()
        for _ in range(reps):
            graphed.replay()
        c_stop.record()
        torch.cuda.synchronize()
        graph_ms = c_start.elapsed_time(c_stop) / reps
        config2 = {
            "workload": f"ARES Segment ({N_ELEMENTS} elements), {args.particles} particles, ONE "
                        "setting (README magnet values), linear maps -- BASELINE configs[1]; the "
                        "64 MB working set is L2-resident and the call is launch-latency bound",
            "ms_per_step": eager_ms,
            "value": args.particles * N_ELEMENTS / (eager_ms * 1e-3),
            "graph_ms_per_step": graph_ms,
            "graph_value": args.particles * N_ELEMENTS / (graph_ms * 1e-3),
            "unit": UNIT,
            "achieved_gbs_graph": args.particles * 64 / (graph_ms * 1e-3) / 1e9,
        }

    # ---- CPU baseline (rank 0, single-GPU runs only) --------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = time_cpu_oracle(args.settings, args.cpu_sample_settings, args.particles, 3, 1)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, per_rank),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "observables": observables,
            "config2": config2,
            "gpu_launches": launches,
            "clocks": clocks,
            "particle_tracks_per_s": args.settings * args.particles / (ms_per_step * 1e-3),
            "mean_survival": survival_mean,
            "setup_broadcast_bytes": setup_bytes,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
