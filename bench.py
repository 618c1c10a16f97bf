"""bench.py -- particle-steps/s of Segment.track(ParticleBeam) (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ares|space_charge]

``--workload ares`` (default, the headline): one step = one ``Segment.track`` call on the
195-element ARES lattice, 1e6 particles (one beam shared by all settings), 4096 vectorised
magnet settings sharded over the ranks (BASELINE configs[2]; fits one B200: 131 GB of output).
particle-steps = settings x particles x 195.
``--workload space_charge``: one step = one ``Segment.track`` call on the 50-cell FODO lattice
with 100 SpaceChargeKicks on a 64^3 grid (400 elements), 1e6 particles per beam, 1024
independent beams sharded over the ranks (BASELINE configs[4]; configs[3] is its one-beam case
and is reported in the ``space_charge`` section of the default line).

Keys of the one JSON line printed by rank 0:
  value        device-timed whole-job throughput, inputs resident in HBM (CUDA events, K steps,
               barrier + synchronize on both sides, max over ranks)
  e2e          same metric through the host-buffer API (cheetah_b200.host): inputs uploaded from
               pinned host memory and the outgoing beam downloaded to host memory and read by a
               consumer (a checksum over every chunk) inside the timed region
  roofline     dominant kernel: algorithmic bytes per launch / mean launch time (CUDA events on
               the launching stream during the timed steps) against MEASURED_PEAKS.json hbm_gbs
  parity_check this rank-0 GPU's output on a sample of the workload against the CPU oracle
  cpu_baseline the reference's own CPU implementation (oracle/_ref = the unmodified reference
               installed by oracle/build_ref.py; the oracle port if that is absent) timed on this
               box's host cores on a bounded sample of the same workload (rank 0, N=1 only)
  reference_gpu the unmodified reference on this B200 (``segment.to("cuda")``, eager and
               ``torch.compile``), chunked over the settings -- "the existing GPU implementation"
  space_charge / dense / observables / config2: the other BASELINE configs and kernel families
``--impl reference`` runs that CPU reference as the reference arm.
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
SC_CELLS, SC_GRID = 50, 64


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--workload", choices=["ares", "space_charge"], default="ares")
    p.add_argument("--settings", type=int, default=4096, help="total vectorised settings (ares)")
    p.add_argument("--beams", type=int, default=1024, help="total beams (space_charge)")
    p.add_argument("--particles", type=int, default=1_000_000)
    p.add_argument("--cpu-sample-settings", type=int, default=8)
    p.add_argument("--e2e-steps", type=int, default=2)
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-observables", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-space-charge", action="store_true")
    p.add_argument("--no-dense", action="store_true")
    p.add_argument("--no-reference-gpu", action="store_true")
    p.add_argument("--reference-gpu-child", choices=["eager", "compile"], help=argparse.SUPPRESS)
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
        self.first = 0

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

    def mark(self, timeout: float = 3.0) -> None:
        """Call right before the timed region, with the same load already running: waits until
        nvidia-smi delivers (its start-up can take longer than a 200 ms timed region) and makes
        stop() report the samples from here on."""
        if self.proc is None:
            return
        deadline = time.time() + timeout
        while not self.lines and time.time() < deadline:
            time.sleep(0.01)
        self.first = len(self.lines)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples of the timed region; if it was too short for one, those of the warm-up steps
        # (the same kernels) that ran right before it
        lines = self.lines[self.first:] or self.lines
        for line in lines:
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
# the reference on the host cores (cpu_baseline and --impl reference)
# ----------------------------------------------------------------------------------------
def cpu_reference_tracker(lattice: list, particles: torch.Tensor, dtype):
    """(callable returning (particles, survival) of one pass, kind): the unmodified reference's
    ``Segment.track`` (segment.py:545-574; its transfer-map cache warm after the first call) when
    ``oracle/_ref`` holds it, else the oracle port."""
    from oracle import reference

    if reference.available():
        segment = reference.segment(lattice, dtype=dtype)
        beam = reference.particle_beam(particles, 1e8, dtype=dtype)

        def run():
            out = segment.track(beam)
            return out.particles, out.survival_probabilities

        return run, "reference"
    from oracle import track_oracle as oracle

    beam = workloads.oracle_beam(particles, dtype)

    def run():
        out = oracle.track(lattice, beam)
        return out["particles"], out["survival_probabilities"]

    return run, "port"


def time_cpu_ares(n_settings_total: int, begin: int, end: int, n_particles: int, steps: int,
                  warmup: int, keep_output: bool = False) -> dict:
    dtype = torch.float32
    sample = end - begin
    lattice = workloads.ares_config3(n_settings_total, dtype, begin, end)
    particles = workloads.twiss_beam_particles(n_particles)
    run, kind = cpu_reference_tracker(lattice, particles, dtype)
    threads = torch.get_num_threads()
    for _ in range(warmup):
        run()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = run()
        times.append(time.perf_counter() - t0)
    assert out[0].shape == (sample, n_particles, 7)
    mean = sum(times) / len(times)
    name = ("cheetah.Segment.track of the unmodified reference (oracle/_ref)" if kind == "reference"
            else "oracle port (oracle/track_oracle.py)")
    return {
        "value": sample * n_particles * N_ELEMENTS / mean,
        "unit": UNIT,
        "cores": threads,
        "kind": kind,
        "sample": (
            f"{name}: settings {begin}..{end - 1} of {n_settings_total} x {n_particles} particles x "
            f"{N_ELEMENTS} elements per pass, float32, torch CPU {threads} threads, mean of {steps} "
            f"passes after {warmup} warm-up ({mean * 1e3:.0f} ms/pass)"
        ),
        "ms_per_pass": mean * 1e3,
        "output": out if keep_output else None,
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core
    torch.set_num_threads(os.cpu_count() or 1)
    steps = max(1, min(args.steps, 200))
    warmup = max(1, min(args.warmup, 20))
    if args.workload == "space_charge":
        import bench_space_charge as sc

        n_elements, kicks = 8 * SC_CELLS, 2 * SC_CELLS
        base = sc.cpu_kick_baseline(args.particles, SC_GRID, n_elements, kicks,
                                    reps=max(1, min(steps, 5)))
        config = space_charge_config(args, per_rank=-(-args.beams // args.gpus))
        ms = base["ms_per_kick"] * kicks
    else:
        base = time_cpu_ares(args.settings, 0, min(args.cpu_sample_settings, args.settings),
                             args.particles, steps, warmup)
        from cheetah_b200 import sharding

        begin, end = sharding.shard_bounds(args.settings, 0, args.gpus)
        config = workload_config(args, per_rank=end - begin)
        ms = base["ms_per_pass"]
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config,
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


def space_charge_config(args, per_rank: int) -> dict:
    return {
        "workload": (
            f"{SC_CELLS} FODO cells x [Quadrupole, Drift/2, SpaceChargeKick({SC_GRID}^3), Drift/2] x 2 "
            f"= {8 * SC_CELLS} elements, {2 * SC_CELLS} kicks, {args.particles} particles per beam, "
            f"{args.beams} independent beams (per-beam particles, total charge linspace(1e-11, "
            "1e-9)), 1e8 eV -- BASELINE configs[4]"
        ),
        "beams": args.beams,
        "beams_per_rank": per_rank,
        "particles": args.particles,
        "n_elements": 8 * SC_CELLS,
        "kicks": 2 * SC_CELLS,
        "grid": SC_GRID,
        "parallelism": f"beams sharded over {args.gpus} rank(s), no per-step collective",
        "l2": "inputs/outputs larger than L2: each kick streams beams_per_rank x particles x 56 B",
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


class Context:
    """Process group, device and the barrier / max-over-ranks helpers of one bench process."""

    def __init__(self, args) -> None:
        import torch.distributed as dist

        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device"
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world}"
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self) -> None:
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, value: float) -> float:
        t = torch.tensor([value], device=self.device, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum_over_ranks(self, value: float) -> float:
        t = torch.tensor([value], device=self.device, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])

    def close(self) -> None:
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def device_timed(ctx: Context, fn, steps: int, warmup: int) -> float:
    """ms per step: CUDA events on the current stream, barrier + synchronize on both sides, max
    over ranks."""
    for _ in range(warmup):
        fn()
    ctx.barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        fn()
    stop.record()
    ctx.barrier()
    return ctx.max_over_ranks(start.elapsed_time(stop)) / steps


# ----------------------------------------------------------------------------------------
# the reference on this GPU (child process: a compile that hangs must not take the bench down)
# ----------------------------------------------------------------------------------------
def reference_gpu_child(args) -> None:
    """Times the unmodified reference's ``Segment.track`` on cuda:0 over the ARES settings in
    chunks (the reference materialises ~10 (chunk, N, 7) temporaries per call); prints one JSON
    object on the real stdout."""
    from oracle import reference

    mode = args.reference_gpu_child
    device, dtype = torch.device("cuda", 0), torch.float32
    chunk = 64
    n_chunks = min(4, -(-args.settings // chunk))
    particles = workloads.twiss_beam_particles(args.particles)
    beam = reference.particle_beam(particles, 1e8, device=device, dtype=dtype)
    segments = []
    for c in range(n_chunks):
        lattice = workloads.ares_config3(args.settings, dtype, c * chunk,
                                         min((c + 1) * chunk, args.settings))
        segments.append(reference.segment(lattice, device=device, dtype=dtype))
    n_settings = sum(min((c + 1) * chunk, args.settings) - c * chunk for c in range(n_chunks))
    tracks = [s.track for s in segments]
    if mode == "compile":
        tracks = [torch.compile(t) for t in tracks]
    t0 = time.perf_counter()
    for track in tracks:  # warm-up: transfer-map caches, compilation
        out = track(beam)
    torch.cuda.synchronize()
    first_pass_s = time.perf_counter() - t0
    del out
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    start.record()
    for _ in range(reps):
        for track in tracks:
            out = track(beam)
            del out
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps
    emit({
        "mode": mode, "ms_per_pass": ms, "settings": n_settings, "chunk_settings": chunk,
        "value": n_settings * args.particles * N_ELEMENTS / (ms * 1e-3), "unit": UNIT,
        "first_pass_s": first_pass_s,
        "sample": f"cheetah.Segment.track of the unmodified reference on cuda:0 ({mode}), "
                  f"{n_settings} of {args.settings} settings in chunks of {chunk} x "
                  f"{args.particles} particles, float32, warm transfer-map cache, mean of {reps} "
                  "passes (CUDA events)",
    })


def reference_gpu_section(args) -> dict | None:
    from oracle import reference

    if not reference.available():
        return {"unavailable": "oracle/_ref (the installed reference) is absent"}
    out = {}
    for mode, limit in (("eager", 240), ("compile", 420)):
        cmd = [sys.executable, str(REPO / "bench.py"), "--reference-gpu-child", mode,
               "--settings", str(args.settings), "--particles", str(args.particles)]
        try:
            result = subprocess.run(cmd, capture_output=True, text=True, timeout=limit)
            lines = [ln for ln in result.stdout.splitlines() if ln.startswith("{")]
            if result.returncode == 0 and lines:
                out[mode] = json.loads(lines[-1])
            else:
                out[mode] = {"failed": (result.stderr or "").strip().splitlines()[-1:]}
        except subprocess.TimeoutExpired:
            out[mode] = {"failed": f"no result within {limit} s"}
    return out


# ----------------------------------------------------------------------------------------
# ARES workload (BASELINE configs[2]) + the sections for the other configs
# ----------------------------------------------------------------------------------------
def ares_parity(ctx: Context, args, beam, reference_out) -> dict:
    """Settings 0..S-1 of the workload on this GPU against (a) the float64 CPU oracle and (b) the
    float32 output of the CPU reference that cpu_baseline has just produced."""
    from oracle import lattice_io
    from oracle import track_oracle as oracle
    from tests import golden_utils as gu

    dtype = torch.float32
    sample = min(args.cpu_sample_settings, args.settings)
    lattice = workloads.ares_config3(args.settings, dtype, 0, sample)
    segment = workloads.product_segment(lattice, ctx.device, dtype)
    out = segment.track(beam)
    ours_p, ours_s = out.particles.cpu(), out.survival_probabilities.cpu()
    particles = workloads.twiss_beam_particles(args.particles)
    truth = oracle.track(lattice_io.cast(lattice, torch.float64),
                         workloads.oracle_beam(particles, torch.float64))
    result = {
        "what": f"settings 0..{sample - 1} x {args.particles} particles of this workload, GPU float32 "
                "vs the float64 CPU oracle (max |error| / max |coordinate| per column; survival "
                "masks compared exactly)",
        "max_col_err": gu.column_scaled_error(ours_p, truth["particles"]),
        "mask_flips": int((ours_s.double() != truth["survival_probabilities"]).sum()),
        "tolerance": 2e-6,
        "survivors": int(truth["survival_probabilities"].sum()),
    }
    if reference_out is not None:
        ref_p, ref_s = reference_out
        result["max_col_err_vs_cpu_reference_f32"] = gu.column_scaled_error(ours_p, ref_p.double())
        result["mask_flips_vs_cpu_reference_f32"] = int((ours_s != ref_s).sum())
        result["cpu_reference_f32_own_err"] = gu.column_scaled_error(ref_p, truth["particles"])
    result["ok"] = bool(result["max_col_err"] < 2e-6 and result["mask_flips"] == 0)
    return result


def ares_dense_section(ctx: Context, args, beam, per_rank: int, begin: int, end: int) -> dict:
    """The other two branches of the apply kernel on the same workload.  `coupled` (56
    multiply-adds): ARES with both solenoids powered and one quadrupole tilted (x-y coupling,
    solenoid.py:74-116, track_methods.py:345-382).  `tau_coupled` (all 72): the same followed by a
    CustomTransferMap with a tau column and a changed delta row, so that no sparsity flag holds.
    The top-level keys are the coupled variant (what this section has reported since round 2)."""
    from cheetah_b200 import tracking

    dtype = torch.float32
    peak, _ = peak_hbm()
    nbytes = per_rank * args.particles * 32 + args.particles * 32

    traffic_path = REPO / "profiles" / "apply_maps_traffic.json"
    branches = json.loads(traffic_path.read_text()).get("branches", {}) \
        if traffic_path.exists() else {}

    def measure(lattice, label, branch):
        # DRAM bytes of one ncu --set full capture of this branch at 256 settings, scaled per
        # (particle, setting) to this launch (see the headline's traffic_note)
        per_unit = branches.get(branch.split()[0], {}).get("dram_bytes_per_particle_setting")
        segment = workloads.product_segment(lattice, ctx.device, dtype)
        out = None
        for _ in range(2):
            del out
            out = segment.track(beam)
        ctx.barrier()
        tracking.apply_events = []
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = max(3, args.steps // 2)
        start.record()
        for _ in range(steps):
            del out
            out = segment.track(beam)
        stop.record()
        ctx.barrier()
        apply_ms = [a.elapsed_time(b) for a, b in tracking.apply_events]
        tracking.apply_events = None
        survival = float(out.survival_probabilities.mean())
        del out
        torch.cuda.empty_cache()
        ms = ctx.max_over_ranks(start.elapsed_time(stop)) / steps
        mean_apply = sum(apply_ms) / len(apply_ms)
        return {
            "workload": label,
            "ms_per_step": ms,
            "value": args.settings * args.particles * N_ELEMENTS / (ms * 1e-3),
            "unit": UNIT,
            "mean_survival": survival,
            "roofline": {
                "kernel": f"apply_shared_beam_kernel<float, NAP=3, ELLIPTICAL=0>, {branch}",
                "bound": "hbm", "achieved": nbytes / (mean_apply * 1e-3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": nbytes / (mean_apply * 1e-3) / 1e9 / peak,
                "mean_launch_ms": mean_apply, "algorithmic_bytes_per_launch": nbytes,
                "traffic": per_unit * per_rank * args.particles if per_unit else None,
            },
        }

    result = measure(
        workloads.ares_config3_dense(args.settings, dtype, begin, end),
        "ARES x 1e6 particles x 4096 settings with ARLIMSOG1A/B powered (k = 0.5, -0.4 1/m) and "
        "AREAMQZM2 tilted by 0.3 rad: x-y coupled maps", "coupled branch (56 multiply-adds)")
    result["tau_coupled"] = measure(
        workloads.ares_config3_tau_coupled(args.settings, dtype, begin, end),
        "the same followed by a CustomTransferMap with a tau column and a changed delta row: no "
        "sparsity flag holds", "dense branch (72 multiply-adds)")
    return result


def peak_hbm() -> tuple[float, str]:
    peaks_path = REPO / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        return float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class Checksum:
    """Host-side consumer of the e2e path: reads every downloaded chunk -- counts the surviving
    (particle, setting) pairs in the mask and sums all outgoing coordinates -- so the result
    demonstrably exists on the host.  Called from HostTracker's worker threads; each chunk is
    split over a pool of reader threads (numpy releases the GIL) because one core reads ~10 GB/s
    and the PCIe link delivers ~55."""

    def __init__(self, readers: int | None = None) -> None:
        from concurrent.futures import ThreadPoolExecutor

        self.survivors = 0.0
        self.total = 0.0
        self.chunks = 0
        self.lock = threading.Lock()
        self.readers = readers or max(2, min(12, (os.cpu_count() or 4) - 4))
        self.pool = ThreadPoolExecutor(max_workers=self.readers)

    @staticmethod
    def _piece(coordinates, survival):
        import numpy as np

        if survival.dtype == np.uint8:
            survivors = float(np.count_nonzero(survival))
        else:
            survivors = float(survival.sum(dtype=np.float64))
        # float32 pairwise sums per setting (SIMD, ~10 GB/s per core), float64 across settings
        return survivors, float(coordinates.sum(axis=1).sum(dtype=np.float64))

    def __call__(self, begin, end, coordinates_host, survival_host) -> None:
        coordinates = coordinates_host.numpy().reshape(coordinates_host.shape[0], -1)
        survival = survival_host.numpy().reshape(survival_host.shape[0], -1)
        count = coordinates.shape[0]
        pieces = min(self.readers, count)
        bounds = [count * i // pieces for i in range(pieces + 1)]
        futures = [self.pool.submit(self._piece, coordinates[a:b], survival[a:b])
                   for a, b in zip(bounds[:-1], bounds[1:])]
        survivors = total = 0.0
        for future in futures:
            s_piece, t_piece = future.result()
            survivors += s_piece
            total += t_piece
        with self.lock:
            self.survivors += survivors
            self.total += total
            self.chunks += 1


def run_ares(args) -> None:
    import cheetah_b200  # noqa: F401  (fails loudly if the CUDA library is missing)
    import cheetah_b200 as cb
    from cheetah_b200 import _capi, lattice_description, sharding, tracking
    from cheetah_b200.host import HostTracker

    ctx = Context(args)
    rank, world, device = ctx.rank, ctx.world, ctx.device
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
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    for _ in range(max(1, args.warmup)):
        del out
        out = segment.track(beam)
    survival_mean = float(out.survival_probabilities.mean())
    ctx.barrier()
    tracking.apply_events = []
    launches_before = _capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    ctx.barrier()
    start.record()
    for _ in range(args.steps):
        del out
        out = segment.track(beam)
    stop.record()
    ctx.barrier()
    local_ms = start.elapsed_time(stop)
    launches = _capi.launch_count() - launches_before
    apply_ms = [a.elapsed_time(b) for a, b in tracking.apply_events]
    tracking.apply_events = None
    clocks = sampler.stop()
    del out
    torch.cuda.empty_cache()
    ms_per_step = ctx.max_over_ranks(local_ms) / args.steps
    value = args.settings * args.particles * N_ELEMENTS / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches) -----------------------------
    peak, peak_kind = peak_hbm()
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
        "kernel": "apply_shared_beam_kernel<float, NAP=3, ELLIPTICAL=0> (ch_apply_maps: one beam under "
                  "consecutive settings; 4 particles per thread, 256 threads), sparse branch",
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
        "share_of_step": mean_apply_ms / (local_ms / args.steps),
    }

    # ---- the dense branch of the same kernel ---------------------------------------------------
    dense = None
    if not args.no_dense:
        dense = ares_dense_section(ctx, args, beam, per_rank, begin, end)

    # ---- fused observables: moments of the outgoing beam, no (B, N, 7) array in HBM ----------
    observables = None
    if not args.no_observables:
        observed = None

        def observe():
            nonlocal observed
            observed = segment.track_moments(beam)

        o_ms = device_timed(ctx, observe, args.steps, 2)
        cov_ms = device_timed(ctx, lambda: segment.track_moments(beam, covariance=True),
                              max(3, args.steps // 2), 1)
        observables = {
            "what": "Segment.track_moments: mu, sigma (6 each) and surviving-particle count per "
                    "setting from observe_shared_beam_kernel (packed FFMA2 pairs); outgoing particles never "
                    "written",
            "value": args.settings * args.particles * N_ELEMENTS / (o_ms * 1e-3),
            "unit": UNIT,
            "ms_per_step": o_ms,
            "covariance_ms_per_step": cov_ms,
            # settings that lose (almost) every particle have no finite sigma, as in the reference
            "mean_sigma_x": float(
                observed.sigma[..., 0][observed.sigma[..., 0].isfinite()].mean()
            ),
        }

    # ---- end to end through the host-buffer API -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_description = workloads.ares_config3(args.settings, dtype, begin, end)
        host_segment = cb.Segment(
            elements=lattice_description.build(host_description, dtype=dtype)
        )
        host_beam = cb.ParticleBeam(
            particles=particles.to(dtype).pin_memory(), energy=torch.tensor(1e8, dtype=dtype),
            species=cb.Species("electron", dtype=dtype),
        )
        tracker = HostTracker(host_segment, args.particles, per_rank, device=device, dtype=dtype,
                              chunk_settings=64, ring=4)
        # reader threads of the host consumer: the box's cores are shared by the ranks
        readers = max(2, min(12, ((os.cpu_count() or 8) - 4) // world))
        checksum = Checksum(readers)
        tracker.track(host_beam, consumer=checksum)  # warm-up (pins, lowers, first-touch)
        ctx.barrier()
        checksum = Checksum(readers)
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            tracker.track(host_beam, consumer=checksum)
            torch.cuda.synchronize()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e_h2d, e2e_d2h = tracker.h2d_bytes, tracker.d2h_bytes  # of one HostTracker.track
        host_survival = checksum.survivors / args.e2e_steps / (per_rank * args.particles)
        e2e_observables = None
        if not args.no_observables:
            tracker.track_moments(host_beam)
            ctx.barrier()
            reps = max(args.e2e_steps, 5)
            t0 = time.perf_counter()
            for _ in range(reps):
                moments_host = tracker.track_moments(host_beam)
            ctx.barrier()
            eo_s = ctx.max_over_ranks((time.perf_counter() - t0) / reps)
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
            "api": "cheetah_b200.host.HostTracker.track (CPU tensors in; outgoing coordinates "
                   f"({tracker.bytes_per_particle_setting} B per particle and setting) land in a "
                   "pinned host ring and a host consumer checksums every chunk; bytes are per rank)",
            "d2h_gbs_per_rank": e2e_d2h / e2e_s / 1e9,
            "d2h_gbs_all_ranks": ctx.sum_over_ranks(e2e_d2h) / e2e_s / 1e9,
            "host_consumer": {"chunks_read": checksum.chunks // args.e2e_steps,
                              "mean_survival_seen_on_host": host_survival},
        }
        del tracker
        torch.cuda.empty_cache()

    # ---- BASELINE configs[1]: same lattice, one setting (README magnet values), B = 1 -----------
    config2 = None
    if rank == 0 and world == 1 and not args.no_observables:
        segment2 = workloads.product_segment(workloads.ares_config2(dtype), device, dtype)
        for _ in range(3):
            segment2.track(beam)
        torch.cuda.synchronize()
        c_start, c_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50

        def batch_median(call):
            # these calls last tens of microseconds: the host's launch rate is part of what is
            # measured, so take the median of five batches instead of one
            samples = []
            for _ in range(5):
                c_start.record()
                for _ in range(reps):
                    call()
                c_stop.record()
                torch.cuda.synchronize()
                samples.append(c_start.elapsed_time(c_stop) / reps)
            return sorted(samples)[2]

        eager_ms = batch_median(lambda: segment2.track(beam))
        graphed = cb.GraphedTrack(segment2, beam)
        for _ in range(3):
            graphed.replay()
        torch.cuda.synchronize()
        graph_ms = batch_median(graphed.replay)
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
        del graphed

    # ---- BASELINE configs[3] and this rank's share of configs[4] -------------------------------
    space_charge = None
    if not args.no_space_charge:
        space_charge = space_charge_section(ctx, args)

    # ---- CPU baseline + parity of this workload (rank 0, single-GPU runs only) -----------------
    cpu_baseline = parity = reference_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample_settings, args.settings)
        base = time_cpu_ares(args.settings, 0, sample, args.particles, 3, 1, keep_output=True)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        parity = ares_parity(ctx, args, beam, base["output"])
        del base
        if not args.no_reference_gpu:
            torch.cuda.empty_cache()
            reference_gpu = reference_gpu_section(args)

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
            "parity_check": parity,
            "reference_gpu": reference_gpu,
            "dense": dense,
            "observables": observables,
            "config2": config2,
            "space_charge": space_charge,
            "gpu_launches": launches,
            "clocks": clocks,
            "particle_tracks_per_s": args.settings * args.particles / (ms_per_step * 1e-3),
            "mean_survival": survival_mean,
            "setup_broadcast_bytes": setup_bytes,
        }
        emit(line)
    ctx.close()


# ----------------------------------------------------------------------------------------
# space charge: section of the default line, and --workload space_charge
# ----------------------------------------------------------------------------------------
def space_charge_section(ctx: Context, args) -> dict:
    """configs[3] (one beam, rank 0 of a single-GPU run) and this rank's share of configs[4]
    (1024 beams / max(world, 8) ranks: 128 beams, one GPU's share of the 8-GPU configuration)."""
    import bench_space_charge as sc

    out = {}
    if ctx.rank == 0 and ctx.world == 1:
        out["config4"] = sc.section(
            ctx.device, args.particles, 1, SC_CELLS, SC_GRID, steps=3, warmup=2,
            label="50 FODO cells, 400 elements, 100 SpaceChargeKicks (64^3), 1e6 particles, one "
                  "beam -- BASELINE configs[3]")
        out["parity_check"] = sc.parity_check()
        torch.cuda.empty_cache()
    share = max(1, args.beams // max(ctx.world, 8))
    first = ctx.rank * share
    section = sc.section(
        ctx.device, args.particles, share, SC_CELLS, SC_GRID, steps=2, warmup=1, graph=False,
        first_beam=first, total_beams=args.beams,
        label=f"the same lattice, {share} independent beams on this GPU = one GPU's share of the "
              f"{args.beams}-beam BASELINE configs[4] on 8 GPUs")
    ctx.barrier()
    section["ms_per_step"] = ctx.max_over_ranks(section["ms_per_step"])
    # every rank tracks its own `share` beams with no collective: the job's throughput is the sum
    # (at 8 ranks this IS BASELINE configs[4]: 1024 beams on 8 GPUs)
    section["all_ranks"] = {
        "beams": share * ctx.world, "n_gpus": ctx.world,
        "value": share * ctx.world * args.particles * section["n_elements"]
                 / (section["ms_per_step"] * 1e-3),
        "unit": UNIT,
        "particle_kicks_per_s": share * ctx.world * args.particles * section["kicks"]
                                / (section["ms_per_step"] * 1e-3),
    }
    out["config5_share"] = section
    torch.cuda.empty_cache()
    return out


def run_space_charge(args) -> None:
    import cheetah_b200  # noqa: F401
    from cheetah_b200 import _capi, sharding
    import bench_space_charge as sc

    ctx = Context(args)
    rank, world, device = ctx.rank, ctx.world, ctx.device
    dtype = torch.float32
    begin, end = sharding.shard_bounds(args.beams, rank, world)
    per_rank = end - begin
    n = args.particles
    description = workloads.fodo_space_charge(SC_CELLS, SC_GRID, dtype)
    segment = workloads.product_segment(description, device, dtype)
    n_elements, kicks = len(description), 2 * SC_CELLS
    # one GPU holds at most `group` beams at a time (in + out + workspace); larger shards are
    # tracked group after group inside the step
    group = min(per_rank, 256)
    beams = [sc.make_beam(n, min(group, per_rank - g), device, dtype, begin + g, args.beams)
             for g in range(0, per_rank, group)]

    def step():
        for beam in beams:
            out = segment.track(beam)
            del out

    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    for _ in range(max(1, args.warmup)):
        step()
    ctx.barrier()
    launches_before = _capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark()
    ctx.barrier()
    start.record()
    for _ in range(args.steps):
        step()
    stop.record()
    ctx.barrier()
    launches = _capi.launch_count() - launches_before
    clocks = sampler.stop()
    ms_per_step = ctx.max_over_ranks(start.elapsed_time(stop)) / args.steps
    value = args.beams * n * n_elements / (ms_per_step * 1e-3)

    stages = sc.stage_table(beams[0], SC_GRID, reps=3)
    dominant = max((s for s in stages if s["stage"] != "moments"), key=lambda s: s["us"])
    peak, peak_kind = peak_hbm()
    us_per_beam_kick = ms_per_step * 1e3 / (kicks * per_rank)
    roofline = {
        "kernel": dominant["kernel"], "bound": "hbm", "achieved": dominant["achieved_gbs"],
        "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": dominant["frac_of_hbm_peak"],
        "traffic": dominant["traffic"], "algorithmic_bytes_per_launch": dominant["algorithmic_bytes"],
        "mean_launch_ms": dominant["us"] * 1e-3,
        "whole_kick": {
            "us_per_beam_kick": us_per_beam_kick,
            "particle_bytes_per_beam_kick": n * 76,
            "frac_of_particle_byte_floor": (n * 76 / (peak * 1e9)) / (us_per_beam_kick * 1e-6),
        },
    }

    # ---- e2e: host particles in, host particles out (one group of beams at a time) -----------
    e2e = None
    if not args.no_e2e:
        from cheetah_b200.host import track_host

        sample = beams[0]
        sample_beams = sample.particles.shape[0] if sample.particles.dim() == 3 else 1
        host_particles = sample.particles.cpu().pin_memory()
        host_charges = sample.particle_charges.cpu().pin_memory()
        buffers = None
        checksum = 0.0

        def host_step():
            nonlocal buffers, checksum
            out, _, buffers = track_host(segment, host_particles, 1e8, host_charges,
                                         device=device, buffers=buffers)
            checksum = float(out[..., 1].sum(dtype=torch.float64))

        host_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        ctx.barrier()
        e2e_s = ctx.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {
            "value": world * sample_beams * n * n_elements / e2e_s, "unit": UNIT,
            "h2d_bytes_per_step": sample_beams * n * 36, "d2h_bytes_per_step": sample_beams * n * 28,
            "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
            "api": "cheetah_b200.host.track_host: the first group of this rank's beams "
                   f"({sample_beams} beams) from pinned host memory through Segment.track and back "
                   "to pinned host memory, a host checksum over the result; value = ranks x "
                   "beams x particles x 400 / time",
            "host_checksum": checksum,
        }

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = sc.cpu_kick_baseline(n, SC_GRID, n_elements, kicks)
        cpu_baseline = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
    parity = sc.parity_check() if rank == 0 else None

    if rank == 0:
        emit({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": space_charge_config(args, per_rank), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "e2e": e2e, "parity_check": parity,
            "gpu_launches": launches, "clocks": clocks,
            "particle_kicks_per_s": args.beams * n * kicks / (ms_per_step * 1e-3),
            "us_per_beam_kick": us_per_beam_kick, "stages": stages,
        })
    ctx.close()


def main() -> None:
    args = parse_args()
    if args.reference_gpu_child:
        reference_gpu_child(args)
        return
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "space_charge":
        run_space_charge(args)
    else:
        run_ares(args)


if __name__ == "__main__":
    main()
