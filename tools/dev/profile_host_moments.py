"""cProfile of HostTracker.track_moments (host tensors in, moments out) on config 3 (development aid)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import cheetah_b200 as cb  # noqa: E402
import workloads  # noqa: E402
from cheetah_b200 import lattice_description  # noqa: E402
from cheetah_b200.host import HostTracker  # noqa: E402

settings = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dtype = torch.float32
host_segment = cb.Segment(elements=lattice_description.build(
    workloads.ares_config3(settings, dtype), dtype=dtype))
particles = workloads.twiss_beam_particles(1_000_000)
host_beam = cb.ParticleBeam(particles=particles.to(dtype).pin_memory(),
                            energy=torch.tensor(1e8, dtype=dtype),
                            species=cb.Species("electron", dtype=dtype))
tracker = HostTracker(host_segment, 1_000_000, settings, device="cuda", dtype=dtype)
for _ in range(3):
    tracker.track_moments(host_beam)
t0 = time.perf_counter()
for _ in range(10):
    tracker.track_moments(host_beam)
print(f"track_moments host to host: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms per call")
profiler = cProfile.Profile()
profiler.enable()
for _ in range(10):
    tracker.track_moments(host_beam)
profiler.disable()
pstats.Stats(profiler).sort_stats("cumulative").print_stats(30)
