"""BASELINE configs[1] (ARES, 1e6 particles, one setting): eager and CUDA-graph replay timings
of Segment.track, checked against the eager result (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402
import workloads  # noqa: E402

device, dtype = torch.device("cuda", 0), torch.float32
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
beam = workloads.product_beam(workloads.twiss_beam_particles(n), device, dtype)
segment = workloads.product_segment(workloads.ares_config2(dtype), device, dtype)
for _ in range(3):
    eager = segment.track(beam)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 200
a.record()
for _ in range(reps):
    segment.track(beam)
b.record()
torch.cuda.synchronize()
eager_ms = a.elapsed_time(b) / reps
graphed = cb.GraphedTrack(segment, beam)
for _ in range(3):
    out = graphed.replay()
torch.cuda.synchronize()
a.record()
for _ in range(reps):
    out = graphed.replay()
b.record()
torch.cuda.synchronize()
graph_ms = a.elapsed_time(b) / reps
same = torch.equal(out.particles, eager.particles) and torch.equal(
    out.survival_probabilities, eager.survival_probabilities)
print(f"config 2, N={n}: eager {eager_ms * 1e3:.1f} us, graph replay {graph_ms * 1e3:.1f} us "
      f"({n * 64 / graph_ms / 1e6:.0f} GB/s), graph == eager: {same}")
