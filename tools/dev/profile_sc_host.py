"""Host-side profile (cProfile) of one eager Segment.track through 100 space-charge kicks, one beam."""
import cProfile
import pstats
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import bench_space_charge as sc  # noqa: E402
import workloads  # noqa: E402

device, dtype = torch.device("cuda", 0), torch.float32
segment = workloads.product_segment(workloads.fodo_space_charge(50, 64, dtype), device, dtype)
beam = sc.make_beam(1_000_000, 1, device, dtype)
for _ in range(3):
    segment.track(beam)
torch.cuda.synchronize()
profiler = cProfile.Profile()
profiler.enable()
for _ in range(3):
    segment.track(beam)
torch.cuda.synchronize()
profiler.disable()
stats = pstats.Stats(profiler)
stats.sort_stats("cumulative").print_stats(45)
