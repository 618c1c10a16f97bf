"""cProfile of the host side of Segment.track for a small beam (development aid)."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tools.quick_apply_bench import build  # noqa: E402


def main():
    segment, beam = build(1, 10_000)
    for _ in range(20):
        segment.track(beam)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2000):
        segment.track(beam)
    torch.cuda.synchronize()
    print(f"eager Segment.track, ARES x 10k particles: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per call")
    profiler = cProfile.Profile()
    profiler.enable()
    for _ in range(2000):
        segment.track(beam)
    profiler.disable()
    stats = pstats.Stats(profiler)
    stats.sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
