"""Quick device-side timing of one SpaceChargeKick and the FODO config (development aid)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402
from cheetah_b200 import _capi  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (time.perf_counter() - t0) / reps * 1e3


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    grid = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = "cuda"
    t = lambda v: torch.tensor(v, device=dev)  # noqa: E731
    beam = cb.ParticleBeam.from_parameters(
        num_particles=n, total_charge=torch.tensor(1e-10), energy=torch.tensor(1e8),
        device=dev, dtype=torch.float32, generator=torch.Generator().manual_seed(0),
    )
    kick = cb.SpaceChargeKick(effect_length=t(1.0), grid_shape=(grid, grid, grid))
    ms, wall = timed(lambda: kick.track(beam))
    print(f"one kick  N={n} grid={grid}^3: {ms:.3f} ms device, {wall:.3f} ms wall, "
          f"{n / ms / 1e3:.1f} M particle-kicks/s")

    def drift_with_kick(length):
        return [cb.Drift(length=t(length / 2)),
                cb.SpaceChargeKick(effect_length=t(length), grid_shape=(grid, grid, grid)),
                cb.Drift(length=t(length / 2))]

    cells = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    elements = []
    for _ in range(cells):
        elements += [cb.Quadrupole(length=t(0.2), k1=t(4.2)), *drift_with_kick(1.0),
                     cb.Quadrupole(length=t(0.2), k1=t(-4.2)), *drift_with_kick(1.0)]
    segment = cb.Segment(elements)
    ms, wall = timed(lambda: segment.track(beam), reps=3, warm=1)
    kicks = 2 * cells
    print(f"FODO x{cells} ({len(elements)} elements, {kicks} kicks): {ms:.2f} ms device, {wall:.2f} ms wall, "
          f"{n * len(elements) / ms / 1e3:.1f} M particle-steps/s, {n * kicks / ms / 1e3:.1f} M particle-kicks/s, "
          f"{ms / kicks * 1e3:.1f} us per kick+run")
    graphed = cb.GraphedTrack(segment, beam)
    ms, wall = timed(lambda: graphed.replay(), reps=3, warm=1)
    print(f"FODO x{cells} CUDA graph replay: {ms:.2f} ms device, {wall:.2f} ms wall, "
          f"{n * len(elements) / ms / 1e3:.1f} M particle-steps/s, {ms / kicks * 1e3:.1f} us per kick+run")
    eager = segment.track(beam)
    replayed = graphed.replay()
    eager2 = segment.track(beam)
    scale = eager.particles.std(dim=0)
    print("eager vs eager (atomics order), max |diff| / column std:",
          ((eager.particles - eager2.particles).abs().amax(dim=0) / scale.clamp_min(1e-30)).tolist())
    print("graph vs eager, max |diff| / column std:",
          ((eager.particles - replayed.particles).abs().amax(dim=0) / scale.clamp_min(1e-30)).tolist())
    print("graph == eager:", torch.equal(eager.particles, replayed.particles) or
          float((eager.particles - replayed.particles).abs().max()))
    print("launches", _capi.launch_count())


if __name__ == "__main__":
    main()
