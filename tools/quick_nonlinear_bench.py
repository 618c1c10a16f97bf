"""Small driver for profiling the non-linear and diagnostics kernels under ncu (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402
import workloads  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    device, dtype = "cuda", torch.float32
    t = lambda v: torch.tensor(v, device=device, dtype=dtype)  # noqa: E731
    particles = workloads.parameters_beam_particles(n).to(device=device, dtype=dtype)
    species = cb.Species("electron", device=device, dtype=dtype)
    batched = cb.ParticleBeam(particles.expand(B, n, 7).contiguous(), t(1e8), species=species)
    shared = cb.ParticleBeam(particles, t(1e8), species=species)
    if len(sys.argv) > 3 and sys.argv[3] in ("drift_kick_drift", "second_order"):  # fused line
        import bench_nonlinear
        from cheetah_b200 import lattice_description
        method = sys.argv[3]
        description = bench_nonlinear.fodo(method, 5, B, dtype)
        segment = cb.Segment(lattice_description.build(description, device=device, dtype=dtype))
        for _ in range(3):
            segment.track(shared)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            segment.track(shared)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{method} x 20 elements, {B} settings x {n}: {ms:.3f} ms, "
              f"{B * n * 20 / ms / 1e6:.1f} G particle-steps/s")
        return
    cases = {
        "drift_dkd": cb.Segment([cb.Drift(length=t(1.0), tracking_method="drift_kick_drift")]),
        "quad_second_order": cb.Segment([cb.Quadrupole(length=t(0.2), k1=t(4.2),
                                                       tracking_method="second_order")]),
        "dipole_dkd": cb.Segment([cb.Dipole(length=t(0.5), angle=t(0.2),
                                            tracking_method="drift_kick_drift")]),
        "quad_dkd": cb.Segment([cb.Quadrupole(length=t(0.2), k1=t(4.2), num_steps=5,
                                              tracking_method="drift_kick_drift")]),
    }
    if len(sys.argv) > 3:  # one single-element case by name
        cases = {sys.argv[3]: cases[sys.argv[3]]}
    for name, segment in cases.items():
        for _ in range(3):
            out = segment.track(batched)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            out = segment.track(batched)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{name}: {ms:.3f} ms, {B * n * 56 / ms / 1e6:.0f} GB/s")
    screen = cb.Screen(is_active=True, pixel_size=t([2e-5, 2e-5]))
    for _ in range(3):
        screen.track(shared)
        image = screen.reading
    torch.cuda.synchronize()
    print("screen sum", float(image.sum()))


if __name__ == "__main__":
    main()
