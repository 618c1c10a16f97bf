"""Quick device-side timing of Segment.track on ARES (development aid, not bench.py)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402
from cheetah_b200 import _capi  # noqa: E402
from oracle import lattice_io  # noqa: E402
from tests import golden_utils as gu  # noqa: E402


def build(batch, n, dtype=torch.float32, device="cuda"):
    lattice = gu.ares_lattice(dtype)
    g = torch.Generator().manual_seed(1)
    for el in lattice:
        if el["type"] == "Quadrupole" and batch > 1:
            el["k1"] = ((torch.rand(batch, generator=g) * 2 - 1) * 5.0).to(dtype)
        elif el["type"] in ("HorizontalCorrector", "VerticalCorrector") and batch > 1:
            el["angle"] = ((torch.rand(batch, generator=g) * 2 - 1) * 2e-5).to(dtype)
    for name in ("ARLISLHG1", "ARBCSLHB1"):
        gu.set_attr(lattice, name, "x_max", torch.tensor(2e-3, dtype=dtype))
        gu.set_attr(lattice, name, "y_max", torch.tensor(2e-3, dtype=dtype))
    segment = gu.product_segment(lattice, device, dtype)
    beam = cb.ParticleBeam.from_twiss(
        num_particles=n, beta_x=3.14, beta_y=42.0, device=device, dtype=dtype,
        generator=torch.Generator().manual_seed(0),
    )
    return segment, beam


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    batches = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 16, 256]
    for batch in batches:
        segment, beam = build(batch, n)
        for _ in range(3):
            out = segment.track(beam)
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        t0 = time.perf_counter()
        start.record()
        for _ in range(reps):
            out = segment.track(beam)
        stop.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = start.elapsed_time(stop) / reps
        bytes_ = batch * n * 32 + n * 28
        print(
            f"B={batch:5d} N={n}: {ms:8.3f} ms/step device, {wall:8.3f} ms wall, "
            f"{bytes_ / ms / 1e6:8.1f} GB/s algorithmic, survival {float(out.survival_probabilities.mean()):.3f}, "
            f"launches {_capi.launch_count()}"
        )
        del out


if __name__ == "__main__":
    main()
