"""Segment.track_moments on ARES (observables epilogue only) for profiling (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tools.quick_apply_bench import build  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    covariance = len(sys.argv) > 3 and sys.argv[3] == "cov"
    segment, beam = build(batch, n)
    for _ in range(3):
        observed = segment.track_moments(beam, covariance=covariance)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        observed = segment.track_moments(beam, covariance=covariance)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"track_moments B={batch} N={n} cov={covariance}: {ms:.3f} ms, "
          f"{batch * n / ms / 1e6:.1f} G particle-settings/s, sigma_x[0] = {float(observed.sigma[0, 0]):.4e}")


if __name__ == "__main__":
    main()
