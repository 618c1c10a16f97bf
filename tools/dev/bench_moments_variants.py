"""Segment.track_moments on the three bench lattices (sparse / coupled / dense maps), with and
without the covariance sums (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import workloads  # noqa: E402

settings = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
device, dtype = torch.device("cuda", 0), torch.float32
beam = workloads.product_beam(workloads.twiss_beam_particles(1_000_000), device, dtype)
for name, make in (("sparse", workloads.ares_config3), ("coupled", workloads.ares_config3_dense),
                   ("dense", workloads.ares_config3_tau_coupled)):
    segment = workloads.product_segment(make(settings, dtype), device, dtype)
    for covariance in (False, True):
        for _ in range(2):
            segment.track_moments(beam, covariance=covariance)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            segment.track_moments(beam, covariance=covariance)
        b.record()
        torch.cuda.synchronize()
        print(f"{name:8s} covariance={covariance!s:5s}: {a.elapsed_time(b) / 5:7.3f} ms per "
              f"{settings} x 1e6 step")
