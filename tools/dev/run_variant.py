"""One Segment.track / track_moments call pattern on a bench workload, for ncu captures
(development aid):  python tools/dev/run_variant.py {sparse|coupled|dense} SETTINGS [moments|cov]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import workloads  # noqa: E402

kind, settings = sys.argv[1], int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "track"
device, dtype = torch.device("cuda", 0), torch.float32
beam = workloads.product_beam(workloads.twiss_beam_particles(1_000_000), device, dtype)
description = {"sparse": workloads.ares_config3, "coupled": workloads.ares_config3_dense,
               "dense": workloads.ares_config3_tau_coupled}[kind](settings, dtype)
segment = workloads.product_segment(description, device, dtype)
for _ in range(4):
    if mode == "track":
        out = segment.track(beam)
        del out
    else:
        segment.track_moments(beam, covariance=mode == "cov")
torch.cuda.synchronize()
