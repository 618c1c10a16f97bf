"""Segment.track in float64 on the config-3 workload shape (development aid): 56 B row + 8 B
survival written per (particle, setting)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import workloads  # noqa: E402

settings = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1_000_000
device, dtype = torch.device("cuda", 0), torch.float64
beam = workloads.product_beam(workloads.twiss_beam_particles(n), device, dtype)
for name, description in (("sparse", workloads.ares_config3(settings, dtype)),
                          ("coupled", workloads.ares_config3_dense(settings, dtype))):
    segment = workloads.product_segment(description, device, dtype)
    out = None
    for _ in range(2):
        del out
        out = segment.track(beam)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        del out
        out = segment.track(beam)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    gbs = settings * n * 64 / ms / 1e6
    print(f"{name:8s} float64, B={settings}: {ms:7.3f} ms  {gbs:7.1f} GB/s ({gbs / 6553:.3f} of 6553)")
    del segment, out
