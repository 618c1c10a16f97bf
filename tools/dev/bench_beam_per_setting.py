"""Segment.track with a beam PER setting (vectorised beams x settings, the RL-rollout shape):
reads 28 B and writes 32 B per (particle, setting) (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import cheetah_b200 as cb  # noqa: E402
import workloads  # noqa: E402

settings = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = 1_000_000
device, dtype = torch.device("cuda", 0), torch.float32
base = workloads.twiss_beam_particles(n).to(dtype).to(device)
particles = base.unsqueeze(0).repeat(settings, 1, 1)
particles[..., :6] *= (1.0 + 0.1 * torch.rand(settings, 1, 1, device=device))
beam = cb.ParticleBeam(particles=particles, energy=torch.tensor(1e8, device=device, dtype=dtype),
                       device=device, dtype=dtype)
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
    gbs = settings * n * 60 / ms / 1e6
    print(f"{name:8s} beam per setting, B={settings}: {ms:7.3f} ms  {gbs:7.1f} GB/s "
          f"({gbs / 6553:.3f} of 6553)")
    del segment
