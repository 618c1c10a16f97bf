"""Run-to-run reproducibility of one kick and of k kicks (atomic ordering is the only source)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402

dev = "cuda"
t = lambda v: torch.tensor(v, device=dev)  # noqa: E731
beam = cb.ParticleBeam.from_parameters(num_particles=1_000_000, total_charge=torch.tensor(1e-10),
                                       energy=torch.tensor(1e8), device=dev, dtype=torch.float32,
                                       generator=torch.Generator().manual_seed(0))
kick = cb.SpaceChargeKick(effect_length=t(1.0), grid_shape=(64, 64, 64))
a = kick.track(beam).particles.clone()
b = kick.track(beam).particles.clone()
d = (a - beam.particles)
print("one kick: max|kick| per column", d.abs().amax(dim=0).tolist())
print("one kick: run-to-run max|diff| / max|kick|", ((a - b).abs().amax(dim=0) / d.abs().amax(dim=0).clamp_min(1e-30)).tolist())


def cell(k1):
    return [cb.Quadrupole(length=t(0.2), k1=t(k1)), cb.Drift(length=t(0.5)),
            cb.SpaceChargeKick(effect_length=t(1.0), grid_shape=(64, 64, 64)), cb.Drift(length=t(0.5))]


for cells in (1, 5, 25, 50):
    elements = []
    for _ in range(cells):
        elements += cell(4.2) + cell(-4.2)
    seg = cb.Segment(elements)
    x = seg.track(beam).particles
    y = seg.track(beam).particles
    print(f"{2 * cells:4d} kicks: run-to-run max|diff|/std", ((x - y).abs().amax(dim=0) / x.std(dim=0).clamp_min(1e-30)).tolist()[:6],
          " rms diff/std", ((x - y).square().mean(dim=0).sqrt() / x.std(dim=0).clamp_min(1e-30)).tolist()[:6],
          " sigma_x", float(x[:, 0].std()))
