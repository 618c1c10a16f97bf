"""Device-side timing of Segment.track on the ARES workloads of bench.py (development aid):
config 3 (sparse branch), the solenoid + tilted-quadrupole variant (coupled branch) and a
variant with an off-crest cavity map replaced by a tau-coupled custom map (dense branch)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import workloads  # noqa: E402


def time_track(segment, beam, reps=5):
    out = None
    for _ in range(3):
        del out  # one (settings, N, 7) array at a time: 115 GB at 4096 settings
        out = segment.track(beam)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        del out
        out = segment.track(beam)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, float(out.survival_probabilities.mean())


def main():
    settings = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = 1_000_000
    device, dtype = torch.device("cuda", 0), torch.float32
    beam = workloads.product_beam(workloads.twiss_beam_particles(n), device, dtype)
    for name, description in (("sparse", workloads.ares_config3(settings, dtype)),
                              ("coupled", workloads.ares_config3_dense(settings, dtype)),
                              ("dense", workloads.ares_config3_tau_coupled(settings, dtype))):
        segment = workloads.product_segment(description, device, dtype)
        ms, survival = time_track(segment, beam)
        gbs = (settings * n * 32 + n * 32) / ms / 1e6
        print(f"{name:8s} B={settings}: {ms:7.3f} ms  {gbs:7.1f} GB/s  ({gbs / 6553:.3f} of 6553)  "
              f"survival {survival:.3f}")
        del segment


if __name__ == "__main__":
    main()
