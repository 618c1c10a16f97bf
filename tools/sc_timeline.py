"""In-situ timeline of the C-ABI calls of one Segment.track (development aid): wraps every
library function with CUDA events recorded on the stream the call is enqueued on and prints, per
function, call count, summed duration and the start/end offsets of the first kick's calls."""
import sys
from collections import defaultdict
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import cheetah_b200 as cb  # noqa: E402
from cheetah_b200 import _capi  # noqa: E402
import workloads  # noqa: E402


class Proxy:
    def __init__(self, lib):
        self._lib = lib
        self.records = []
        self.enabled = False

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("ch_") or name in ("ch_last_error", "ch_kernel_launch_count",
                                                   "ch_abi_version", "ch_nonlinear_constants_len"):
            return fn

        def wrapped(*args):
            if not self.enabled:
                return fn(*args)
            current = torch.cuda.current_stream()
            stream = current if args[-1] == current.cuda_stream else torch.cuda.ExternalStream(args[-1])
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            status = fn(*args)
            b.record(stream)
            self.records.append((name, a, b))
            return status

        return wrapped


def main():
    beams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n, device, dtype = 1_000_000, torch.device("cuda", 0), torch.float32
    segment = workloads.product_segment(workloads.fodo_space_charge(cells, 64, dtype), device, dtype)
    particles = workloads.parameters_beam_particles(n).to(device=device, dtype=dtype)
    if beams > 1:
        particles = particles.expand(beams, n, 7).contiguous()
    charges = torch.full((beams, n) if beams > 1 else (n,), 1e-10 / n, dtype=dtype, device=device)
    beam = cb.ParticleBeam(particles, torch.tensor(1e8, device=device), particle_charges=charges,
                           species=cb.Species("electron", device=device, dtype=dtype))
    beam._unit_seventh = True
    import os

    if os.environ.get("CH_TIMELINE_UNFUSED"):  # every stage as its own pass
        from cheetah_b200 import tracking

        tracking.fuse_space_charge = False
    proxy = Proxy(_capi.lib())
    _capi._lib = proxy
    for _ in range(2):
        segment.track(beam)
    torch.cuda.synchronize()
    proxy.enabled = True
    origin = torch.cuda.Event(enable_timing=True)
    origin.record()
    segment.track(beam)
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    totals = defaultdict(lambda: [0, 0.0])
    print(f"whole track: {origin.elapsed_time(end):.3f} ms, {len(proxy.records)} calls")
    for i, (name, a, b) in enumerate(proxy.records):
        totals[name][0] += 1
        totals[name][1] += a.elapsed_time(b)
        if i < 24:
            print(f"  {name:28s} start {origin.elapsed_time(a):8.3f} ms  end {origin.elapsed_time(b):8.3f} ms"
                  f"  ({a.elapsed_time(b) * 1e3:8.1f} us)")
    for name, (count, ms) in sorted(totals.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:28s} x{count:4d}  {ms:9.3f} ms total  {ms / count * 1e3:9.1f} us each")


if __name__ == "__main__":
    main()
