"""Latency of ch_compose_maps alone on ARES (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cheetah_b200 import tracking  # noqa: E402
from tools.quick_apply_bench import build  # noqa: E402


def main():
    for batch in (1, 16, 4096):
        segment, beam = build(batch, 1000)
        program = tracking._plan(list(segment.elements), beam.particles.device, (), segment)
        section = program.stages[0]
        for _ in range(5):
            tracking._compose(program, section, beam.energy, beam.species, torch.float32)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            tracking._compose(program, section, beam.energy, beam.species, torch.float32)
        b.record()
        torch.cuda.synchronize()
        eager = a.elapsed_time(b) / 50 * 1e3
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(50):
                tracking._compose(program, section, beam.energy, beam.species, torch.float32)
        graph.replay()
        torch.cuda.synchronize()
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        print(f"B={batch}: compose {eager:.1f} us per eager call, "
              f"{a.elapsed_time(b) / 50 * 1e3:.1f} us per launch inside a CUDA graph")


if __name__ == "__main__":
    main()
