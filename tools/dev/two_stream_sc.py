"""Experiment: one 128-beam space-charge track against two concurrent 64-beam tracks on two CUDA
streams driven by two host threads (do stages with different bottlenecks overlap?)."""
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import bench_space_charge as sc  # noqa: E402
import workloads  # noqa: E402


def main():
    beams = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    parts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    device, dtype = torch.device("cuda", 0), torch.float32
    segment = workloads.product_segment(workloads.fodo_space_charge(cells, 64, dtype), device, dtype)
    whole = sc.make_beam(1_000_000, beams, device, dtype)
    per = beams // parts
    pieces = [sc.make_beam(1_000_000, per, device, dtype, i * per, beams) for i in range(parts)]

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    single = timed(lambda: segment.track(whole))
    streams = [torch.cuda.Stream(device) for _ in range(parts)]

    def split():
        def work(i):
            with torch.cuda.stream(streams[i]):
                segment.track(pieces[i])
        threads = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()

    concurrent = timed(split)
    kicks = 2 * cells
    print(f"{beams} beams, {kicks} kicks: single call {single:.1f} ms ({single / kicks:.2f} ms/kick), "
          f"{parts} x {per} beams on {parts} streams {concurrent:.1f} ms ({concurrent / kicks:.2f} ms/kick)")


if __name__ == "__main__":
    main()
