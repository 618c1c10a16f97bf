"""profiles/sc_traffic.json from an `ncu --set full` report of one space-charge kick: DRAM bytes
(read + written) per beam and stage, summed over the kernels of the stage and averaged over the
captured kicks.  bench_space_charge.py scales them by the number of beams for `traffic`.

    python tools/sc_traffic_from_ncu.py gpurun_out/.../sc_b128.ncu-rep 128 > profiles/sc_traffic.json
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

STAGES = {
    "sc_moments_kernel": "moments",
    "sc_deposit_kernel": "deposit",
    "sc_green_lattice_kernel": "green",
    "fftr_even_z_kernel": "green",
    "fftr_even_strided_kernel": "green",
    "fft_even_pass_kernel": "green",
    "sc_quad_sum_kernel": "poisson",
    "fftr_r2c_z_kernel": "poisson",
    "fftr_strided_kernel": "poisson",
    "fftr_c2r_z_kernel": "poisson",
    "fft_r2c_z_kernel": "poisson",
    "fft_strided_kernel": "poisson",
    "fft_c2r_z_kernel": "poisson",
    "sc_field_brick_kernel": "field_gather",
    "sc_gather_brick_kernel": "field_gather",
    "sc_field_kernel": "field",
    "sc_gather_kick_kernel": "gather",
}
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path: str, beams: int) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    name_col = header.index("Kernel Name")
    cols = [header.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum")]
    time_col = header.index("gpu__time_duration.sum")
    per_kernel = defaultdict(lambda: [0, 0.0, 0.0])
    for row in rows[2:]:
        kernel = next((k for k in STAGES if k in row[name_col]), None)
        if kernel is None:
            continue
        total = sum(float(row[c]) * UNITS[units[c]] for c in cols)
        entry = per_kernel[kernel]
        entry[0] += 1
        entry[1] += total
        entry[2] += float(row[time_col]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[time_col]]
    # launches of each kernel in ONE kick (the capture window may straddle two kicks)
    per_kick_launches = {"fftr_even_strided_kernel": 2, "fft_even_pass_kernel": 3,
                         "fftr_strided_kernel": 3, "fft_strided_kernel": 3}
    kicks = 1
    stages = defaultdict(float)
    detail = {}
    for kernel, (launches, total, us) in per_kernel.items():
        per_launch = total / launches
        stages[STAGES[kernel]] += per_launch * per_kick_launches.get(kernel, 1) / beams
        detail[kernel] = {"launches_captured": launches, "dram_bytes_per_launch": per_launch,
                          "us_per_launch_under_ncu": us / launches}
    json.dump({
        "source": path, "beams": beams, "kicks_captured": kicks,
        "what": "dram__bytes_read.sum + dram__bytes_write.sum per beam and kick, by stage",
        "dram_bytes_per_beam": dict(stages), "kernels": detail,
    }, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
