"""Per-opcode executed-instruction and stall-sample totals of one kernel in an .ncu-rep captured
with --import-source on (development aid):  python tools/ncu_opcode_mix.py REPORT UNITS
UNITS = number of (warp, setting) pairs the launch processed, to normalise the counts."""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, units):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header = rows[1]
    i_src, i_exec = header.index("Source"), header.index("Instructions Executed")
    i_stall = header.index("Warp Stall Sampling (All Samples)")
    ops, stalls, data = collections.Counter(), collections.Counter(), []
    for r in rows[2:]:
        try:
            n, st = int(r[i_exec]), int(r[i_stall])
        except (ValueError, IndexError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_src])
        op = m.group(2).split(".")[0] if m else r[i_src][:10]
        ops[op] += n
        stalls[op] += st
        data.append((r[i_src], n, st))
    total, total_stall = sum(ops.values()), sum(stalls.values())
    print(f"{total} warp instructions, {total / units:.1f} per unit, {total_stall} stall samples")
    for op, n in ops.most_common(30):
        print(f"{op:10s} {n / total * 100:6.2f} %  {n / units:7.1f} per unit   "
              f"{stalls[op] / total_stall * 100:5.1f} % of stall samples")
    print("-- blocks of 40 SASS instructions --")
    for i in range(0, len(data), 40):
        block = data[i:i + 40]
        executed = sum(d[1] for d in block) / units
        if executed >= 1.0:
            print(f"{i:5d} {executed:7.1f} per unit  {sum(d[2] for d in block) / total_stall * 100:5.1f} %"
                  f" stalls   {block[0][0].strip()[:70]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
