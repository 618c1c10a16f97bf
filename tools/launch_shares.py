"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares."""
import csv
import sys
from collections import defaultdict


def main(path: str, title: str) -> None:
    rows = []
    with open(path) as f:
        lines = [line for line in f if not line.startswith("==")]
    reader = csv.DictReader(lines)
    for row in reader:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        value = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        rows.append((row["Kernel Name"], value * scale))
    total = sum(v for _, v in rows)
    groups = defaultdict(list)
    for name, value in rows:
        groups[name].append(value)
    print(f"# {title}")
    print("# gpu__time_duration.sum per launch (cold-cache, serialised under ncu: compare SHARES)")
    print(f"{'kernel':92s} {'launches':>8s} {'mean_us':>12s} {'share_%':>8s}")
    for name, values in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name[:90]:92s} {len(values):8d} {sum(values) / len(values):12.1f} "
              f"{100 * sum(values) / total:8.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
