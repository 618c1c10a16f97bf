"""SASS evidence for the Blackwell-specific instructions each kernel family relies on: counts and
the first occurrences, from `cuobjdump -sass` of the built library (no GPU needed).

    python tools/sass_excerpts.py > profiles/r02_sass_excerpts.txt
"""
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "cheetah_b200" / "lib" / "libcheetah_b200.so"
# kernel name fragment -> mnemonics to look for
FAMILIES = {
    "apply_maps_kernel": ["UBLKCP", "SYNCS", "LDS.128", "FFMA"],
    "observe_maps_kernel": ["FFMA2", "FMUL2", "FADD2", "UBLKCP"],
    "compose_maps_kernel": ["DFMA", "SHFL"],
    "sc_deposit_kernel": ["RED.E.ADD.F32x4", "REDG.E.ADD.F32x4", "UBLKCP"],
    "sc_deposit_fixed_kernel": ["RED.E.ADD.64", "REDG.E.ADD.64", "ATOMG.E.ADD.64"],
    "sc_gather_brick_kernel": ["LDG.E.ENL2.256", "LDG.E.256", "LDCU", "UBLKCP", "MUFU.RSQ"],
    "sc_gather_kick_kernel": ["LDG.E.ENL2.256", "LDG.E.256", "UBLKCP"],
    "sc_field_brick_kernel": ["STG.E.128", "LDS"],
    "fftr_strided_kernel": ["FFMA", "LDG.E.64", "STG.E.64"],
    "sc_green_lattice_kernel": ["DFMA", "MUFU.RCP64H", "DMUL"],
    "track_nonlinear": ["DFMA", "UBLKCP"],
}


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    print(f"# cuobjdump -sass {LIB.name}: architectures {arch}")
    kernels = defaultdict(list)
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line.split("/*")[1].split("*/")[1].strip().rstrip(";").strip())
    for fragment, mnemonics in FAMILIES.items():
        matching = {k: v for k, v in kernels.items() if fragment in k}
        if not matching:
            continue
        biggest = max(matching, key=lambda k: len(matching[k]))
        print(f"\n## {fragment}: {len(matching)} instantiation(s); largest = {biggest[:100]}"
              f" ({len(matching[biggest])} instructions)")
        for mnemonic in mnemonics:
            hits = [ins for ins in matching[biggest] if mnemonic in ins]
            total = sum(1 for v in matching.values() for ins in v if mnemonic in ins)
            if not total:
                continue
            print(f"  {mnemonic:18s} {len(hits):5d} in it, {total:6d} over all instantiations"
                  + (f"   e.g. {hits[0]}" if hits else ""))


if __name__ == "__main__":
    sys.exit(main())
