"""Build ``lib/libcheetah_b200.so`` with nvcc for sm_100a (in-tree, no JIT cache)."""

from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
SOURCES = sorted((ROOT / "csrc").glob("*.cu"))
HEADERS = sorted((ROOT / "csrc").glob("*.cuh")) + [ROOT.parent / "include" / "cheetah_b200.h"]
OUTPUT = ROOT / "lib" / "libcheetah_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-I", str(ROOT.parent / "include"),
]


def up_to_date() -> bool:
    if not OUTPUT.exists():
        return False
    built = OUTPUT.stat().st_mtime
    return all(src.stat().st_mtime <= built for src in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return OUTPUT
    OUTPUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", str(OUTPUT)]
    cmd += [str(src) for src in SOURCES]
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{result.stdout}\n{result.stderr}")
    if verbose:
        print(result.stderr)
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
