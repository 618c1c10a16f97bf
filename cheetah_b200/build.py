"""Build ``lib/libcheetah_b200.so`` with nvcc for sm_100a (in-tree, no JIT cache).

Every ``csrc/*.cu`` is compiled to its own object (in parallel; an object is rebuilt when its
source, any header or the flags are newer) and the objects are linked into one shared library.
``build(force=True)`` -- what ``__graft_entry__.build()`` runs -- recompiles everything.
"""

from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
SOURCES = sorted((ROOT / "csrc").glob("*.cu"))
HEADERS = sorted((ROOT / "csrc").glob("*.cuh")) + [ROOT.parent / "include" / "cheetah_b200.h"]
OUTPUT = ROOT / "lib" / "libcheetah_b200.so"
OBJECTS = ROOT / "lib" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", str(ROOT.parent / "include"),
]


def _object(src: Path) -> Path:
    return OBJECTS / (src.stem + ".o")


def _includes(path: Path, seen: set | None = None) -> list[Path]:
    """Headers of this repository that ``path`` includes, transitively (quoted includes only)."""
    import re

    seen = set() if seen is None else seen
    for name in re.findall(r'^\s*#include\s+"([^"]+)"', path.read_text(), flags=re.M):
        for candidate in (path.parent / name, ROOT.parent / "include" / name):
            if candidate.exists() and candidate not in seen:
                seen.add(candidate)
                _includes(candidate, seen)
    return sorted(seen)


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    built = target.stat().st_mtime
    return any(dep.stat().st_mtime > built for dep in deps)


def up_to_date() -> bool:
    return not _stale(OUTPUT, SOURCES + HEADERS + [Path(__file__)])


def _compile(src: Path, verbose: bool) -> str:
    cmd = ["nvcc", *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", str(src),
           "-o", str(_object(src))]
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{result.stdout}\n{result.stderr}")
    return result.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and up_to_date():
        return OUTPUT
    OBJECTS.mkdir(parents=True, exist_ok=True)
    todo = [src for src in SOURCES
            if force or _stale(_object(src), [src, *_includes(src), Path(__file__)])]
    with ThreadPoolExecutor(max_workers=min(len(todo) or 1, os.cpu_count() or 1)) as pool:
        logs = list(pool.map(lambda src: _compile(src, verbose), todo))
    if verbose:
        print("\n".join(logs))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(OUTPUT)]
    cmd += [str(_object(src)) for src in SOURCES]
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
        raise RuntimeError(f"nvcc link failed:\n{result.stdout}\n{result.stderr}")
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
