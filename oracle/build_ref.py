"""Install the UNMODIFIED reference (desy-ml/cheetah, /root/reference) into ``oracle/_ref/``.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference is pure Python, so "building" it is a
``pip install --no-deps --target oracle/_ref`` of the checkout (from a copy under /tmp because
/root/reference is read-only and setuptools writes ``build/`` next to ``setup.py``).  No
reference source is copied into the repository history: ``oracle/_ref/`` is git-ignored, but
NOT gpurun-ignored, so the installed package travels to the GPU box like the built ``.so``.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box)

Users: ``bench.py --impl reference`` (the reference arm, ``kind: "reference"``), the
``reference_gpu`` section of bench.py, ``tests/test_integration_gpu.py`` (the INTEGRATION.md
binding executed with real ``cheetah`` objects) -- all through ``oracle/reference.py``.
"""

from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ORACLE = Path(__file__).resolve().parent
TARGET = ORACLE / "_ref"
SOURCE = Path("/root/reference")
STAMP = TARGET / "INSTALLED_FROM"


def reference_revision() -> str:
    head = SOURCE / ".git" / "HEAD"
    try:
        text = head.read_text().strip()
        if text.startswith("ref:"):
            text = (SOURCE / ".git" / text.split()[1]).read_text().strip()
        return text
    except OSError:
        return "unknown"


def build(force: bool = False) -> Path | None:
    """Install the reference; returns the target directory, or None when there is no checkout
    to install from (then whatever ``oracle/_ref`` already holds is used as is)."""
    if not SOURCE.exists():
        return TARGET if (TARGET / "cheetah").exists() else None
    if (TARGET / "cheetah").exists() and STAMP.exists() and not force:
        return TARGET
    if TARGET.exists():
        shutil.rmtree(TARGET)
    with tempfile.TemporaryDirectory() as tmp:
        copy = Path(tmp) / "reference"
        shutil.copytree(SOURCE, copy, ignore=shutil.ignore_patterns(".git", "docs", "images"))
        result = subprocess.run(
            [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
             "--no-deps", "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(copy)],
            capture_output=True, text=True,
        )
    if result.returncode != 0:
        raise RuntimeError(f"pip install of the reference failed:\n{result.stdout}\n{result.stderr}")
    STAMP.write_text(f"{SOURCE} @ {reference_revision()} (pip install --no-deps --target)\n")
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
