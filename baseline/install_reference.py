#!/usr/bin/env python
"""Install the UNMODIFIED reference package into ``baseline/_ref`` (git-ignored, travels to the GPU box with gpurun).

    python baseline/install_reference.py

The reference's build backend is ``poetry-core`` (``pyproject.toml:56-58``), which is not in this image and cannot be
fetched (no network), so ``pip install /root/reference`` fails in the backend import.  The package is pure Python: this
script copies the tree to /tmp (``/root/reference`` is read-only), swaps ONLY the ``[build-system]`` table for setuptools
(no source file is touched) and lets pip build and install the wheel:

    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/...

``--no-deps``: the pinned dependencies (torch 2.3.1, normflows, jsonpickle, matplotlib, ...) are not resolvable offline;
the arm runs on the image's torch, the missing non-numeric packages are stubbed by ``oracle/ref_shim.py`` (test / baseline
infrastructure) at import time.  Nothing under ``baseline/_ref`` is imported by the product (``awesome_b200/``)."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("AWESOME_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "awesome")):
        print(f"reference tree not found at {SRC}; nothing installed")
        return 1
    tmp = tempfile.mkdtemp(prefix="awesome_ref_")
    try:
        shutil.copytree(os.path.join(SRC, "awesome"), os.path.join(tmp, "awesome"))
        for f in ("README.md", "LICENSE"):
            if os.path.exists(os.path.join(SRC, f)):
                shutil.copy(os.path.join(SRC, f), tmp)
        version = "0.1.0"
        m = re.search(r'^version\s*=\s*"([^"]+)"', open(os.path.join(SRC, "pyproject.toml")).read(), re.M)
        if m:
            version = m.group(1)
        with open(os.path.join(tmp, "pyproject.toml"), "w") as fh:
            fh.write('[build-system]\nrequires = ["setuptools"]\nbuild-backend = "setuptools.build_meta"\n\n'
                     f'[project]\nname = "awesome"\nversion = "{version}"\n\n'
                     '[tool.setuptools.packages.find]\ninclude = ["awesome*"]\n')
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DST, tmp]
        print(" ".join(cmd), flush=True)
        return subprocess.run(cmd).returncode
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    raise SystemExit(main())
