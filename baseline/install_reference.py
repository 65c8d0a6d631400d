"""TEST / BENCH INFRASTRUCTURE ONLY — stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, travels to the
GPU box with the gpurun snapshot) so that bench.py can time the reference itself beside the GPU.

    python baseline/install_reference.py            # needs /root/reference (build container)

Recipe (the one offline install the task allows):
    pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>
The reference ships no packaging metadata (its pyproject.toml only configures pytest/isort, and setuptools' automatic
discovery refuses the flat layout with two top-level packages), so the install runs from a copy under /tmp to which a
five-line setup.py naming the packages is added.  No reference source file is modified, and nothing under
baseline/_ref is tracked by git.  oracle/reference_shims.py imports it from there (or from /root/reference)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("SFM_REFERENCE_SOURCE", "/root/reference")

SETUP_PY = """from setuptools import find_packages, setup
setup(name="structure_from_motion_reference", version="0", packages=find_packages(include=["lib", "lib.*"]))
"""


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "lib", "ransac", "ransac.py"))


def install(force: bool = False) -> str:
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "lib")):
        raise RuntimeError(f"no reference at {SOURCE}")
    tmp = tempfile.mkdtemp(prefix="sfm_ref_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, copy, ignore=shutil.ignore_patterns(".git", "__pycache__", "data"))
        with open(os.path.join(copy, "setup.py"), "w") as f:
            f.write(SETUP_PY)
        os.remove(os.path.join(copy, "pyproject.toml"))  # only pytest/isort settings; it hides setup.py from pip
        shutil.rmtree(TARGET, ignore_errors=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, copy])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not installed():
        raise RuntimeError("pip finished but baseline/_ref/lib/ransac/ransac.py is missing")
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
