"""Recipe for oracle/_ref: the REFERENCE ITSELF (holoviz/datashader, pure Python + numba) made importable on the
GPU box, so that `bench.py --impl reference` and `cpu_baseline` time the reference's own numba kernels
(`kind: "reference"`) instead of the C port.

    python oracle/make_ref.py            # needs /root/reference; writes only under oracle/_ref/

The reference is interpreted, so "building" it is staging its package from where it lies (/root/reference/datashader,
tests and example data left out) next to the three import shims it needs in this image (xarray / toolz /
multipledispatch stand-ins that hold containers and functional helpers, no arithmetic: tests/golden/_shims).
oracle/_ref/ is git-ignored - reference sources never enter the history - but not gpurun-ignored, so it travels to
the GPU box like the built .so files.  TEST / BENCH INFRASTRUCTURE ONLY: nothing under datashader_b200/ imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/datashader"
OUT = os.path.join(HERE, "_ref")
SHIMS = os.path.join(os.path.dirname(HERE), "tests", "golden", "_shims")


def make_ref(force=False):
    if not os.path.isdir(REF):
        return None
    dst = os.path.join(OUT, "datashader")
    if os.path.isdir(dst) and not force:
        return OUT
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("tests", "__pycache__", "*.pyc", "*.nc", "*.png", "*.tif"))
    shutil.copytree(SHIMS, os.path.join(OUT, "_shims"), ignore=shutil.ignore_patterns("__pycache__"))
    return OUT


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
