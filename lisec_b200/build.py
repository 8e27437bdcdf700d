"""Build liblisec_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery).

    python -m lisec_b200.build          # incremental
    python -m lisec_b200.build --clean
"""
from __future__ import annotations

import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblisec_b200.so")


def build(clean: bool = False, quiet: bool = True) -> str:
    if clean:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=quiet)
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building liblisec_b200.so failed:\n" + r.stdout + r.stderr)
    if not quiet:
        sys.stdout.write(r.stdout)
    if not os.path.exists(LIB):
        raise RuntimeError("make succeeded but %s is missing" % LIB)
    return LIB


if __name__ == "__main__":
    print(build(clean="--clean" in sys.argv, quiet=False))
