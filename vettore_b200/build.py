"""Builds libvettore_b200.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libvettore_b200.so")


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    jobs = jobs or max(1, (os.cpu_count() or 2))
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    if force:
        cmd.append("-B")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libvettore_b200.so failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
