"""Builds libseekr_b200.so in-tree: ``python -m seekr_b200.build`` (nvcc, sm_100a; no GPU needed)."""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(jobs=8, verbose=False):
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j%d" % jobs]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libseekr_b200.so failed")
    return os.path.join(HERE, "lib", "libseekr_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))
