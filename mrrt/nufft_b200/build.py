"""Build libb200nufft.so in-tree for sm_100a:  python -m mrrt.nufft_b200.build"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False, jobs=None):
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-C", csrc, "clean"])
    jobs = jobs or os.cpu_count() or 4
    subprocess.check_call(["make", "-C", csrc, "-j%d" % jobs])
    return os.path.join(HERE, "libb200nufft.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
