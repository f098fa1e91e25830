"""Build tests/cpu_harness/libharness.so with g++ (TEST INFRASTRUCTURE; see harness.cpp)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "libharness.so")


def build(force=False):
    srcs = [os.path.join(HERE, "harness.cpp")] + [os.path.join(ROOT, "mpx_b200", "csrc", f)
                                                   for f in ("common.cuh", "srbd.cuh", "wb.cuh", "wb_tree.cuh")]
    srcs.append(os.path.join(ROOT, "include", "mpxb.h"))
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", srcs[0], "-o", OUT]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
