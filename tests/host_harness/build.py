"""Builds tests/host_harness/libharness.so (TEST INFRASTRUCTURE; host build of the RUMI_HD kernel arithmetic)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libharness.so")


def build():
    src = os.path.join(HERE, "harness.cpp")
    deps = [src] + [os.path.join(HERE, "../../rumi_slam_b200/csrc", f)
                    for f in ("orb_geom.h", "orb_math.cuh", "octree_core.cuh", "orb_common.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=gnu++17", "-fPIC", "-shared", "-x", "c++",
                               src, "-o", LIB, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build())
