"""Builds tests/hostsim/libhostsim.so: the device math headers compiled for the host CPU
(TEST-ONLY, see hostsim.cpp).  Rebuilt when any csrc header or hostsim.cpp is newer."""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pos_evolution_b200", "csrc")
OUT = os.path.join(HERE, "libhostsim.so")


def build(force=False):
    srcs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "hostsim.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-x", "c++", "-I", CSRC,
           os.path.join(HERE, "hostsim.cpp"), "-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
