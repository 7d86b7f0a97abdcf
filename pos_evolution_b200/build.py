"""Builds pos_evolution_b200/libb200pos.so (the only native artefact of the product) with nvcc for
sm_100a.  In-tree so the .so travels with the repo snapshot to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "b200pos.cu")
OUT = os.path.join(HERE, "libb200pos.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in os.listdir(d)] + [os.path.join(HERE, "..", "include", "b200pos.h")]


def up_to_date():
    return os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--shared",
           "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-o", OUT, SRC, "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return OUT


ASAN_OUT = os.path.join(HERE, "libb200pos_asan.so")


def build_asan():
    """The same library with the HOST side (argument checks, marshaling, std::vector staging of the C ABI) instrumented by
    AddressSanitizer + UBSan; device code unchanged.  Test tooling (tools/sanitize.sh loads it through B2_LIB), never the product."""
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-lineinfo", "--shared", "-Xcompiler", "-fPIC",
           "-Xcompiler", "-fsanitize=address", "-Xcompiler", "-fsanitize=undefined", "-Xcompiler", "-fno-omit-frame-pointer", "-Xcompiler", "-g",
           "-o", ASAN_OUT, SRC, "-ldl", "-Xlinker", "-lasan", "-Xlinker", "-lubsan"]
    subprocess.check_call(cmd)
    return ASAN_OUT


if __name__ == "__main__":
    if "--asan" in sys.argv:
        print(build_asan())
    else:
        print(build(force=True, verbose="-v" in sys.argv))
