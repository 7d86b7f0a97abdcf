"""ctypes access to tests/hostsim/libhostsim.so plus Montgomery marshaling helpers (test-only)."""
import ctypes
import importlib.util
import os

import numpy as np

from oracle.bls12_381 import P

HERE = os.path.dirname(os.path.abspath(__file__))
RMONT = 1 << 384
RINV = pow(RMONT, -1, P)


def load():
    spec = importlib.util.spec_from_file_location("hostsim_build", os.path.join(HERE, "hostsim", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return ctypes.CDLL(mod.build())


def limbs(v, n=12):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def fp_m(v):
    """int -> Montgomery limbs (np.uint32[12])."""
    return np.array(limbs(v * RMONT % P), dtype=np.uint32)


def fp_v(a):
    """Montgomery limbs -> int."""
    return sum(int(x) << (32 * i) for i, x in enumerate(a[:12])) * RINV % P


def fp2_m(v):
    return np.concatenate([fp_m(v[0]), fp_m(v[1])])


def fp2_v(a):
    return (fp_v(a[0:12]), fp_v(a[12:24]))


def fp12_m(f):
    out = []
    for six in f:
        for two in six:
            out.append(fp2_m(two))
    return np.concatenate(out)


def fp12_v(a):
    g = [fp2_v(a[24 * k:24 * k + 24]) for k in range(6)]
    return ((g[0], g[1], g[2]), (g[3], g[4], g[5]))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def buf(b):
    return (ctypes.c_uint8 * len(b)).from_buffer_copy(bytes(b))
