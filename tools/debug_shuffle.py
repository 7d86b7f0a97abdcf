import ctypes, hashlib, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pos_evolution_b200.engine import Engine
from oracle import fast
eng = Engine(0)
lib = eng.lib
for n, rounds in ((2, 1), (300, 2)):
    seed = hashlib.sha256(b"shuffle" + n.to_bytes(4, "little")).digest()
    ref = fast.shuffle_permutation(n, seed, rounds)
    got = eng.shuffle_committees(seed, n, rounds)
    nblk = (n + 255) // 256
    src = np.zeros(rounds * nblk * 32, dtype=np.uint8)
    piv = np.zeros(rounds, dtype=np.uint64)
    lib.b2_debug_shuffle_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_uint32]
    rc = lib.b2_debug_shuffle_tables(eng.h, src.ctypes.data_as(ctypes.c_void_p), src.size, piv.ctypes.data_as(ctypes.c_void_p), rounds)
    print("n", n, "rounds", rounds, "rc", rc, "equal", np.array_equal(ref, got))
    for r in range(rounds):
        hp = int.from_bytes(hashlib.sha256(seed + bytes([r])).digest()[:8], "little") % n
        print("  round", r, "pivot gpu", int(piv[r]), "host", hp)
        for blk in range(nblk):
            h = hashlib.sha256(seed + bytes([r]) + blk.to_bytes(4, "little")).digest()
            g = bytes(src[(r * nblk + blk) * 32:(r * nblk + blk) * 32 + 32])
            print("   blk", blk, "src equal", g == h, g[:6].hex(), h[:6].hex())
# 33-byte hash through the batch API for comparison
seed = hashlib.sha256(b"shuffle" + (2).to_bytes(4, "little")).digest()
m = np.frombuffer(seed + b"\x00", dtype=np.uint8).reshape(1, 33)
print("batch33", bytes(eng.sha256_batch(m, 33)[0]).hex()[:16], hashlib.sha256(seed + b"\x00").hexdigest()[:16])
