"""Synthetic committees made with the CPU oracle only (TEST INFRASTRUCTURE / CPU-baseline input).

Used by bench.py's `--impl reference` arm and by tests that need a full-size committee without the
GPU: keys are an arithmetic progression sk_i = sk_0 + i*delta, so pubkeys and individual signatures
are generated with one point addition each plus a batch inversion."""
import hashlib

from .bls12_381 import R, E1, E2, G1, g1_compress, g2_compress
from .hash_to_curve import hash_to_g2


def _h(b):
    return hashlib.sha256(b).digest()


def committee(tag: int, size: int = 512):
    """-> (pubkeys [48 B] * size, individual signatures [96 B] * size, message 32 B), all valid."""
    sk0 = int.from_bytes(_h(b"b200pos/synth/sk0" + tag.to_bytes(8, "little")), "big") % R or 1
    delta = int.from_bytes(_h(b"b200pos/synth/skd"), "big") % R or 1
    msg = _h(b"b200pos/synth/msg" + tag.to_bytes(8, "little"))
    h = hash_to_g2(msg)
    p, pstep = E1.mul(G1, sk0), E1.mul(G1, delta)
    s, sstep = E2.mul(h, sk0), E2.mul(h, delta)
    pts1, pts2 = [], []
    for _ in range(size):
        pts1.append(p)
        pts2.append(s)
        p = E1.add(p, pstep)
        s = E2.add(s, sstep)
    pks = [g1_compress((x, y, 1)) for (x, y) in E1.batch_to_affine(pts1)]
    sigs = [g2_compress((a[0], a[1], (1, 0))) for a in E2.batch_to_affine(pts2)]
    return pks, sigs, msg
