"""RFC 9380 hash_to_curve for BLS12381G2_XMD:SHA-256_SSWU_RO_ -- CPU oracle (TEST INFRASTRUCTURE).

The reference never contains this code: it is reached through
``is_valid_indexed_attestation`` (/root/reference/pos-evolution.md:736, :976) ->
``bls.FastAggregateVerify`` -> py_ecc ``hash_to_G2``.  Restated from RFC 9380 sections 5.2,
5.3.1, 6.6.2, 8.8.2 and appendix E.3; pinned by the RFC's K.1 and J.10.1 vectors
(tests/test_oracle_kat.py).
"""
import hashlib

from .bls12_381 import (P, E2, F2_ONE, F2_ZERO, H_EFF_G2, f2_add, f2_sub, f2_mul, f2_sqr, f2_neg,
                        f2_inv, f2_sqrt, f2_is_zero)

DST_POP = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"


def expand_message_xmd(msg: bytes, dst: bytes, len_in_bytes: int) -> bytes:
    b_in_bytes, s_in_bytes = 32, 64
    ell = (len_in_bytes + b_in_bytes - 1) // b_in_bytes
    if ell > 255 or len_in_bytes > 65535 or len(dst) > 255:
        raise ValueError("expand_message_xmd: bad lengths")
    dst_prime = dst + bytes([len(dst)])
    z_pad = bytes(s_in_bytes)
    l_i_b = len_in_bytes.to_bytes(2, "big")
    b0 = hashlib.sha256(z_pad + msg + l_i_b + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:len_in_bytes]


def hash_to_field_fp2(msg: bytes, count: int, dst: bytes):
    L, m = 64, 2
    uniform = expand_message_xmd(msg, dst, count * m * L)
    out = []
    for i in range(count):
        e = []
        for j in range(m):
            off = L * (j + i * m)
            e.append(int.from_bytes(uniform[off:off + L], "big") % P)
        out.append(tuple(e))
    return out


# SSWU parameters for the 3-isogenous curve E2': y^2 = x^3 + A'x + B'
ISO_A = (0, 240)
ISO_B = (1012, 1012)
SSWU_Z = (P - 2, P - 1)          # -(2 + i)


def sgn0_fp2(a):
    s0 = a[0] & 1
    z0 = a[0] == 0
    return s0 | (z0 & (a[1] & 1))


def map_to_curve_sswu(u):
    """Simplified SWU (RFC 9380 section 6.6.2, straight-line semantics) -> affine point on E2'."""
    u2 = f2_sqr(u)
    zu2 = f2_mul(SSWU_Z, u2)
    tv1 = f2_add(f2_sqr(zu2), zu2)                   # Z^2 u^4 + Z u^2
    if f2_is_zero(tv1):
        x1 = f2_mul(ISO_B, f2_inv(f2_mul(SSWU_Z, ISO_A)))
    else:
        x1 = f2_mul(f2_mul(f2_neg(ISO_B), f2_inv(ISO_A)), f2_add(F2_ONE, f2_inv(tv1)))
    gx1 = f2_add(f2_add(f2_mul(f2_sqr(x1), x1), f2_mul(ISO_A, x1)), ISO_B)
    y1 = f2_sqrt(gx1)
    if y1 is not None:
        x, y = x1, y1
    else:
        x = f2_mul(zu2, x1)
        gx2 = f2_add(f2_add(f2_mul(f2_sqr(x), x), f2_mul(ISO_A, x)), ISO_B)
        y = f2_sqrt(gx2)
        assert y is not None
    if sgn0_fp2(u) != sgn0_fp2(y):
        y = f2_neg(y)
    return x, y


# 3-isogeny E2' -> E2, RFC 9380 appendix E.3; coefficients low -> high degree, (re, im)
_K = 0x5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97d6
_L = 0x1530477c7ab4113b59a4c18b076d11930f7da5d4a07f649bf54439d87d27e500fc8c25ebf8c92f6812cfc71c71c6d706
ISO_XNUM = [
    (_K, _K),
    (0, 0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71a),
    (0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71e,
     0x8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38d),
    (0x171d6541fa38ccfaed6dea691f5fb614cb14b4e7f4e810aa22d6108f142b85757098e38d0f671c7188e2aaaaaaaa5ed1, 0),
]
ISO_XDEN = [(0, P - 72), (12, P - 12), (1, 0)]
ISO_YNUM = [
    (_L, _L),
    (0, 0x5c759507e8e333ebb5b7a9a47d7ed8532c52d39fd3a042a88b58423c50ae15d5c2638e343d9c71c6238aaaaaaaa97be),
    (0x11560bf17baa99bc32126fced787c88f984f87adf7ae0c7f9a208c6b4f20a4181472aaa9cb8d555526a9ffffffffc71c,
     0x8ab05f8bdd54cde190937e76bc3e447cc27c3d6fbd7063fcd104635a790520c0a395554e5c6aaaa9354ffffffffe38f),
    (0x124c9ad43b6cf79bfbf7043de3811ad0761b0f37a1e26286b0e977c69aa274524e79097a56dc4bd9e1b371c71c718b10, 0),
]
ISO_YDEN = [(P - 432, P - 432), (0, P - 216), (18, P - 18), (1, 0)]


def _horner(coeffs, x):
    acc = coeffs[-1]
    for c in reversed(coeffs[:-1]):
        acc = f2_add(f2_mul(acc, x), c)
    return acc


def iso_map_g2(x, y):
    """E2' affine -> E2 affine (None when a denominator vanishes = point at infinity)."""
    xd = _horner(ISO_XDEN, x)
    yd = _horner(ISO_YDEN, x)
    if f2_is_zero(xd) or f2_is_zero(yd):
        return None
    xn = _horner(ISO_XNUM, x)
    yn = _horner(ISO_YNUM, x)
    return f2_mul(xn, f2_inv(xd)), f2_mul(y, f2_mul(yn, f2_inv(yd)))


def clear_cofactor_g2(pj):
    return E2.mul(pj, H_EFF_G2)


def map_to_curve_g2(u):
    aff = iso_map_g2(*map_to_curve_sswu(u))
    return E2.INF if aff is None else E2.from_affine(*aff)


def hash_to_g2(msg: bytes, dst: bytes = DST_POP):
    """hash_to_curve (random-oracle variant) -> Jacobian point in the r-torsion of E2."""
    u0, u1 = hash_to_field_fp2(msg, 2, dst)
    q = E2.add(map_to_curve_g2(u0), map_to_curve_g2(u1))
    return clear_cofactor_g2(q)
