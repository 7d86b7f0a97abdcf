"""The CUDA library's math headers, compiled for the host (tests/hostsim), against the oracle.
This is how the algorithm layer is debugged without a GPU; the GPU tests (test_gpu_*.py) then only
have to establish that the PTX carry chains and the kernel orchestration agree with it."""
import ctypes
import hashlib
import random

import numpy as np
import pytest

import hs
from oracle import bls_sig as B
from oracle.bls12_381 import (P, R, E1, E2, G1, G2, F12_ONE, f2_mul, f2_sqr, f2_inv, f2_sqrt, f12_mul, f12_sqr, f12_inv,
                              f12_frob, f12_frob_n, f12_conj, f12_pow, miller_loop, final_exponentiation, pairing,
                              g1_compress, g2_compress, g1_decompress, g2_decompress, g2_in_subgroup, f2_add)
from oracle.hash_to_curve import (expand_message_xmd, hash_to_g2, map_to_curve_sswu, clear_cofactor_g2, DST_POP, H_EFF_G2,
                                  map_to_curve_g2)

lib = hs.load()
rnd = random.Random(11)


def rfp():
    return rnd.randrange(P)


def rf2():
    return (rfp(), rfp())


def rf12():
    return tuple(tuple(rf2() for _ in range(3)) for _ in range(2))


def call(fn, *arrs, out_words):
    out = np.zeros(out_words, dtype=np.uint32)
    fn(*[hs.ptr(a) for a in arrs], hs.ptr(out))
    return out


def test_fp_ops_random_and_edges():
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (P + 1) // 2, (1 << 380), hs.RMONT % P]
    pairs = [(a, b) for a in edge for b in edge] + [(rfp(), rfp()) for _ in range(3000)]
    for a, b in pairs:
        am, bm = hs.fp_m(a), hs.fp_m(b)
        assert hs.fp_v(call(lib.hs_fp_mul, am, bm, out_words=12)) == a * b % P
        assert hs.fp_v(call(lib.hs_fp_add, am, bm, out_words=12)) == (a + b) % P
        assert hs.fp_v(call(lib.hs_fp_sub, am, bm, out_words=12)) == (a - b) % P
        assert hs.fp_v(call(lib.hs_fp_neg, am, out_words=12)) == -a % P
    for _ in range(10):
        a = rfp()
        assert hs.fp_v(call(lib.hs_fp_inv, hs.fp_m(a), out_words=12)) == pow(a, -1, P)


def test_fp2_ops():
    for _ in range(300):
        a, b = rf2(), rf2()
        assert hs.fp2_v(call(lib.hs_fp2_mul, hs.fp2_m(a), hs.fp2_m(b), out_words=24)) == f2_mul(a, b)
        assert hs.fp2_v(call(lib.hs_fp2_sqr, hs.fp2_m(a), out_words=24)) == f2_sqr(a)
    for _ in range(5):
        a = rf2()
        assert hs.fp2_v(call(lib.hs_fp2_inv, hs.fp2_m(a), out_words=24)) == f2_inv(a)


def test_fp2_sqrt_all_branches():
    cases = [rf2() for _ in range(24)] + [(rfp(), 0) for _ in range(4)] + [(0, rfp()) for _ in range(2)] + [(0, 0), (1, 0), (P - 1, 0)]
    n_sq = n_non = 0
    for a in cases:
        for v in (f2_sqr(a), a):
            out = np.zeros(24, dtype=np.uint32)
            ok = lib.hs_fp2_sqrt(hs.ptr(hs.fp2_m(v)), hs.ptr(out))
            assert bool(ok) == (f2_sqrt(v) is not None)
            if ok:
                assert f2_sqr(hs.fp2_v(out)) == v
                n_sq += 1
            else:
                n_non += 1
    assert n_sq > 20 and n_non > 5


def test_fp12_ops():
    for _ in range(10):
        a, b = rf12(), rf12()
        assert hs.fp12_v(call(lib.hs_fp12_mul, hs.fp12_m(a), hs.fp12_m(b), out_words=144)) == f12_mul(a, b)
        assert hs.fp12_v(call(lib.hs_fp12_sqr, hs.fp12_m(a), out_words=144)) == f12_sqr(a)
        assert hs.fp12_v(call(lib.hs_fp12_frob, hs.fp12_m(a), out_words=144)) == f12_frob(a)
        assert hs.fp12_v(call(lib.hs_fp12_frob2, hs.fp12_m(a), out_words=144)) == f12_frob_n(a, 2)
        l0, l1, l4 = rf2(), rf2(), rf2()
        sparse = ((l0, l1, (0, 0)), ((0, 0), l4, (0, 0)))
        got = call(lib.hs_fp12_mul_by_014, hs.fp12_m(a), hs.fp2_m(l0), hs.fp2_m(l1), hs.fp2_m(l4), out_words=144)
        assert hs.fp12_v(got) == f12_mul(a, sparse)
    a = rf12()
    assert hs.fp12_v(call(lib.hs_fp12_inv, hs.fp12_m(a), out_words=144)) == f12_inv(a)
    # cyclotomic squaring is only valid after the easy part of the final exponentiation
    t = f12_mul(f12_conj(a), f12_inv(a))
    m = f12_mul(f12_frob_n(t, 2), t)
    assert hs.fp12_v(call(lib.hs_fp12_cyc_sqr, hs.fp12_m(m), out_words=144)) == f12_sqr(m)


def _k8(k):
    return np.array(hs.limbs(k, 8), dtype=np.uint32)


def test_g1_g2_group_ops_and_encodings():
    inf1, inf2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)
    for k1, k2 in ((1, 1), (1, 2), (5, R - 5), (123456789, 987654321), (R - 1, 1), (7, 7)):
        for mixed in (0, 1):
            o = ctypes.create_string_buffer(48)
            assert lib.hs_g1_add(hs.buf(g1_compress(E1.mul(G1, k1))), hs.buf(g1_compress(E1.mul(G1, k2))), o, mixed) == 0
            assert o.raw == g1_compress(E1.mul(G1, (k1 + k2) % R))
            o = ctypes.create_string_buffer(96)
            assert lib.hs_g2_add(hs.buf(g2_compress(E2.mul(G2, k1))), hs.buf(g2_compress(E2.mul(G2, k2))), o, mixed) == 0
            assert o.raw == g2_compress(E2.mul(G2, (k1 + k2) % R))
    o = ctypes.create_string_buffer(48)
    lib.hs_g1_add(hs.buf(inf1), hs.buf(g1_compress(G1)), o, 1)
    assert o.raw == g1_compress(G1)
    lib.hs_g1_add(hs.buf(g1_compress(G1)), hs.buf(inf1), o, 0)
    assert o.raw == g1_compress(G1)
    for k in (1, 2, 0xdeadbeefcafebabe, R - 1, rnd.randrange(R)):
        o = ctypes.create_string_buffer(48)
        lib.hs_g1_mul(hs.buf(g1_compress(G1)), hs.ptr(_k8(k)), o)
        assert o.raw == g1_compress(E1.mul(G1, k))
    k = rnd.randrange(R)
    o = ctypes.create_string_buffer(96)
    lib.hs_g2_mul(hs.buf(g2_compress(G2)), hs.ptr(_k8(k)), o)
    assert o.raw == g2_compress(E2.mul(G2, k))
    # decode errors
    bad1 = [bytes(48), bytes([0xE0]) + bytes(47), bytes([0xC0]) + bytes(46) + b"\x01", bytes([0x9F]) + b"\xff" * 47]
    out = np.zeros(24, dtype=np.uint32)
    for b in bad1:
        assert lib.hs_g1_decompress(hs.buf(b), hs.ptr(out)) == 2
    assert lib.hs_g1_decompress(hs.buf(inf1), hs.ptr(out)) == 1
    # x not on curve
    x = 1
    while True:
        try:
            g1_decompress(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big"))
            x += 1
        except Exception:
            break
    assert lib.hs_g1_decompress(hs.buf(bytes([0x80]) + x.to_bytes(47, "big")), hs.ptr(out)) == 2
    out2 = np.zeros(48, dtype=np.uint32)
    assert lib.hs_g2_decompress(hs.buf(bytes(96)), hs.ptr(out2)) == 2
    assert lib.hs_g2_decompress(hs.buf(inf2), hs.ptr(out2)) == 1
    assert lib.hs_g2_decompress(hs.buf(g2_compress(G2)), hs.ptr(out2)) == 0


def _aff_m(field_m, aff):
    return np.concatenate([field_m(aff[0]), field_m(aff[1])])


def test_subgroup_checks():
    for k in (1, 5, rnd.randrange(R)):
        assert lib.hs_g1_in_subgroup_exact(hs.ptr(_aff_m(hs.fp_m, E1.to_affine(E1.mul(G1, k))))) == 1
        a2 = _aff_m(hs.fp2_m, E2.to_affine(E2.mul(G2, k)))
        assert lib.hs_g2_in_subgroup_exact(hs.ptr(a2)) == 1 and lib.hs_g2_in_subgroup_psi(hs.ptr(a2)) == 1
    # points on the curves but outside the r-torsion
    x, found = 1, 0
    while found < 3:
        try:
            p = g1_decompress(bytes([0x80]) + x.to_bytes(47, "big"))
            if not E1.is_inf(E1.mul(p, R)):
                assert lib.hs_g1_in_subgroup_exact(hs.ptr(_aff_m(hs.fp_m, E1.to_affine(p)))) == 0
                found += 1
        except Exception:
            pass
        x += 1
    for u in ((1, 2), (3, 4), (5, 6), (7, 8), (9, 10), (11, 12)):
        q = map_to_curve_g2(u)                      # on E2, not in G2 before cofactor clearing
        assert not g2_in_subgroup(q)
        a2 = _aff_m(hs.fp2_m, E2.to_affine(q))
        assert lib.hs_g2_in_subgroup_exact(hs.ptr(a2)) == 0 and lib.hs_g2_in_subgroup_psi(hs.ptr(a2)) == 0
        o = ctypes.create_string_buffer(96)
        lib.hs_g2_clear_cofactor(hs.ptr(a2), o)
        assert o.raw == g2_compress(E2.mul(q, H_EFF_G2))
    # small-order / cofactor-torsion points must also be rejected by the psi test: take a point of E2 and kill its G2 part
    q = E2.mul(map_to_curve_g2((13, 14)), R)        # order divides the cofactor
    assert not E2.is_inf(q)
    a2 = _aff_m(hs.fp2_m, E2.to_affine(q))
    assert lib.hs_g2_in_subgroup_psi(hs.ptr(a2)) == 0


def test_sha256_and_expand_message():
    for n in (0, 1, 55, 56, 63, 64, 65, 200):
        m = bytes(range(256))[:n] * 1
        o = ctypes.create_string_buffer(32)
        lib.hs_sha256(hs.buf(m) if n else None, n, o)
        assert o.raw == hashlib.sha256(m).digest()
    for msg, dst in ((b"", b"QUUX-V01-CS02-with-expander-SHA256-128"), (b"abcdef0123456789", DST_POP), (bytes(32), DST_POP)):
        o = ctypes.create_string_buffer(256)
        lib.hs_expand_message_xmd_256(hs.buf(msg) if msg else None, len(msg), hs.buf(dst), len(dst), o)
        assert o.raw == expand_message_xmd(msg, dst, 256)


def test_sswu_map_matches_rfc_semantics():
    for u in [rf2() for _ in range(8)] + [(0, 0), (1, 0), (0, 1)]:
        out = np.zeros(48, dtype=np.uint32)
        lib.hs_sswu_map(hs.ptr(hs.fp2_m(u)), hs.ptr(out))
        assert (hs.fp2_v(out[:24]), hs.fp2_v(out[24:])) == map_to_curve_sswu(u)


def test_hash_to_g2_rfc_vector_and_random():
    o = ctypes.create_string_buffer(96)
    dst = b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_"
    lib.hs_hash_to_g2(None, 0, hs.buf(dst), len(dst), o)
    assert o.raw == g2_compress(hash_to_g2(b"", dst))
    for i in range(4):
        m = hashlib.sha256(bytes([i])).digest()
        lib.hs_hash_to_g2(hs.buf(m), 32, hs.buf(DST_POP), len(DST_POP), o)
        assert o.raw == g2_compress(hash_to_g2(m, DST_POP))


def test_pairing_bit_exact_vs_oracle():
    for a, b in ((1, 1), (3, 5), (rnd.randrange(R), rnd.randrange(R))):
        pj, qj = E1.mul(G1, a), E2.mul(G2, b)
        out = np.zeros(144, dtype=np.uint32)
        assert lib.hs_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(out), 1, 0) == 0
        assert hs.fp12_v(out) == pairing(pj, qj)
        # Jacobian P with Z != 1 (the aggregated-pubkey form): 3P
        assert lib.hs_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(out), 1, 1) == 0
        assert hs.fp12_v(out) == pairing(E1.mul(pj, 3), qj)
    f = rf12()
    assert hs.fp12_v(call(lib.hs_final_exp, hs.fp12_m(f), out_words=144)) == final_exponentiation(f)
    out = np.zeros(144, dtype=np.uint32)
    lib.hs_pairing(hs.buf(bytes([0xC0]) + bytes(47)), hs.buf(g2_compress(G2)), hs.ptr(out), 1, 0)
    assert hs.fp12_v(out) == F12_ONE


def test_fp_sqr_dedicated():
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 380), (1 << 381) - 1 - ((1 << 381) - 1 >= P) * ((1 << 381) - P), hs.RMONT % P,
            int("ffffffff" * 11 + "0", 16) % P, int("0fffffffffffffff" * 6, 16) % P]
    for a in edge + [rfp() for _ in range(5000)]:
        assert hs.fp_v(call(lib.hs_fp_sqr, hs.fp_m(a), out_words=12)) == a * a % P
    # raw limb patterns (the Montgomery image is what the chains see): all-ones limbs below p, alternating, sparse
    half = (1 << 192) - 1
    for raw in (P - 1, P - 2, int("ffffffff" * 11, 16), int("ffffffff00000000" * 5 + "ffffffff", 16), 1 << 352, (1 << 352) - 1,
                # the 6-limb halves a0, a1 of the Karatsuba split: equal, zero, all-ones, a0 < a1, a0 > a1 by one
                half, half << 192, (0x1234 << 192) | 0x1234, (7 << 192) | 6, (6 << 192) | 7, ((half >> 8) << 192) | half, (1 << 192) | half, 1 << 192, 1):
        raw %= P
        arr = np.array(hs.limbs(raw), dtype=np.uint32)
        got = call(lib.hs_fp_sqr, arr, out_words=12)
        assert sum(int(x) << (32 * i) for i, x in enumerate(got)) == raw * raw * hs.RINV % P


def test_team_pairing_matches_oracle_and_serial():
    """3-lane cooperative Miller loop + final exponentiation (csrc/team.cuh, three host threads) == oracle pairing."""
    for a, b in ((1, 1), (5, 7), (rnd.randrange(R), rnd.randrange(R))):
        pj, qj = E1.mul(G1, a), E2.mul(G2, b)
        out = np.zeros(144, dtype=np.uint32)
        assert lib.hs_team_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(out), 1, 0) == 0
        assert hs.fp12_v(out) == pairing(pj, qj)
        assert lib.hs_team_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(out), 1, 1) == 0
        assert hs.fp12_v(out) == pairing(E1.mul(pj, 3), qj)
        # Miller value alone must equal the serial device code bit for bit
        o1, o2 = np.zeros(144, dtype=np.uint32), np.zeros(144, dtype=np.uint32)
        lib.hs_team_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(o1), 0, 1)
        lib.hs_pairing(hs.buf(g1_compress(pj)), hs.buf(g2_compress(qj)), hs.ptr(o2), 0, 1)
        assert np.array_equal(o1, o2)
    f = rf12()
    assert hs.fp12_v(call(lib.hs_team_final_exp, hs.fp12_m(f), out_words=144)) == final_exponentiation(f)
    out = np.zeros(144, dtype=np.uint32)
    lib.hs_team_pairing(hs.buf(g1_compress(G1)), hs.buf(bytes([0xC0]) + bytes(95)), hs.ptr(out), 1, 0)
    assert hs.fp12_v(out) == F12_ONE
