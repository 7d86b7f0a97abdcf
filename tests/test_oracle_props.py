"""Algebraic property tests of the oracle."""
import pytest

from oracle import bls_sig as B
from oracle.bls12_381 import (P, R, E1, E2, G1, G2, F12_ONE, f12_pow, f12_frob, pairing, miller_loop,
                              final_exponentiation, g1_compress, g1_decompress, g2_compress, g2_decompress,
                              DeserializationError, f2_sqrt, f2_sqr)


def test_pairing_bilinear_and_order():
    e = pairing(G1, G2)
    assert e != F12_ONE
    assert f12_pow(e, R) == F12_ONE
    assert pairing(E1.mul(G1, 7), E2.mul(G2, 11)) == f12_pow(e, 77)
    assert pairing(E1.neg(G1), G2) == f12_pow(e, R - 1)


def test_final_exp_matches_plain_power():
    f = miller_loop(E1.to_affine(E1.mul(G1, 3)), E2.to_affine(E2.mul(G2, 5)))
    assert final_exponentiation(f) == f12_pow(f, 3 * (P**12 - 1) // R)
    assert f12_frob(f) == f12_pow(f, P)


def test_compress_roundtrip_and_flags():
    for k in (1, 2, 3, 0xdeadbeef, R - 1):
        p, q = E1.mul(G1, k), E2.mul(G2, k)
        assert E1.eq(g1_decompress(g1_compress(p)), p)
        assert E2.eq(g2_decompress(g2_compress(q)), q)
    assert E1.is_inf(g1_decompress(bytes([0xC0]) + bytes(47)))
    assert E2.is_inf(g2_decompress(bytes([0xC0]) + bytes(95)))
    for bad in (bytes(48), bytes([0xE0]) + bytes(47), bytes([0xC0]) + bytes(46) + b"\x01",
                bytes([0x9F]) + b"\xff" * 47):
        with pytest.raises(DeserializationError):
            g1_decompress(bad)
    with pytest.raises(DeserializationError):
        g2_decompress(bytes(96))


def test_fp2_sqrt():
    for a in ((3, 4), (0, 5), (7, 0), (P - 1, 0), (123456789, 987654321)):
        sq = f2_sqr(a)
        r = f2_sqrt(sq)
        assert r is not None and f2_sqr(r) == sq


def test_sign_verify_aggregate_semantics():
    m = b"\x11" * 32
    sks = [1000 + i for i in range(5)]
    pks = [B.SkToPk(k) for k in sks]
    sigs = [B.Sign(k, m) for k in sks]
    assert all(B.Verify(pk, m, s) for pk, s in zip(pks, sigs))
    agg = B.Aggregate(sigs)
    assert agg == B.Aggregate(list(reversed(sigs)))
    assert B.Aggregate([B.Aggregate(sigs[:2]), B.Aggregate(sigs[2:])]) == agg
    assert B.FastAggregateVerify(pks, m, agg)
    assert B.FastAggregateVerify(pks, m, agg) == B.Verify(B.AggregatePKs(pks), m, agg)
    assert not B.FastAggregateVerify(pks, b"\x12" * 32, agg)
    assert not B.FastAggregateVerify(pks[:-1], m, agg)
    assert not B.FastAggregateVerify([], m, agg)
    assert not B.FastAggregateVerify(pks, m, bytes([0xC0]) + bytes(95))          # infinity signature
    assert not B.FastAggregateVerify(pks + [bytes([0xC0]) + bytes(47)], m, agg)    # infinity pubkey
    assert not B.FastAggregateVerify(pks, m, bytes(96))                            # undecodable
    neg = g1_compress(E1.neg(g1_decompress(pks[0])))
    assert not B.FastAggregateVerify([pks[0], neg], m, agg)                        # aggregate pk = infinity
    with pytest.raises(ValueError):
        B.Aggregate([])
