"""IETF BLS signatures, proof-of-possession ciphersuite, py_ecc semantics -- CPU oracle
(TEST INFRASTRUCTURE).

This is the ``bls`` facade the reference calls (``bls.Verify`` at
/root/reference/pos-evolution.md:165) and whose ``Aggregate`` / ``FastAggregateVerify``
BASELINE.json's north_star names.  The reference does not contain the code (SURVEY.md
section 8c); behaviour restated from draft-irtf-cfrg-bls-signature-05 sections 2.3-2.9, 3.3.4
and py_ecc ``G2ProofOfPossession`` (pyspec's default backend):

  * Verify / FastAggregateVerify never raise: any malformed input -> False;
  * Aggregate raises on an empty list or an undecodable signature, performs NO
    subgroup check;
  * KeyValidate = decodable, not infinity, in the r-torsion.
"""
from .bls12_381 import (R, E1, E2, G1, g1_compress, g1_decompress, g2_compress, g2_decompress,
                        g1_in_subgroup, g2_in_subgroup, pairing_product_is_one, DeserializationError)
from .hash_to_curve import hash_to_g2, DST_POP


def SkToPk(sk: int) -> bytes:
    if not 0 < sk < R:
        raise ValueError("secret key out of range")
    return g1_compress(E1.mul(G1, sk))


def Sign(sk: int, message: bytes) -> bytes:
    if not 0 < sk < R:
        raise ValueError("secret key out of range")
    return g2_compress(E2.mul(hash_to_g2(message, DST_POP), sk))


def KeyValidate(pk: bytes) -> bool:
    try:
        p = g1_decompress(bytes(pk))
    except DeserializationError:
        return False
    return (not E1.is_inf(p)) and g1_in_subgroup(p)


def _core_verify_point(pk_point, message: bytes, signature: bytes) -> bool:
    try:
        sig = g2_decompress(bytes(signature))
    except DeserializationError:
        return False
    if not g2_in_subgroup(sig):
        return False
    h = hash_to_g2(bytes(message), DST_POP)
    # e(pk, H(m)) * e(-g1, sig) == 1
    return pairing_product_is_one([(pk_point, h), (E1.neg(G1), sig)])


def Verify(pk: bytes, message: bytes, signature: bytes) -> bool:
    try:
        if not KeyValidate(pk):
            return False
        return _core_verify_point(g1_decompress(bytes(pk)), message, signature)
    except Exception:
        return False


def Aggregate(signatures) -> bytes:
    if len(signatures) < 1:
        raise ValueError("Aggregate: empty list")
    acc = E2.INF
    for s in signatures:
        acc = E2.add(acc, g2_decompress(bytes(s)))        # raises if undecodable
    return g2_compress(acc)


def AggregatePKs(pubkeys) -> bytes:
    """py_ecc ``_AggregatePKs``: every key must pass KeyValidate."""
    if len(pubkeys) < 1:
        raise ValueError("AggregatePKs: empty list")
    acc = E1.INF
    for pk in pubkeys:
        if not KeyValidate(pk):
            raise ValueError("AggregatePKs: invalid key")
        acc = E1.add(acc, g1_decompress(bytes(pk)))
    return g1_compress(acc)


def FastAggregateVerify(pubkeys, message: bytes, signature: bytes) -> bool:
    try:
        if len(pubkeys) < 1:
            return False
        acc = E1.INF
        for pk in pubkeys:
            if not KeyValidate(pk):
                return False
            acc = E1.add(acc, g1_decompress(bytes(pk)))
        if E1.is_inf(acc):                                  # KeyValidate(aggregate) fails on infinity
            return False
        return _core_verify_point(acc, message, signature)
    except Exception:
        return False


def AggregateVerify(pubkeys, messages, signature: bytes) -> bool:
    try:
        if len(pubkeys) < 1 or len(pubkeys) != len(messages):
            return False
        sig = g2_decompress(bytes(signature))
        if not g2_in_subgroup(sig):
            return False
        pairs = []
        for pk, m in zip(pubkeys, messages):
            if not KeyValidate(pk):
                return False
            pairs.append((g1_decompress(bytes(pk)), hash_to_g2(bytes(m), DST_POP)))
        pairs.append((E1.neg(G1), sig))
        return pairing_product_is_one(pairs)
    except Exception:
        return False
