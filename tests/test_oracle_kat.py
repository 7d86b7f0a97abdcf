"""Pins the oracle to the public known-answer vectors of the standards it restates
(SURVEY.md appendix B): RFC 9380 K.1 / J.10.1, eth2 bls/sign vectors, SkToPk, ZCash encodings."""
from oracle import bls_sig as B
from oracle.bls12_381 import (P, R, X_ABS, H1, E1, E2, G1, G2, G1_X, G1_Y, G2_X, G2_Y, g1_compress, g2_compress,
                              g1_in_subgroup, g2_in_subgroup)
from oracle.hash_to_curve import expand_message_xmd, hash_to_g2, H_EFF_G2, map_to_curve_g2

SK = 0x263dbd792f5b1be47ed85f8938c0f29586af0d3ac7b977f21c278fe1462040e3


def test_curve_constants():
    x = -X_ABS
    assert R == x**4 - x**2 + 1
    assert P == (x - 1) ** 2 * R // 3 + x
    assert P % 4 == 3
    assert H1 == (x - 1) ** 2 // 3
    assert (P**12 - 1) % R == 0
    assert E1.on_curve_affine(G1_X, G1_Y) and E2.on_curve_affine(G2_X, G2_Y)
    assert g1_in_subgroup(G1) and g2_in_subgroup(G2)
    assert g1_compress(G1).hex().startswith("97f1d3a7")
    assert g2_compress(G2).hex().startswith("93e02b60")


def test_rfc9380_expand_message_xmd():
    out = expand_message_xmd(b"", b"QUUX-V01-CS02-with-expander-SHA256-128", 32)
    assert out.hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"
    out = expand_message_xmd(b"abc", b"QUUX-V01-CS02-with-expander-SHA256-128", 32)
    assert out.hex() == "d8ccab23b5985ccea865c6c97b6e5b8350e794e603b4b97902f53a8a0d605615"


def test_rfc9380_hash_to_g2():
    p = hash_to_g2(b"", b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_")
    x, y = E2.to_affine(p)
    assert x == (0x0141ebfbdca40eb85b87142e130ab689c673cf60f1a3e98d69335266f30d9b8d4ac44c1038e9dcdd5393faf5c41fb78a,
                 0x05cb8437535e20ecffaef7752baddf98034139c38452458baeefab379ba13dff5bf5dd71b72418717047f5b0f37da03d)
    assert y == (0x0503921d7f6a12805e72940b963c0cf3471c7b2a524950ca195d11062ee75ec076daf2d4bc358c4b190c0c98064fdd92,
                 0x12424ac32561493f3fe3c260708a12b7c620e7be00099a974e259ddc7d1f6395c3c811cdd19f1e8dbf3e9ecfdcbab8d6)


def test_h_eff_lands_in_subgroup():
    q = map_to_curve_g2((5, 7))
    assert not g2_in_subgroup(q)
    assert g2_in_subgroup(E2.mul(q, H_EFF_G2))


def test_eth2_sk_to_pk():
    assert B.SkToPk(SK).hex() == ("a491d1b0ecd9bb917989f0e74f0dea0422eac4a873e5e2644f368dffb9a6e20f"
                                  "d6e10c1b77654d067c0618f6e5a7f79a")


def test_eth2_sign_vectors():
    assert B.Sign(SK, b"\x56" * 32).hex() == (
        "882730e5d03f6b42c3abc26d3372625034e1d871b65a8a6b900a56dae22da98abbe1b68f85e49fe7652a55ec3d0591c2"
        "0767677e33e5cbb1207315c41a9ac03be39c2e7668edc043d6cb1d9fd93033caa8a1c5b0e84bedaeb6c64972503a43eb")
    assert B.Sign(SK, b"\xab" * 32).hex() == (
        "91347bccf740d859038fcdcaf233eeceb2a436bcaaee9b2aa3bfb70efe29dfb2677562ccbea1c8e061fb9971b0753c24"
        "0622fab78489ce96768259fc01360346da5b9f579e5da0d941e4c6ba18a0e64906082375394f337fa1af2b7127b0d121")
    assert B.Sign(SK, b"\x00" * 32).hex() == (
        "b6ed936746e01f8ecf281f020953fbf1f01debd5657c4a383940b020b26507f6076334f91e2366c96e9ab279fb515809"
        "0352ea1c5b0c9274504f4f0e7053af24802e51e4568d164fe986834f41e55c8e850ce1f98458c0cfc9ab380b55285a55")
