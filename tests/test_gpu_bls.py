"""GPU parity tests for the BLS half of the path: every result that crosses the C ABI is compared
byte for byte with the oracle (compressed points, verdict bytes).  Run on the B200 box:
    python -m pytest tests -m gpu"""
import hashlib

import numpy as np
import pytest

import scenarios
from oracle import bls_sig as B
from oracle import spec as S
from oracle.bls12_381 import R, E1, E2, G1, G2, g1_compress, g2_compress, g1_decompress
from oracle.hash_to_curve import hash_to_g2, map_to_curve_g2

pytestmark = pytest.mark.gpu

INF1 = bytes([0xC0]) + bytes(47)
INF2 = bytes([0xC0]) + bytes(95)


@pytest.fixture(scope="module")
def eng():
    from pos_evolution_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def world(eng):
    """64-validator minimal-preset state, registry loaded on the GPU (plus 4 special registry rows)."""
    pks = scenarios.pubkeys(64)
    spec, state = scenarios.minimal_state(64, slot=9, pks=pks)
    neg0 = g1_compress(E1.neg(g1_decompress(pks[0])))
    # a point on E1 outside the r-torsion
    x = 1
    while True:
        try:
            p = g1_decompress(bytes([0x80]) + x.to_bytes(47, "big"))
            if not E1.is_inf(E1.mul(p, R)):
                break
        except Exception:
            pass
        x += 1
    extra = [neg0, INF1, bytes(48), g1_compress(p)]      # 64: -pk0 (valid), 65: infinity, 66: undecodable, 67: not in G1
    allpk = np.frombuffer(b"".join(pks + extra), dtype=np.uint8).reshape(-1, 48)
    eff = np.array([v.effective_balance for v in state.validators] + [32 * 10**9] * 4, dtype=np.uint64)
    valid = eng.registry_load(allpk, eff)
    assert valid.tolist() == [1] * 65 + [0, 0, 0]
    return spec, state, pks


def _bits(rows, stride=None):
    n = max(len(r) for r in rows)
    stride = stride or max(1, (n + 7) // 8)
    out = np.zeros((len(rows), stride), dtype=np.uint8)
    for a, r in enumerate(rows):
        for j, b in enumerate(r):
            if b:
                out[a, j >> 3] |= 1 << (j & 7)
    return out


def test_sk_to_pk_and_sign(eng):
    sks = [1, 2, 0x263dbd792f5b1be47ed85f8938c0f29586af0d3ac7b977f21c278fe1462040e3, R - 1, scenarios.secret_key(5)]
    pk = eng.sk_to_pk(sks)
    for i, k in enumerate(sks):
        assert bytes(pk[i]) == B.SkToPk(k)
    msgs = np.frombuffer(b"\x56" * 32 + b"\xab" * 32 + b"\x00" * 32, dtype=np.uint8).reshape(3, 32)
    sig = eng.sign([sks[2]] * 3 + [sks[4]], [0, 1, 2, 1], msgs)
    assert bytes(sig[0]).hex().startswith("882730e5d03f6b42")       # eth2 bls/sign vectors
    assert bytes(sig[1]).hex().startswith("91347bccf740d859")
    assert bytes(sig[2]).hex().startswith("b6ed936746e01f8e")
    for i, (k, m) in enumerate(((sks[2], 0), (sks[2], 1), (sks[2], 2), (sks[4], 1))):
        assert bytes(sig[i]) == B.Sign(k, bytes(msgs[m]))


def test_hash_to_g2(eng):
    msgs = [hashlib.sha256(bytes([i])).digest() for i in range(40)]
    out = eng.hash_to_g2(np.frombuffer(b"".join(msgs), dtype=np.uint8))
    for i, m in enumerate(msgs):
        assert bytes(out[i]) == g2_compress(hash_to_g2(m))


def test_g1_aggregate_all_minimal_committees(eng, world):
    spec, state, pks = world
    members, off, rows = [], [0], []
    rng = np.random.default_rng(5)
    for slot in range(8, 16):
        for idx in range(2):
            c = spec.get_beacon_committee(state, slot, idx)
            members += c
            off.append(len(members))
            rows.append([bool(b) for b in rng.integers(0, 2, size=len(c))] if slot % 2 else [True] * len(c))
    # edge cases: empty bits; pk0 + (-pk0) -> infinity; a selected invalid key; unselected invalid key is harmless
    for mem, row in (([0, 1, 2, 3], [False] * 4), ([0, 64], [True, True]), ([1, 66, 2], [True, True, True]), ([1, 65, 2], [True, False, True]),
                     ([5, 67], [True, True])):
        members += mem
        off.append(len(members))
        rows.append(row)
    out, status = eng.g1_aggregate(members, off, _bits(rows))
    for a, row in enumerate(rows):
        sel = [members[off[a] + j] for j, b in enumerate(row) if b]
        acc = E1.INF
        bad = False
        for v in sel:
            if v >= 65:
                bad = True
                continue
            acc = E1.add(acc, g1_decompress(pks[v] if v < 64 else g1_compress(E1.neg(g1_decompress(pks[0])))))
        exp_status = (1 if bad else 0) | (2 if not sel else 0) | (4 if E1.is_inf(acc) else 0)
        assert status[a] == exp_status, a
        assert bytes(out[a]) == g1_compress(acc), a


def test_bls_aggregate_segments(eng, world):
    m = hashlib.sha256(b"agg").digest()
    sigs = scenarios.individual_signatures(list(range(10)), m)
    junk = bytes(96)
    not_g2 = g2_compress(map_to_curve_g2((1, 2)))          # decodable, outside G2: Aggregate does NOT check the subgroup
    all_sigs = sigs + [INF2, junk, not_g2]
    segs = [list(range(10)), [0], [3, 4, 5], [], [0, 10], [1, 11], [2, 12], [10, 10]]
    flat, off = [], [0]
    for s in segs:
        flat += [all_sigs[i] for i in s]
        off.append(len(flat))
    out, status = eng.aggregate(np.frombuffer(b"".join(flat), dtype=np.uint8), off)
    for k, s in enumerate(segs):
        try:
            exp = B.Aggregate([all_sigs[i] for i in s])
            assert status[k] == 0 and bytes(out[k]) == exp, k
        except Exception:
            assert status[k] == (2 if not s else 1), k


def test_fast_aggregate_verify_vs_oracle(eng, world):
    spec, state, pks = world
    atts = [
        scenarios.make_attestation(spec, state, 8, 0),
        scenarios.make_attestation(spec, state, 8, 1, bits=[True, False, True, True]),
        scenarios.make_attestation(spec, state, 5, 1),
        scenarios.make_attestation(spec, state, 8, 0, corrupt="flip_bit"),
        scenarios.make_attestation(spec, state, 8, 0, corrupt="wrong_message"),
        scenarios.make_attestation(spec, state, 8, 1, corrupt="wrong_signer_set"),
        scenarios.make_attestation(spec, state, 8, 0, bits=[False] * 4),
        scenarios.make_attestation(spec, state, 6, 1, bits=[False, False, True, False]),
    ]
    members, off, rows, msgs, sigs, expect = [], [0], [], [], [], []
    for att in atts:
        c = spec.get_beacon_committee(state, att.data.slot, att.data.index)
        members += c
        off.append(len(members))
        rows.append(att.aggregation_bits)
        dom = spec.get_domain(state, S.DOMAIN_BEACON_ATTESTER, att.data.target.epoch)
        msgs.append(spec.compute_signing_root(att.data, dom))
        sigs.append(att.signature)
        sel = [v for v, b in zip(c, att.aggregation_bits) if b]
        expect.append(B.FastAggregateVerify([pks[v] for v in sel], msgs[-1], att.signature))
    # signature-level edge cases on a valid committee
    c = spec.get_beacon_committee(state, 8, 0)
    good = atts[0]
    for sig in (INF2, bytes(96), g2_compress(map_to_curve_g2((3, 4)))):
        members += c
        off.append(len(members))
        rows.append([True] * 4)
        msgs.append(msgs[0])
        sigs.append(sig)
        expect.append(B.FastAggregateVerify([pks[v] for v in c], msgs[0], sig))
    # pubkey-level edge cases: aggregate = infinity, invalid key selected
    for mem in ([0, 64], [1, 66]):
        members += mem
        off.append(len(members))
        rows.append([True, True])
        msgs.append(msgs[0])
        sigs.append(good.signature)
        expect.append(False)
    assert expect[:3] == [True, True, True] and not any(expect[3:7]) and expect[7] is True and not any(expect[8:])
    ok = eng.fast_aggregate_verify(members, off, _bits(rows), np.frombuffer(b"".join(msgs), dtype=np.uint8),
                                   np.frombuffer(b"".join(sigs), dtype=np.uint8))
    assert ok.tolist() == [int(e) for e in expect]
    # the pyspec-literal form on explicit pubkeys gives the same verdicts
    flat, poff = [], [0]
    neg0 = g1_compress(E1.neg(g1_decompress(pks[0])))
    table = pks + [neg0, INF1, bytes(48)]
    for a in range(len(rows)):
        sel = [members[off[a] + j] for j, b in enumerate(rows[a]) if b]
        flat += [table[v] for v in sel]
        poff.append(len(flat))
    ok2 = eng.fast_aggregate_verify_pks(np.frombuffer(b"".join(flat), dtype=np.uint8), poff,
                                        np.frombuffer(b"".join(msgs), dtype=np.uint8), np.frombuffer(b"".join(sigs), dtype=np.uint8))
    assert ok2.tolist() == [int(e) for e in expect]
    # random-linear-combination batch mode: the same verdict vector (this batch is ONE group with rejected members: its equation
    # fails and every member is decided by the per-aggregate fallback), for both entry points, for two different seeds
    for seed in (b"seed-1", b"seed-2"):
        eng.set_verify_mode(True, hashlib.sha256(seed).digest())
        try:
            ok3 = eng.fast_aggregate_verify(members, off, _bits(rows), np.frombuffer(b"".join(msgs), dtype=np.uint8),
                                            np.frombuffer(b"".join(sigs), dtype=np.uint8))
            ok4 = eng.fast_aggregate_verify_pks(np.frombuffer(b"".join(flat), dtype=np.uint8), poff,
                                                np.frombuffer(b"".join(msgs), dtype=np.uint8), np.frombuffer(b"".join(sigs), dtype=np.uint8))
        finally:
            eng.set_verify_mode(False)
        assert ok3.tolist() == [int(e) for e in expect] and ok4.tolist() == [int(e) for e in expect]
    # groups that pass as a whole: the three honest aggregates repeated 24 times = 72 aggregates = 3 groups (32 + 32 + 8), no
    # fallback; then one corrupted signature in the middle group: exactly that aggregate flips
    hm, ho, hr, hmsg, hsig = [], [0], [], [], []
    for rep in range(24):
        for a in range(3):
            hm += members[off[a]:off[a + 1]]
            ho.append(len(hm))
            hr.append(rows[a])
            hmsg.append(msgs[a])
            hsig.append(sigs[a])
    eng.set_verify_mode(True, hashlib.sha256(b"seed-3").digest())
    try:
        okh = eng.fast_aggregate_verify(hm, ho, _bits(hr), np.frombuffer(b"".join(hmsg), dtype=np.uint8), np.frombuffer(b"".join(hsig), dtype=np.uint8))
        assert okh.tolist() == [1] * 72
        hsig[40] = hsig[41]                              # a valid signature, of another message / committee
        okh = eng.fast_aggregate_verify(hm, ho, _bits(hr), np.frombuffer(b"".join(hmsg), dtype=np.uint8), np.frombuffer(b"".join(hsig), dtype=np.uint8))
        assert okh.tolist() == [1] * 40 + [0] + [1] * 31
        # swapping two signatures inside a group keeps sum S_i: only the random scalars catch it
        hsig[40] = hsig[37]
        hsig[3], hsig[4] = hsig[4], hsig[3]
        okh = eng.fast_aggregate_verify(hm, ho, _bits(hr), np.frombuffer(b"".join(hmsg), dtype=np.uint8), np.frombuffer(b"".join(hsig), dtype=np.uint8))
        assert okh.tolist() == [1] * 3 + [0, 0] + [1] * 67
    finally:
        eng.set_verify_mode(False)


def test_aggregate_then_verify_linearity_512(eng):
    """Size-independent property at the BASELINE committee size: individual signatures of 512 validators,
    bls.Aggregate on the GPU, FastAggregateVerify on the GPU; flipping one participation bit must flip the verdict."""
    n = 512
    sks = np.zeros((n, 8), dtype=np.uint32)
    ks = [scenarios.secret_key(i) for i in range(n)]
    for i, k in enumerate(ks):
        for j in range(8):
            sks[i, j] = (k >> (32 * j)) & 0xFFFFFFFF
    pk = eng.sk_to_pk(sks)
    eng.registry_load(pk, np.full(n, 32 * 10**9, dtype=np.uint64))
    msg = np.frombuffer(hashlib.sha256(b"linearity").digest(), dtype=np.uint8).reshape(1, 32)
    sig = eng.sign(sks, np.zeros(n, dtype=np.uint32), msg)
    agg, st = eng.aggregate(sig, [0, n])
    assert st[0] == 0
    assert bytes(agg[0]) == B.Sign(sum(ks) % R, bytes(msg[0]))          # linearity: sum of signatures = signature of sum
    bits = np.full((2, 64), 0xFF, dtype=np.uint8)
    bits[1, 7] ^= 0x10
    members = np.concatenate([np.arange(n), np.arange(n)]).astype(np.uint32)
    ok = eng.fast_aggregate_verify(members, [0, n, 2 * n], bits, np.concatenate([msg, msg]), np.concatenate([agg, agg]))
    assert ok.tolist() == [1, 0]
