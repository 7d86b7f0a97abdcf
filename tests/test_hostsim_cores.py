"""The per-thread kernel bodies (csrc/cores.cuh), looped on the host by tests/hostsim, against the oracle:
registry load (decompress + KeyValidate), K2 gather+aggregate, K3 bls.Aggregate, K4-K6 verification."""
import ctypes
import hashlib

import numpy as np

import hs
import scenarios
from oracle import bls_sig as B
from oracle.bls12_381 import E1, E2, G1, R, g1_compress, g1_decompress, g2_compress
from oracle.hash_to_curve import map_to_curve_g2

lib = hs.load()
INF1 = bytes([0xC0]) + bytes(47)
INF2 = bytes([0xC0]) + bytes(95)


def _registry(pks):
    n = len(pks)
    rec = np.zeros(n * 24, dtype=np.uint32)
    valid = np.zeros(n, dtype=np.uint8)
    lib.hs_core_registry_load(hs.buf(b"".join(pks)), ctypes.c_uint64(n), hs.ptr(rec), hs.ptr(valid))
    return rec, valid


def test_registry_aggregate_verify_cores():
    pks = scenarios.pubkeys(8)
    neg0 = g1_compress(E1.neg(g1_decompress(pks[0])))
    table = pks + [neg0, INF1, bytes(48)]
    rec, valid = _registry(table)
    assert valid.tolist() == [1] * 9 + [0, 0]
    msg = hashlib.sha256(b"core").digest()
    aggs = [([0, 1, 2, 3], [1, 1, 1, 1]), ([4, 5, 6, 7], [1, 0, 1, 1]), ([0, 8], [1, 1]), ([1, 10], [1, 1]), ([2, 3], [0, 0])]
    members = np.array(sum((m for m, _ in aggs), []), dtype=np.uint32)
    off = np.cumsum([0] + [len(m) for m, _ in aggs]).astype(np.uint32)
    bits = np.zeros((len(aggs), 1), dtype=np.uint8)
    for a, (_, row) in enumerate(aggs):
        for j, b in enumerate(row):
            bits[a, 0] |= b << j
    n = len(aggs)
    jac = np.zeros(n * 36, dtype=np.uint32)
    status = np.zeros(n, dtype=np.uint8)
    out48 = ctypes.create_string_buffer(48 * n)
    lib.hs_core_g1_aggregate(hs.ptr(rec), hs.ptr(valid), hs.ptr(members), hs.ptr(off), hs.ptr(bits), 1, n, hs.ptr(jac), hs.ptr(status), out48)
    assert status.tolist() == [0, 0, 4, 1, 2 | 4]
    sel0 = [0, 1, 2, 3]
    sel1 = [4, 6, 7]
    acc = E1.INF
    for v in sel0:
        acc = E1.add(acc, g1_decompress(pks[v]))
    assert out48.raw[:48] == g1_compress(acc)
    assert out48.raw[96:144] == INF1
    sigs = [scenarios.sign_aggregate(sel0, msg), scenarios.sign_aggregate(sel1, msg), scenarios.sign_aggregate(sel0, msg),
            scenarios.sign_aggregate(sel0, msg), scenarios.sign_aggregate(sel0, msg)]
    ok = np.zeros(n, dtype=np.uint8)
    lib.hs_core_verify(hs.ptr(jac), hs.ptr(status), hs.buf(msg * n), hs.buf(b"".join(sigs)), n, hs.ptr(ok))
    assert ok.tolist() == [1, 1, 0, 0, 0]
    assert B.FastAggregateVerify([pks[v] for v in sel1], msg, sigs[1])
    # wrong message / infinity signature / non-subgroup signature on a valid aggregate
    bad = [sigs[0], INF2, g2_compress(map_to_curve_g2((9, 9))), bytes(96)]
    msgs = hashlib.sha256(b"other").digest() + msg * 3
    jac4 = np.tile(jac[:36], 4)
    ok = np.zeros(4, dtype=np.uint8)
    lib.hs_core_verify(hs.ptr(jac4), hs.ptr(np.zeros(4, dtype=np.uint8)), hs.buf(msgs), hs.buf(b"".join(bad)), 4, hs.ptr(ok))
    assert ok.tolist() == [0, 0, 0, 0]


def test_g2_aggregate_core():
    m = hashlib.sha256(b"agg-core").digest()
    sigs = scenarios.individual_signatures(list(range(6)), m)
    allsig = sigs + [INF2, bytes(96)]
    segs = [[0, 1, 2, 3, 4, 5], [2], [], [0, 6], [1, 7]]
    flat = [allsig[i] for s in segs for i in s]
    off = np.cumsum([0] + [len(s) for s in segs]).astype(np.uint32)
    out = ctypes.create_string_buffer(96 * len(segs))
    st = np.zeros(len(segs), dtype=np.int32)
    lib.hs_core_g2_aggregate(hs.buf(b"".join(flat)), hs.ptr(off), len(segs), out, hs.ptr(st))
    assert st.tolist() == [0, 0, 2, 0, 1]
    assert out.raw[:96] == B.Aggregate(sigs)
    assert out.raw[96:192] == sigs[2]
    assert out.raw[288:384] == sigs[0]


def test_signing_root_core_matches_oracle():
    from oracle import spec as OS
    from oracle.ssz import compute_domain
    rng = np.random.default_rng(2)
    for _ in range(5):
        d = OS.AttestationData(int(rng.integers(0, 2**62)), int(rng.integers(0, 64)), bytes(rng.integers(0, 256, 32, dtype=np.uint8)),
                               OS.Checkpoint(int(rng.integers(0, 2**40)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))),
                               OS.Checkpoint(int(rng.integers(0, 2**40)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))))
        dom = compute_domain(OS.DOMAIN_BEACON_ATTESTER, b"\x00\x00\x00\x01", bytes(rng.integers(0, 256, 32, dtype=np.uint8)))
        ser = (d.slot.to_bytes(8, "little") + d.index.to_bytes(8, "little") + d.beacon_block_root + d.source.epoch.to_bytes(8, "little") + d.source.root
               + d.target.epoch.to_bytes(8, "little") + d.target.root)
        assert len(ser) == 128
        out = ctypes.create_string_buffer(32)
        lib.hs_signing_root(hs.buf(ser), hs.buf(dom), out)
        assert out.raw == OS.Spec.compute_signing_root(d, dom)


def test_rlc_group_equation_on_the_host():
    """The random-linear-combination batch equation (csrc/cores.cuh, kernels in csrc/rlc.cuh) with the device code compiled for the
    host: scalars == SHA-256(seed || LE32(i) || msg)[0:8] | 1; a group of honest (pk, msg, sig) triples passes; replacing one
    signature by a valid signature of another message, or one key by another valid key, fails the group; an undecodable
    signature is rejected outright (not part of the equation) and the rest still passes."""
    import ctypes
    import hashlib
    from oracle import bls_sig as B
    lib = hs.load()
    seed = hashlib.sha256(b"rlc-seed").digest()
    n = 3
    sks = [scenarios.secret_key(i) for i in range(n)]
    msgs = [hashlib.sha256(bytes([i])).digest() for i in range(n)]
    pks = [B.SkToPk(k) for k in sks]
    sigs = [B.Sign(k, m) for k, m in zip(sks, msgs)]

    def run(pks_, sigs_):
        r = (ctypes.c_uint64 * n)()
        inb = (ctypes.c_uint8 * n)()
        gp = ctypes.c_uint8(9)
        rc = lib.hs_rlc_group(hs.buf(seed), hs.buf(b"".join(pks_)), hs.buf(b"".join(msgs)), hs.buf(b"".join(sigs_)), n, r, inb, ctypes.byref(gp))
        assert rc == 0
        return list(r), list(inb), int(gp.value)

    r, inb, gp = run(pks, sigs)
    for i in range(n):
        want = int.from_bytes(hashlib.sha256(seed + i.to_bytes(4, "little") + msgs[i]).digest()[:8], "little") | 1
        assert r[i] == want
    assert inb == [1, 1, 1] and gp == 1
    assert run(pks, [sigs[0], B.Sign(sks[1], msgs[2]), sigs[2]])[2] == 0          # wrong message under one signature
    assert run([pks[0], pks[2], pks[2]], sigs)[2] == 0                            # wrong key
    r2, inb2, gp2 = run(pks, [sigs[0], bytes(96), sigs[2]])                       # undecodable signature: rejected outright
    assert inb2 == [1, 0, 1] and gp2 == 1
    # a swap of two signatures between members keeps sum S_i but not sum r_i S_i: the random scalars are what catches it
    assert run(pks, [sigs[1], sigs[0], sigs[2]])[2] == 0
