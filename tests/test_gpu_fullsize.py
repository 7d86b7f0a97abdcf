"""Full BASELINE.json size (2^20 validators, 2 048 committees of 512, 10 000 blocks) through the C ABI, checked with
size-independent properties instead of the (far too slow) BLS oracle:
  * every honest aggregate verifies; tampering with t signatures makes exactly their t committees fail (and nothing else);
  * an identity contribution whose bit is cleared leaves the verdict true;
  * checksum of checksums: Aggregate(all 2^20 signatures as one segment) == Aggregate(the 2 048 committee aggregates);
  * update_latest_messages + get_weight + get_head equal the numpy oracle (oracle/fast.py) on the full arrays, through the
    synchronous and the pipelined (depth 3) forms of the epoch.
Inputs are made by the product's own kernels (bench.build_world), like the bench."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def full():
    import bench
    from pos_evolution_b200 import spec as PS
    from pos_evolution_b200.engine import Engine
    eng = Engine(0)
    W = bench.build_world(eng, 0, np, PS)
    yield eng, W, bench
    eng.close()


def test_full_epoch_properties(full):
    eng, W, bench = full
    from pos_evolution_b200.epoch import EpochProcessor
    dev = torch.device("cuda", 0)
    N_AGG, CS = bench.N_AGG, bench.COMMITTEE_SIZE
    ep = EpochProcessor(eng, N_AGG, bench.N_VAL, CS // 8, bench.N_BLOCKS, device=dev, depth=3)
    ep.set_committees(W["members"], W["off"])
    rng = np.random.default_rng(11)
    bits = np.full((N_AGG, CS // 8), 0xFF, dtype=np.uint8)
    sigs = W["sigs"].copy()
    # tamper: t signatures replaced by a valid signature of ANOTHER message (still a point of G2: only the pairing can tell)
    bad_committees = sorted(rng.choice(N_AGG, size=7, replace=False).tolist())
    for a in bad_committees:
        j = int(W["off"][a]) + int(rng.integers(CS))
        other = (a + 1) % N_AGG
        sigs[j] = W["sigs"][int(W["off"][other]) + int(rng.integers(CS))]
    # identity contribution with its bit cleared: verdict stays true
    ident = bytes([0xC0]) + bytes(95)
    id_committees = [a for a in range(3, N_AGG, 401) if a not in bad_committees][:4]
    for a in id_committees:
        k = int(rng.integers(CS))
        sigs[int(W["off"][a]) + k] = np.frombuffer(ident, dtype=np.uint8)
        bits[a, k >> 3] &= ~np.uint8(1 << (k & 7))
    target_epoch = np.full(N_AGG, 5, dtype=np.int64)
    blk = (bench.N_BLOCKS - 1 - (np.arange(N_AGG) % 64)).astype(np.int32)
    d = [torch.as_tensor(sigs, device=dev), torch.as_tensor(bits, device=dev), torch.as_tensor(W["msgs"], device=dev),
         torch.as_tensor(target_epoch, device=dev), torch.as_tensor(blk, device=dev)]
    ok, head = ep.process_epoch_dev(*d, 0, bench.N_BLOCKS - 1, W["boost"])
    torch.cuda.synchronize()
    ok = ok.cpu().numpy()
    expect = np.ones(N_AGG, dtype=np.uint8)
    expect[bad_committees] = 0
    assert np.array_equal(ok, expect)
    assert int(ep.d_agg_status[0].abs().sum().item()) == 0
    # fork choice after the epoch == numpy oracle (identity contributors did not attest: their bit is clear)
    members, off = W["members"], W["off"]
    msg_block, has_msg, equiv, eff, active = [x.copy() for x in W["votes"]]
    from oracle import fast
    m_epoch = np.ones(bench.N_VAL, dtype=np.uint64)
    for a in np.nonzero(expect)[0]:
        sel = [int(members[off[a] + j]) for j in range(CS) if (bits[a, j >> 3] >> (j & 7)) & 1]
        fast.lmd_update(m_epoch, msg_block, has_msg, equiv, sel, int(target_epoch[a]), int(blk[a]))
    parent, roots, leaf_viable = W["tree"]
    w = fast.ghost_weights(parent, msg_block, has_msg, eff, active, equiv, bench.N_BLOCKS - 1, W["boost"])
    want_head = fast.ghost_head(parent, roots, fast.ghost_viable(parent, leaf_viable), w, 0)
    assert int(head.item()) == want_head
    assert np.array_equal(eng.get_weights(bench.N_BLOCKS - 1, W["boost"]), w)
    e, b, h = eng.latest_messages_read()
    assert np.array_equal(h, has_msg) and np.array_equal(e[h == 1], m_epoch[h == 1]) and np.array_equal(b[h == 1], msg_block[h == 1])

    # checksum of checksums on the honest signatures
    agg_all, st_all = eng.aggregate(W["sigs"], np.array([0, bench.N_VAL], dtype=np.uint32))
    agg_c, st_c = eng.aggregate(W["sigs"], W["off"])
    assert int(st_all[0]) == 0 and not st_c.any()
    agg_of_aggs, st2 = eng.aggregate(agg_c, np.array([0, N_AGG], dtype=np.uint32))
    assert int(st2[0]) == 0 and bytes(agg_of_aggs[0]) == bytes(agg_all[0])

    # pipelined form, three honest epochs with later targets: every verdict true, heads == oracle of the running table
    eng.latest_messages_load(np.ones(bench.N_VAL, dtype=np.uint64), *W["votes"][:3])
    bits1 = torch.full((N_AGG, CS // 8), 0xFF, dtype=torch.uint8, device=dev)
    d_sigs = torch.as_tensor(W["sigs"], device=dev)
    tickets, want = [], []
    msg_block, has_msg, equiv, eff, active = [x.copy() for x in W["votes"]]
    m_epoch = np.ones(bench.N_VAL, dtype=np.uint64)
    keep = []
    for k in range(4):
        te = np.full(N_AGG, 7 + k, dtype=np.int64)
        bk = ((blk.astype(np.int64) - 97 * k) % bench.N_BLOCKS).astype(np.int32)
        for a in range(N_AGG):
            fast.lmd_update(m_epoch, msg_block, has_msg, equiv, members[off[a]:off[a + 1]], int(te[a]), int(bk[a]))
        w = fast.ghost_weights(parent, msg_block, has_msg, eff, active, equiv, bench.N_BLOCKS - 1, W["boost"])
        want.append(fast.ghost_head(parent, roots, fast.ghost_viable(parent, leaf_viable), w, 0))
        dk = [torch.as_tensor(te, device=dev), torch.as_tensor(bk, device=dev)]
        keep.append(dk)
        t = ep.submit_dev(d_sigs, bits1, d[2], dk[0], dk[1], 0, bench.N_BLOCKS - 1, W["boost"], last=(k == 3))   # last: team-form tail
        if t is not None:
            tickets.append(t.wait())
    tickets += [t.wait() for t in ep.drain()]
    assert len(tickets) == 4
    for k, (okk, hd) in enumerate(tickets):
        assert hd == want[k], k
    assert int(tickets[-1][0].sum().item()) == N_AGG
    e, b, h = eng.latest_messages_read()
    assert np.array_equal(h, has_msg) and np.array_equal(e[h == 1], m_epoch[h == 1]) and np.array_equal(b[h == 1], msg_block[h == 1])


# ------------------------------------------------------------------------------------------ bytes, not only properties
# The keys of the synthetic world are an arithmetic progression sk_i = sk0 + i*delta, so for any set S of
# validators the oracle knows the aggregate pubkey SkToPk(sum sk_i) and the aggregate signature Sign(sum sk_i, m) without
# adding 512 points in pure Python.  These tests compare the BYTES the GPU produces at full size with those oracle values.
def test_aggregate_signature_bytes_vs_oracle_linearity(full):
    """bls.Aggregate at full size: 64 of the 2 048 committee aggregates (every 32nd), byte for byte against the oracle."""
    eng, W, bench = full
    from oracle import bls_sig as B
    agg, st = eng.aggregate(W["sigs"], W["off"])
    assert not st.any()
    for a in range(0, bench.N_AGG, 32):
        want = B.Sign(bench.committee_secret_sum(W, a), bytes(W["msgs"][a]))
        assert bytes(agg[a]) == want, a


def test_config2_bls_aggregate_32768_signatures(full):
    """BASELINE.json config 2: 32 768 signatures -> 64 segments of 512; all 64 outputs == oracle.Sign(sum sk, m_c)."""
    eng, W, bench = full
    from oracle import bls_sig as B
    n2 = 64 * bench.COMMITTEE_SIZE
    agg, st = eng.aggregate(W["sigs"][:n2], W["off"][:65])
    assert not st.any()
    for a in range(64):
        assert bytes(agg[a]) == B.Sign(bench.committee_secret_sum(W, a), bytes(W["msgs"][a])), a
    # ragged segmentation of the same signatures: 3 uneven segments + an empty one (status 2), checksum of checksums
    seg = np.array([0, 1, 700, 700, n2], dtype=np.uint32)
    agg_r, st_r = eng.aggregate(W["sigs"][:n2], seg)
    assert st_r.tolist() == [0, 0, 2, 0]
    tot, st_t = eng.aggregate(agg_r[[0, 1, 3]], np.array([0, 3], dtype=np.uint32))
    tot2, st_t2 = eng.aggregate(agg, np.array([0, 64], dtype=np.uint32))
    assert int(st_t[0]) == 0 and int(st_t2[0]) == 0 and bytes(tot[0]) == bytes(tot2[0])


@pytest.mark.parametrize("frac,seed", [(0.99, 1), (0.5, 2)])
def test_k2_partial_participation_bytes_vs_oracle(full, frac, seed):
    """K2 (b2_g1_aggregate, the TMA-staged gather) on 512-member committees at 99 % / 50 % random participation: the compressed
    aggregate pubkey of 48 committees == oracle.SkToPk(sum of the selected secret keys)."""
    eng, W, bench = full
    from oracle import bls_sig as B
    rng = np.random.default_rng(seed)
    sel = rng.random((bench.N_AGG, bench.COMMITTEE_SIZE)) < frac
    sel[:, 0] = True
    bits = np.packbits(sel, axis=1, bitorder="little")
    out, status = eng.g1_aggregate(W["members"], W["off"], bits)
    assert not status.any()
    for a in list(range(0, bench.N_AGG, 64)) + [1, 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 2047]:
        assert bytes(out[a]) == B.SkToPk(bench.committee_secret_sum(W, a, bits[a])), a


@pytest.mark.parametrize("name,frac,seed", [("100pct", 1.0, 0), ("99pct", 0.99, 1), ("50pct", 0.5, 2)])
def test_config3_participation_and_corruption(full, name, frac, seed):
    """BASELINE.json config 3: 2 048 x 512 FastAggregateVerify at 100 / 99 / 50 % participation with 1 % of the
    aggregates corrupted: exactly the corrupted ones are rejected; the aggregate signatures of the selected subsets (made by
    the GPU's bls.Aggregate) equal the oracle's Sign(sum of selected sk, m) on a sample."""
    eng, W, bench = full
    from oracle import bls_sig as B
    bits, agg, expect = bench.participation_case(eng, W, np, frac, seed)
    ok = eng.fast_aggregate_verify(W["members"], W["off"], bits, W["msgs"], agg)
    assert np.array_equal(ok, expect)
    assert int((expect == 0).sum()) == 20
    # the same batch through the random-linear-combination mode: every corrupted aggregate is still individually identified
    eng.set_verify_mode(True, bytes(range(32)))
    try:
        ok_rlc = eng.fast_aggregate_verify(W["members"], W["off"], bits, W["msgs"], agg)
    finally:
        eng.set_verify_mode(False)
    assert np.array_equal(ok_rlc, expect)
    good = [a for a in range(0, bench.N_AGG, 256) if expect[a]]
    for a in good:
        assert bytes(agg[a]) == B.Sign(bench.committee_secret_sum(W, a, bits[a]), bytes(W["msgs"][a])), a
