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
