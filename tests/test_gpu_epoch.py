"""GPU parity of the fused epoch pipeline (b2_epoch_dev) and of its software-pipelined form (depth 2, 3, 4) against the oracle:
per-committee aggregate signatures (bytes), verdicts, the latest-message table and the head after every epoch."""
import hashlib

import numpy as np
import pytest
import torch

import scenarios
from oracle import bls_sig as B
from oracle import fast

pytestmark = pytest.mark.gpu

N_VAL, CSIZE, N_AGG, N_BLK = 64, 4, 16, 40


def _world():
    pks = scenarios.pubkeys(N_VAL)
    rng = np.random.default_rng(21)
    members = rng.permutation(N_VAL).astype(np.uint32)
    off = (np.arange(N_AGG + 1) * CSIZE).astype(np.uint32)
    parent, slot, roots, leaf_viable = scenarios.fork_tree(N_BLK, 5)
    eff = (rng.integers(16, 33, size=N_VAL).astype(np.uint64)) * np.uint64(10**9)
    return pks, members, off, (parent, slot, roots, leaf_viable), eff


def _epoch_inputs(k, members, off, rng):
    """Individual signatures for epoch k; aggregates 3 and 7 get one foreign signature (aggregate verifies False),
    aggregate 5 has a partially set bitfield whose signers match (True), aggregate 9 a bitfield that does not match (False)."""
    msgs = [hashlib.sha256(b"epoch%d/agg%d" % (k, a)).digest() for a in range(N_AGG)]
    bits = np.full((N_AGG, 1), 0x0F, dtype=np.uint8)
    sigs = []
    for a in range(N_AGG):
        com = [int(v) for v in members[off[a]:off[a + 1]]]
        signers = list(com)
        if a == 5:
            bits[a, 0] = 0x0B
        if a == 9:
            bits[a, 0] = 0x07                       # bit says 3 signers, but all 4 signed
        s = scenarios.individual_signatures(signers, msgs[a])
        if a in (3, 7):
            s[1] = scenarios.individual_signatures([(com[1] + 1) % N_VAL], msgs[a])[0]
        if a == 5:
            s[2] = bytes([0xC0]) + bytes(95)        # the non-signer contributes the identity: the aggregate equals the 3-signer sum
        sigs += s
    target_epoch = np.full(N_AGG, 2 + k, dtype=np.int64)
    block_idx = ((np.arange(N_AGG) * 7 + 3 * k) % N_BLK).astype(np.int32)
    return msgs, bits, sigs, target_epoch, block_idx


_CACHE = {}


def _epoch_case(k, pks, members, off):
    """Inputs and oracle results of epoch k; the same for every mode, so the (pure-Python, slow) oracle runs once per epoch."""
    if k not in _CACHE:
        msgs, bits, sigs, te, bi = _epoch_inputs(k, members, off, None)
        aggs, oks = _expected(pks, members, off, msgs, bits, sigs)
        _CACHE[k] = (msgs, bits, sigs, te, bi, aggs, oks)
    return _CACHE[k]


def _expected(pks, members, off, msgs, bits, sigs):
    aggs, oks = [], []
    for a in range(N_AGG):
        seg = sigs[off[a]:off[a + 1]]
        agg = B.Aggregate(seg)
        sel = [int(members[off[a] + j]) for j in range(CSIZE) if (bits[a, 0] >> j) & 1]
        aggs.append(agg)
        oks.append(int(B.FastAggregateVerify([pks[v] for v in sel], msgs[a], agg)))
    return aggs, oks


@pytest.mark.parametrize("mode,depth", [("sync", 3), ("pipelined", 2), ("pipelined", 3), ("pipelined", 4), ("pipelined", 8), ("pipelined", 12), ("pipelined_host", 3),
                                        ("pipelined_host", 2), ("sync_host", 3), ("pipelined_team", 6), ("sync_rlc", 3), ("pipelined_rlc", 3),
                                        ("pipelined_host_rlc", 4)])
def test_epoch_pipeline_matches_oracle(mode, depth):
    """*_rlc: the same epochs with FastAggregateVerify in random-linear-combination batches (b2_set_verify_mode): the three rejected
    aggregates sit in the only group, so its equation fails and every member is decided by the per-aggregate fallback;
    pipelined_team: the always-team tail a sharded rank uses."""
    rlc = mode.endswith("_rlc")
    if rlc:
        mode = mode[:-4]
    tail_form = "team" if mode == "pipelined_team" else "thread"
    if mode == "pipelined_team":
        mode = "pipelined"
    from pos_evolution_b200.engine import Engine
    from pos_evolution_b200.epoch import EpochProcessor
    pks, members, off, tree, eff = _world()
    parent, slot, roots, leaf_viable = tree
    eng = Engine(0)
    eng.registry_load(np.frombuffer(b"".join(pks), dtype=np.uint8), eff)
    eng.tree_load(parent, slot, roots, leaf_viable)
    eng.latest_messages_reset()
    dev = torch.device("cuda", 0)
    if rlc:
        eng.set_verify_mode(True, hashlib.sha256(b"test-seed").digest())
    ep = EpochProcessor(eng, N_AGG, N_AGG * CSIZE, 1, N_BLK, device=dev, depth=depth, tail_form=tail_form)
    ep.set_committees(members, off)
    rng = np.random.default_rng(3)

    def collect(t):
        ok, hd = t.wait()
        tickets.append((ok.cpu().numpy().tolist(), hd, t.d_agg_sig.cpu().numpy().copy()))

    m_epoch, m_block, m_has = np.zeros(N_VAL, np.uint64), np.zeros(N_VAL, np.uint32), np.zeros(N_VAL, np.uint8)
    equiv, active = np.zeros(N_VAL, np.uint8), np.ones(N_VAL, np.uint8)
    expected, tickets, keep_alive = [], [], []
    n_epochs = 6
    for k in range(n_epochs):
        msgs, bits, sigs, te, bi, aggs, oks = _epoch_case(k, pks, members, off)
        assert sum(oks) == N_AGG - 3 and oks[5] == 1 and oks[9] == 0
        for a in range(N_AGG):                      # sequential oracle for the LMD table
            if oks[a]:
                sel = [int(members[off[a] + j]) for j in range(CSIZE) if (bits[a, 0] >> j) & 1]
                fast.lmd_update(m_epoch, m_block, m_has, equiv, sel, int(te[a]), int(bi[a]))
        w = fast.ghost_weights(parent, m_block, m_has, eff, active, equiv, -1, 0)
        head = fast.ghost_head(parent, roots, fast.ghost_viable(parent, leaf_viable), w, 0)
        expected.append((aggs, oks, head))
        d = [torch.as_tensor(np.frombuffer(b"".join(sigs), dtype=np.uint8).reshape(-1, 96).copy(), device=dev),
             torch.as_tensor(bits, device=dev), torch.as_tensor(np.frombuffer(b"".join(msgs), dtype=np.uint8).reshape(-1, 32).copy(), device=dev),
             torch.as_tensor(te, device=dev), torch.as_tensor(bi, device=dev)]
        keep_alive.append(d)
        if mode in ("pipelined_host", "sync_host"):
            h = [x.cpu().pin_memory() for x in d]
            keep_alive.append(h)
            if mode == "sync_host":
                ok, hd = ep.process_epoch_host(*h)
                tickets.append((ok.numpy().tolist(), hd, ep.d_agg_sig[0].cpu().numpy().copy()))
            else:
                t = ep.submit_host(*h, last=(k == n_epochs - 1))
                if t is not None:
                    collect(t)
        elif mode == "sync":
            ok, hd = ep.process_epoch_dev(*d)
            torch.cuda.synchronize()
            tickets.append((ok.cpu().numpy().tolist(), int(hd.item()), ep.d_agg_sig[0].cpu().numpy().copy()))
        else:
            t = ep.submit_dev(*d, last=(k % 3 == 2))          # exercises both pairing forms under the pipeline
            if t is not None:
                collect(t)
    if mode in ("pipelined", "pipelined_host"):
        for t in ep.drain():
            collect(t)
    assert len(tickets) == n_epochs
    for k in range(n_epochs):
        aggs, oks, head = expected[k]
        got_ok, got_head, got_agg = tickets[k]
        assert got_ok == oks, k
        assert got_head == head, k
        assert [bytes(r) for r in got_agg] == aggs, k
    e, b, h = eng.latest_messages_read()
    assert np.array_equal(h, m_has) and np.array_equal(e[h == 1], m_epoch[h == 1]) and np.array_equal(b[h == 1], m_block[h == 1])
    eng.close()
