"""Differential tests: the reference's own fenced python blocks (executed from
/root/reference/pos-evolution.md) vs the oracle's restatement vs the numpy array form.
Skipped where /root/reference is absent (GPU box) -- there tests/test_golden.py replays the
vectors these runs produced (tests/golden/gen_golden.py)."""
import copy
import hashlib

import numpy as np
import pytest

import ref_blocks
import scenarios
from oracle import fast
from oracle import spec as S

pytestmark = pytest.mark.skipif(not ref_blocks.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def env():
    import literal
    spec, state = scenarios.minimal_state(64, slot=9)
    return spec, state, literal.namespace(spec)


def test_shuffle_literal_vs_oracle_vs_numpy(env):
    spec, _, ns = env
    for n, tag in ((1, b"a"), (2, b"b"), (37, b"c"), (256, b"d"), (257, b"e"), (1000, b"f")):
        seed = hashlib.sha256(tag).digest()
        lit = [ns["compute_shuffled_index"](ns["uint64"](i), ns["uint64"](n), seed) for i in range(n)]
        assert lit == [spec.compute_shuffled_index(i, n, seed) for i in range(n)]
        assert lit == fast.shuffle_permutation(n, seed, spec.p.SHUFFLE_ROUND_COUNT).tolist()
        assert sorted(lit) == list(range(n))


def test_committees_literal_vs_numpy(env):
    spec, state, ns = env
    epoch = 1
    cps = ns["get_committee_count_per_slot"](state, epoch)
    assert cps == 2 == spec.get_committee_count_per_slot(state, epoch)
    seed = ns["get_seed"](state, epoch, S.DOMAIN_BEACON_ATTESTER)
    assert seed == spec.get_seed(state, epoch, S.DOMAIN_BEACON_ATTESTER)
    active = np.array(spec.get_active_validator_indices(state, epoch), dtype=np.uint32)
    members, off = fast.committees_for_epoch(active, seed, spec.p.SHUFFLE_ROUND_COUNT, cps, spec.p.SLOTS_PER_EPOCH)
    seen = []
    for s in range(spec.p.SLOTS_PER_EPOCH):
        for c in range(cps):
            lit = ns["get_beacon_committee"](state, epoch * spec.p.SLOTS_PER_EPOCH + s, c)
            k = s * cps + c
            assert lit == members[off[k]:off[k + 1]].tolist() == spec.get_beacon_committee(state, 8 + s, c)
            assert len(lit) == 4
            seen += lit
    assert sorted(seen) == list(range(64))          # committees partition the validator set (ref :455)


def _run(fn, state, att):
    st = copy.deepcopy(state)
    try:
        fn(st, att)
        return "ok", st
    except AssertionError:
        return "assert", None


def test_process_attestation_literal_vs_oracle(env):
    spec, state, ns = env
    cases = [
        scenarios.make_attestation(spec, state, 8, 0),
        scenarios.make_attestation(spec, state, 8, 1, bits=[True, False, True, False]),
        scenarios.make_attestation(spec, state, 5, 1),                                    # previous epoch
        scenarios.make_attestation(spec, state, 8, 0, corrupt="flip_bit"),
        scenarios.make_attestation(spec, state, 8, 0, corrupt="wrong_message"),
        scenarios.make_attestation(spec, state, 8, 1, corrupt="wrong_signer_set"),
        scenarios.make_attestation(spec, state, 8, 0, bits=[False] * 4),                  # empty -> invalid
        scenarios.make_attestation(spec, state, 8, 0, bits=[True] * 3),                   # wrong bit length
    ]
    bad_idx = scenarios.make_attestation(spec, state, 8, 0)
    bad_idx.data = S.AttestationData(8, 2, bad_idx.data.beacon_block_root, bad_idx.data.source, bad_idx.data.target)
    cases.append(bad_idx)                                                                 # index >= committee count
    expect = ["ok", "ok", "ok", "assert", "assert", "assert", "assert", "assert", "assert"]
    for att, exp in zip(cases, expect):
        r_lit, st_lit = _run(ns["process_attestation"], state, att)
        r_or, st_or = _run(spec.process_attestation, state, att)
        assert r_lit == r_or == exp
        if exp == "ok":
            assert st_lit.current_epoch_participation == st_or.current_epoch_participation
            assert st_lit.previous_epoch_participation == st_or.previous_epoch_participation
            assert st_lit.balances == st_or.balances
            assert st_lit.balances != state.balances
    # double inclusion: flags already set -> no second reward
    st = copy.deepcopy(state)
    ns["process_attestation"](st, cases[0])
    b1 = list(st.balances)
    ns["process_attestation"](st, cases[0])
    assert st.balances == b1


def _small_store(spec, state, n_blocks=200, seed=3):
    parent, slot, roots, leaf_viable = scenarios.fork_tree(n_blocks, seed)
    rb = [bytes(r) for r in roots]
    just = S.Checkpoint(1, rb[0])
    fin = S.Checkpoint(1, rb[0])
    store = S.Store(time=0, genesis_time=0, justified_checkpoint=just, finalized_checkpoint=fin,
                    best_justified_checkpoint=just, proposer_boost_root=rb[n_blocks - 1], equivocating_indices={3, 17})
    has_child = set(int(p) for p in parent[1:])
    for b in range(n_blocks):
        store.blocks[rb[b]] = S.BeaconBlock(int(slot[b]), rb[parent[b]] if b else bytes(32))
        bs = copy.copy(state)
        if b not in has_child and not leaf_viable[b]:
            bs.current_justified_checkpoint = S.Checkpoint(0, b"\x01" * 32)
        else:
            bs.current_justified_checkpoint = just
        bs.finalized_checkpoint = fin
        store.block_states[rb[b]] = bs
    store.checkpoint_states[just] = state
    rng = np.random.default_rng(seed)
    n = len(state.validators)
    for v in range(n):
        if rng.random() < 0.9:
            store.latest_messages[v] = S.LatestMessage(1, rb[int(n_blocks - 1 - min(n_blocks - 1, rng.geometric(0.05)))])
    return store, parent, slot, roots, leaf_viable, rb


def test_get_head_literal_vs_oracle_vs_numpy(env):
    spec, state, ns = env
    state = copy.deepcopy(state)
    state.validators[5].exit_epoch = 0          # inactive validator is not counted
    store, parent, slot, roots, leaf_viable, rb = _small_store(spec, state)
    head_lit = ns["get_head"](store)
    assert head_lit == spec.get_head(store)
    n = len(state.validators)
    idx_of = {r: i for i, r in enumerate(rb)}
    msg_block = np.zeros(n, dtype=np.uint32)
    has_msg = np.zeros(n, dtype=np.uint8)
    for v, lm in store.latest_messages.items():
        msg_block[v], has_msg[v] = idx_of[lm.root], 1
    eff = np.array([v.effective_balance for v in state.validators], dtype=np.uint64)
    active = np.array([spec.is_active_validator(v, spec.get_current_epoch(state)) for v in state.validators], dtype=np.uint8)
    equiv = np.zeros(n, dtype=np.uint8)
    equiv[list(store.equivocating_indices)] = 1
    boost = fast.proposer_boost_score(eff, active, spec.p.SLOTS_PER_EPOCH, spec.p.PROPOSER_SCORE_BOOST)
    w = fast.ghost_weights(parent, msg_block, has_msg, eff, active, equiv, len(rb) - 1, boost)
    for b in range(0, len(rb), 7):
        assert int(w[b]) == spec.get_latest_attesting_balance(store, rb[b])
    keep = fast.ghost_viable(parent, leaf_viable)
    assert set(rb[b] for b in range(len(rb)) if keep[b]) == set(spec.get_filtered_block_tree(store).keys())
    assert rb[fast.ghost_head(parent, roots, keep, w, 0)] == head_lit


def test_update_latest_messages_literal_vs_numpy(env):
    spec, state, ns = env
    store, parent, slot, roots, leaf_viable, rb = _small_store(spec, state, n_blocks=50)
    n = len(state.validators)
    idx_of = {r: i for i, r in enumerate(rb)}
    msg_epoch = np.zeros(n, dtype=np.uint64)
    msg_block = np.zeros(n, dtype=np.uint32)
    has_msg = np.zeros(n, dtype=np.uint8)
    for v, lm in store.latest_messages.items():
        msg_epoch[v], msg_block[v], has_msg[v] = lm.epoch, idx_of[lm.root], 1
    equiv = np.zeros(n, dtype=np.uint8)
    equiv[list(store.equivocating_indices)] = 1
    for epoch, blk, idxs in ((1, 10, [1, 2, 3, 40]), (2, 20, [2, 3, 17, 41, 63]), (1, 30, [2, 50])):
        att = S.Attestation([], S.AttestationData(0, 0, rb[blk], S.Checkpoint(), S.Checkpoint(epoch, rb[0])), b"")
        ns["update_latest_messages"](store, idxs, att)
        fast.lmd_update(msg_epoch, msg_block, has_msg, equiv, idxs, epoch, blk)
    for v in range(n):
        if has_msg[v]:
            assert store.latest_messages[v] == S.LatestMessage(int(msg_epoch[v]), rb[msg_block[v]])
        else:
            assert v not in store.latest_messages


def test_ffg_literal_vs_oracle(env):
    """process_justification_and_finalization / weigh_justification_and_finalization: the reference's own text (ref :793-852)
    against the oracle's restatement on 80 generated end-of-epoch states; every branch must have fired."""
    import copy
    spec, state0, ns = env[0], env[1], env[2]
    seen = set()
    for seed in range(80):
        st_ref = scenarios.ffg_case(copy.deepcopy(state0), seed)
        st_or = copy.deepcopy(st_ref)
        before = (st_ref.current_justified_checkpoint, st_ref.finalized_checkpoint)
        ns["process_justification_and_finalization"](st_ref)
        spec.process_justification_and_finalization(st_or)
        got, want = scenarios.ffg_outcome(st_or), scenarios.ffg_outcome(st_ref)
        assert got == want, seed
        seen.add(("justified_changed", st_ref.current_justified_checkpoint != before[0]))
        seen.add(("finalized_changed", st_ref.finalized_checkpoint != before[1]))
        seen.add(("bits", tuple(want["justification_bits"][:2])))
        if st_ref.finalized_checkpoint != before[1]:
            seen.add(("finalized_distance", spec.get_current_epoch(st_ref) - st_ref.finalized_checkpoint.epoch))
    assert {("justified_changed", True), ("justified_changed", False), ("finalized_changed", True), ("finalized_changed", False)} <= seen
    assert {("bits", (0, 0)), ("bits", (0, 1)), ("bits", (1, 0)), ("bits", (1, 1))} <= seen
    assert {("finalized_distance", 1), ("finalized_distance", 2), ("finalized_distance", 3)} <= seen
