"""GPU tests of the pyspec-signature layer (pos_evolution_b200.spec / .bls) against the oracle's literal
restatement on the minimal preset: same accept/reject decisions, same state mutation, same head root."""
import copy

import numpy as np
import pytest

import scenarios
from oracle import bls_sig as OB
from oracle import spec as OS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from pos_evolution_b200 import bls as PB
    from pos_evolution_b200 import spec as PS
    from pos_evolution_b200.engine import Engine
    eng = Engine(0)
    PB.use_engine(eng)
    pks = scenarios.pubkeys(64)
    ospec, ostate = scenarios.minimal_state(64, slot=9, pks=pks)
    pspec = PS.Spec(PS.MINIMAL, engine=Engine(0))
    return ospec, ostate, pspec, PS, PB, pks


def _to_product(PS, x):
    """oracle dataclasses -> product dataclasses (same field names)."""
    if isinstance(x, OS.Checkpoint):
        return PS.Checkpoint(x.epoch, x.root)
    if isinstance(x, OS.AttestationData):
        return PS.AttestationData(x.slot, x.index, x.beacon_block_root, _to_product(PS, x.source), _to_product(PS, x.target))
    if isinstance(x, OS.Attestation):
        return PS.Attestation(list(x.aggregation_bits), _to_product(PS, x.data), x.signature)
    raise TypeError(x)


def _state_to_product(PS, s):
    return PS.BeaconState(
        slot=s.slot, fork=PS.Fork(s.fork.previous_version, s.fork.current_version, s.fork.epoch),
        genesis_validators_root=s.genesis_validators_root,
        validators=[PS.Validator(v.pubkey, v.effective_balance, v.slashed, v.activation_epoch, v.exit_epoch) for v in s.validators],
        balances=list(s.balances), randao_mixes=list(s.randao_mixes), block_roots=list(s.block_roots),
        previous_epoch_participation=list(s.previous_epoch_participation), current_epoch_participation=list(s.current_epoch_participation),
        previous_justified_checkpoint=_to_product(PS, s.previous_justified_checkpoint),
        current_justified_checkpoint=_to_product(PS, s.current_justified_checkpoint),
        finalized_checkpoint=_to_product(PS, s.finalized_checkpoint))


def test_bls_facade_matches_oracle(env):
    ospec, ostate, pspec, PS, PB, pks = env
    m = b"\x42" * 32
    sks = [scenarios.secret_key(i) for i in range(4)]
    assert [PB.SkToPk(k) for k in sks] == pks[:4]
    sigs = [PB.Sign(k, m) for k in sks]
    assert sigs == [OB.Sign(k, m) for k in sks]
    agg = PB.Aggregate(sigs)
    assert agg == OB.Aggregate(sigs)
    assert PB.FastAggregateVerify(pks[:4], m, agg) is True
    assert PB.FastAggregateVerify(pks[:3], m, agg) is False
    assert PB.FastAggregateVerify([], m, agg) is False
    assert PB.FastAggregateVerify(pks[:4], m, b"\x00" * 96) is False
    assert PB.FastAggregateVerify(pks[:4], m, b"short") is False
    assert PB.Verify(pks[0], m, sigs[0]) is True and PB.Verify(pks[1], m, sigs[0]) is False
    assert PB.KeyValidate(pks[0]) and not PB.KeyValidate(bytes([0xC0]) + bytes(47)) and not PB.KeyValidate(bytes(48))
    with pytest.raises(ValueError):
        PB.Aggregate([])
    with pytest.raises(ValueError):
        PB.Aggregate([sigs[0], bytes(96)])


def test_process_attestation_matches_oracle(env):
    ospec, ostate, pspec, PS, PB, pks = env
    cases = [
        (scenarios.make_attestation(ospec, ostate, 8, 0), True),
        (scenarios.make_attestation(ospec, ostate, 8, 1, bits=[True, False, True, False]), True),
        (scenarios.make_attestation(ospec, ostate, 5, 1), True),
        (scenarios.make_attestation(ospec, ostate, 8, 0, corrupt="flip_bit"), False),
        (scenarios.make_attestation(ospec, ostate, 8, 0, corrupt="wrong_message"), False),
        (scenarios.make_attestation(ospec, ostate, 8, 1, corrupt="wrong_signer_set"), False),
        (scenarios.make_attestation(ospec, ostate, 8, 0, bits=[False] * 4), False),
        (scenarios.make_attestation(ospec, ostate, 8, 0, bits=[True] * 3), False),
    ]
    for att, valid in cases:
        so = copy.deepcopy(ostate)
        sp = _state_to_product(PS, ostate)
        if valid:
            ospec.process_attestation(so, att)
            pspec.process_attestation(sp, _to_product(PS, att))
            assert sp.balances == so.balances
            assert sp.current_epoch_participation == so.current_epoch_participation
            assert sp.previous_epoch_participation == so.previous_epoch_participation
        else:
            with pytest.raises(AssertionError):
                ospec.process_attestation(so, att)
            with pytest.raises(AssertionError):
                pspec.process_attestation(sp, _to_product(PS, att))
            assert sp.balances == ostate.balances and sp.current_epoch_participation == ostate.current_epoch_participation
    # the batched form: all 16 aggregates of an epoch in one GPU call
    so = copy.deepcopy(ostate)
    so.slot = 16
    so.block_roots = list(ostate.block_roots)
    sp = _state_to_product(PS, so)
    atts = [scenarios.make_attestation(ospec, so, s, i) for s in range(8, 16) for i in range(2)]
    for a in atts:
        ospec.process_attestation(so, a)
    pspec.process_attestations(sp, [_to_product(PS, a) for a in atts])
    assert sp.balances == so.balances and sp.current_epoch_participation == so.current_epoch_participation
    assert sp.previous_epoch_participation == so.previous_epoch_participation


def test_get_head_and_weight_match_oracle(env):
    ospec, ostate, pspec, PS, PB, pks = env
    n_blocks = 120
    parent, slot, roots, leaf_viable = scenarios.fork_tree(n_blocks, 6)
    rb = [bytes(r) for r in roots]
    rng = np.random.default_rng(6)
    ostate2 = copy.deepcopy(ostate)
    ostate2.validators[7].exit_epoch = 0
    pstate = _state_to_product(PS, ostate2)
    has_child = set(int(p) for p in parent[1:])
    for boost in (OS.ZERO32, rb[n_blocks - 1]):
        oj, pj = OS.Checkpoint(1, rb[0]), PS.Checkpoint(1, rb[0])
        ostore = OS.Store(0, 0, oj, oj, oj, boost, {3, 17})
        pstore = PS.Store(0, 0, pj, pj, pj, boost, {3, 17})
        for b in range(n_blocks):
            pr = rb[parent[b]] if b else bytes(32)
            ostore.blocks[rb[b]] = OS.BeaconBlock(int(slot[b]), pr)
            pstore.blocks[rb[b]] = PS.BeaconBlock(int(slot[b]), pr)
            obs, pbs = copy.copy(ostate2), copy.copy(pstate)
            bad = b not in has_child and not leaf_viable[b]
            obs.current_justified_checkpoint = OS.Checkpoint(0, b"\x01" * 32) if bad else oj
            pbs.current_justified_checkpoint = PS.Checkpoint(0, b"\x01" * 32) if bad else pj
            obs.finalized_checkpoint, pbs.finalized_checkpoint = oj, pj
            ostore.block_states[rb[b]], pstore.block_states[rb[b]] = obs, pbs
        ostore.checkpoint_states[oj], pstore.checkpoint_states[pj] = ostate2, pstate
        for v in range(64):
            if rng.random() < 0.9:
                r = rb[int(n_blocks - 1 - min(n_blocks - 1, rng.geometric(0.05)))]
                ostore.latest_messages[v] = OS.LatestMessage(1, r)
                pstore.latest_messages[v] = PS.LatestMessage(1, r)
        assert pspec.get_head(pstore) == ospec.get_head(ostore)
        for b in (0, 1, n_blocks // 2, n_blocks - 1):
            assert pspec.get_weight(pstore, rb[b]) == ospec.get_latest_attesting_balance(ostore, rb[b])
    # on_attestation: validate_on_attestation (store-consistent data only), verify on the GPU, then the host LMD table moves exactly
    # like the oracle's AND the device mirror of the store moves with it (the next get_head needs no re-upload)
    k = max(b for b in range(n_blocks) if slot[b] <= 8)
    att = scenarios.make_attestation(ospec, ostate2, 8, 0, head_root=rb[k], target_root=rb[k])
    ostore.checkpoint_states[att.data.target] = ostate2
    pstore.checkpoint_states[_to_product(PS, att.data.target)] = pstate
    pstore.time = ostore.time = 9 * 12                                   # current slot 9: the slot-8 attestation may be processed
    head_before = pspec.get_head(pstore)
    uploads = dict(pspec.stats)
    ospec.on_attestation(ostore, att)
    pspec.on_attestation(pstore, _to_product(PS, att))
    assert {v: (m.epoch, m.root) for v, m in pstore.latest_messages.items()} == {v: (m.epoch, m.root) for v, m in ostore.latest_messages.items()}
    assert pspec.get_head(pstore) == ospec.get_head(ostore)
    for b in (0, 1, k, n_blocks - 1):
        assert pspec.get_weight(pstore, rb[b]) == ospec.get_latest_attesting_balance(ostore, rb[b])
    assert pspec.stats["store_uploads"] == uploads["store_uploads"] and pspec.stats["registry_uploads"] == uploads["registry_uploads"]
    assert pspec.stats["lmd_device_updates"] == uploads["lmd_device_updates"] + 1
    del head_before
    bad = scenarios.make_attestation(ospec, ostate2, 8, 0, corrupt="flip_bit", head_root=rb[k], target_root=rb[k])
    with pytest.raises(AssertionError):
        pspec.on_attestation(pstore, _to_product(PS, bad))
    # validate_on_attestation: unknown head block, a head block from the future of the attestation, an attestation from the current slot
    unknown = scenarios.make_attestation(ospec, ostate2, 8, 0, head_root=b"\x55" * 32, target_root=rb[k])
    later = scenarios.make_attestation(ospec, ostate2, 8, 0, head_root=rb[n_blocks - 1], target_root=rb[k])
    for a in (unknown, later):
        pstore.checkpoint_states[_to_product(PS, a.data.target)] = pstate
        with pytest.raises(AssertionError):
            pspec.on_attestation(pstore, _to_product(PS, a))
    pstore.time = 8 * 12
    with pytest.raises(AssertionError):
        pspec.on_attestation(pstore, _to_product(PS, att))
    # a block arrives (on_block is the caller's): the mirror notices the new shape and re-uploads; heads keep matching the oracle
    new_root = b"\x77" * 32
    for st_, S_, bs in ((ostore, OS, ostore.block_states[rb[k]]), (pstore, PS, pstore.block_states[rb[k]])):
        st_.blocks[new_root] = S_.BeaconBlock(int(slot[k]) + 1, rb[k])
        st_.block_states[new_root] = bs
    pstore.time = ostore.time = 9 * 12
    assert pspec.get_head(pstore) == ospec.get_head(ostore)
    assert pspec.stats["store_uploads"] == uploads["store_uploads"] + 1


def test_on_attester_slashing_marks_equivocators(env):
    """on_attester_slashing (:1447-1461): two conflicting, correctly signed indexed attestations -> the common signers become
    equivocating and stop counting in get_weight; an unsigned/invalid slashing is rejected and leaves the store untouched."""
    ospec, ostate, pspec, PS, PB, pks = env
    pstate = _state_to_product(PS, ostate)
    com = ospec.get_beacon_committee(ostate, 8, 0)
    att1 = scenarios.make_attestation(ospec, ostate, 8, 0, head_root=b"\x01" * 32)
    att2 = scenarios.make_attestation(ospec, ostate, 8, 0, head_root=b"\x02" * 32)          # same target epoch, different data: double vote
    ia = [PS.IndexedAttestation(sorted(com), _to_product(PS, a.data), a.signature) for a in (att1, att2)]
    just = PS.Checkpoint(1, b"\x07" * 32)
    store = PS.Store(0, 0, just, just, just, PS.ZERO32, set())
    store.block_states[just.root] = pstate
    pspec.on_attester_slashing(store, PS.AttesterSlashing(ia[0], ia[1]))
    assert store.equivocating_indices == set(com)
    store2 = PS.Store(0, 0, just, just, just, PS.ZERO32, set())
    store2.block_states[just.root] = pstate
    with pytest.raises(AssertionError):                     # not slashable: identical data
        pspec.on_attester_slashing(store2, PS.AttesterSlashing(ia[0], ia[0]))
    bad = PS.IndexedAttestation(sorted(com), ia[1].data, ia[0].signature)                   # signature of the other message
    with pytest.raises(AssertionError):
        pspec.on_attester_slashing(store2, PS.AttesterSlashing(ia[0], bad))
    assert store2.equivocating_indices == set()


def test_decode_attestations_roundtrip(env):
    """Spec.decode_attestations(serialize_attestation(a)) == a for real (signed) attestations, through the device decode
    (b2_attestations_decode); the decoded objects are processed like the originals; a malformed encoding comes back as None; and
    the product's encoder agrees with the oracle's."""
    from oracle import ssz
    ospec, ostate, pspec, PS, PB, pks = env
    atts = [_to_product(PS, scenarios.make_attestation(ospec, ostate, 8, 0)),
            _to_product(PS, scenarios.make_attestation(ospec, ostate, 8, 1, bits=[True, False, True, True])),
            _to_product(PS, scenarios.make_attestation(ospec, ostate, 5, 1))]
    enc = [PS.serialize_attestation(a) for a in atts]
    for a, e in zip(atts, enc):
        assert e == ssz.serialize_attestation(list(a.aggregation_bits), PS.serialize_attestation_data(a.data), bytes(a.signature))
    got = pspec.decode_attestations(enc + [enc[0][:-1] + b"\x00"])
    assert got[3] is None
    for a, g in zip(atts, got[:3]):
        assert list(g.aggregation_bits) == list(a.aggregation_bits) and g.data == a.data and bytes(g.signature) == bytes(a.signature)
    sp, sq = _state_to_product(PS, ostate), _state_to_product(PS, ostate)
    pspec.process_attestations(sp, got[:3])
    pspec.process_attestations(sq, atts)
    assert sp.balances == sq.balances and sp.current_epoch_participation == sq.current_epoch_participation


def test_get_head_fast_path_at_one_million_validators():
    """the pyspec-signature get_head(store) re-marshalled the whole Store on every call.  Now the store is
    mirrored on the device: the first call uploads (registry: decompress + KeyValidate of 2^20 pubkeys, block tree, 2^20 latest
    messages), later calls are the C-ABI get_head plus O(1) host checks -- under a millisecond -- and votes that arrive through
    update_latest_messages move the device table incrementally.  Heads are checked against a from-scratch Spec (fresh upload)."""
    import time
    from pos_evolution_b200 import spec as PS
    from pos_evolution_b200.engine import Engine
    n, n_blocks = 1 << 20, 400
    pks = scenarios.pubkeys(64)
    rng = np.random.default_rng(12)
    bal = rng.choice([32, 32, 32, 31, 24, 16], size=n)
    validators = [PS.Validator(pks[i & 63], int(bal[i]) * 10**9) for i in range(n)]
    state = PS.BeaconState(slot=64, fork=PS.Fork(), genesis_validators_root=bytes(32), validators=validators, balances=[], randao_mixes=[],
                           block_roots=[], previous_epoch_participation=[], current_epoch_participation=[])
    parent, slot, roots, leaf_viable = scenarios.fork_tree(n_blocks, 9)
    rb = [bytes(r) for r in roots]
    just = PS.Checkpoint(1, rb[0])
    store = PS.Store(65 * 12, 0, just, just, just, rb[n_blocks - 1], set(int(v) for v in rng.choice(n, size=1000, replace=False)))
    leaf_state = PS.BeaconState(slot=64, fork=PS.Fork(), genesis_validators_root=bytes(32), validators=[], balances=[], randao_mixes=[], block_roots=[],
                                previous_epoch_participation=[], current_epoch_participation=[], current_justified_checkpoint=just, finalized_checkpoint=just)
    for b in range(n_blocks):
        store.blocks[rb[b]] = PS.BeaconBlock(int(slot[b]), rb[parent[b]] if b else bytes(32))
        store.block_states[rb[b]] = leaf_state
    store.checkpoint_states[just] = state
    voted = (n_blocks - 1 - np.minimum(n_blocks - 1, rng.geometric(0.02, size=n))).astype(np.int64)
    msgs = [PS.LatestMessage(1, rb[b]) for b in range(n_blocks)]
    store.latest_messages = {v: msgs[voted[v]] for v in range(n) if v % 100 != 7}
    sp = PS.Spec(PS.MAINNET, engine=Engine(0))
    t0 = time.perf_counter()
    h0 = sp.get_head(store)
    first_s = time.perf_counter() - t0
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        h = sp.get_head(store)
        lat.append(time.perf_counter() - t0)
        assert h == h0
    p50 = sorted(lat)[len(lat) // 2]
    print("Spec.get_head at 2^20 validators: first call %.2f s, then p50 %.0f us" % (first_s, p50 * 1e6))
    assert p50 < 1e-3, "Spec.get_head after the first call must be the device call plus O(1) host work"
    assert sp.stats["store_uploads"] == 1 and sp.stats["registry_uploads"] == 1
    # a committee's worth of votes for a fresh fork tip, through the spec function: incremental on the device
    att = PS.Attestation([True] * 512, PS.AttestationData(64, 0, rb[n_blocks - 2], PS.Checkpoint(0, rb[0]), PS.Checkpoint(2, rb[0])), bytes(96))
    for k in range(8):
        sp.update_latest_messages(store, [int(v) for v in rng.choice(n, size=512, replace=False)], att)
    t0 = time.perf_counter()
    h1 = sp.get_head(store)
    assert time.perf_counter() - t0 < 1e-3 and sp.stats["store_uploads"] == 1 and sp.stats["lmd_device_updates"] == 8
    w1 = sp.get_weight(store, rb[n_blocks - 2])
    fresh = PS.Spec(PS.MAINNET, engine=Engine(0))
    assert fresh.get_head(store) == h1 and fresh.get_weight(store, rb[n_blocks - 2]) == w1
    # the weight of the justified root == every counted balance + the boost: checked against plain Python over the store
    eq = store.equivocating_indices
    total = sum(validators[v].effective_balance for v in store.latest_messages if v not in eq)
    boost = (n // 32) * (int(bal.astype(object).sum()) * 10**9 // n) * 40 // 100
    assert sp.get_weight(store, rb[0]) == total + boost
