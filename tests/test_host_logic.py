"""CPU tests of the product's host layer (no GPU): the C-ABI library loads and exports every symbol that
include/b200pos.h declares, and the Python-side marshaling (shuffle, SSZ signing roots, bit packing,
store -> array conversion) agrees with the oracle's restatement."""
import copy
import hashlib
import os
import re

import numpy as np
import pytest

import scenarios
from oracle import fast
from oracle import spec as OS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from pos_evolution_b200 import _lib
    lib = _lib.load()                       # raises if a prototype's symbol is missing
    hdr = open(os.path.join(ROOT, "include", "b200pos.h")).read()
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("b2_ctx")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pos_evolution_b200._lib import B2Error
    from pos_evolution_b200.engine import Engine
    with pytest.raises(B2Error):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pos_evolution_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_shuffle_and_committees_match_oracle():
    from pos_evolution_b200 import spec as PS
    for n, tag, rounds in ((1, b"a", 10), (37, b"b", 10), (300, b"c", 90), (5000, b"d", 90)):
        seed = hashlib.sha256(tag).digest()
        assert np.array_equal(PS.shuffle_permutation(n, seed, rounds), fast.shuffle_permutation(n, seed, rounds))
    ospec, ostate = scenarios.minimal_state(64, slot=9, pks=[bytes(48)] * 64)

    class _NoGpu:
        pass
    ps = PS.Spec(PS.MINIMAL, engine=_NoGpu())
    for slot in range(8, 16):
        for idx in range(2):
            assert ps.get_beacon_committee(ostate, slot, idx) == ospec.get_beacon_committee(ostate, slot, idx)
    seed = hashlib.sha256(b"x").digest()
    for i in range(20):
        assert ps.compute_shuffled_index(i, 20, seed) == ospec.compute_shuffled_index(i, 20, seed)
    assert ps.get_beacon_proposer_index(ostate) == ospec.get_beacon_proposer_index(ostate)


def test_signing_root_matches_oracle():
    from pos_evolution_b200 import spec as PS
    ospec, ostate = scenarios.minimal_state(8, slot=9, pks=[bytes(48)] * 8)

    class _NoGpu:
        pass
    ps = PS.Spec(PS.MINIMAL, engine=_NoGpu())
    od = OS.AttestationData(9, 1, b"\x11" * 32, OS.Checkpoint(0, b"\x22" * 32), OS.Checkpoint(1, b"\x33" * 32))
    pd = PS.AttestationData(9, 1, b"\x11" * 32, PS.Checkpoint(0, b"\x22" * 32), PS.Checkpoint(1, b"\x33" * 32))
    dom_o = ospec.get_domain(ostate, OS.DOMAIN_BEACON_ATTESTER, 1)
    dom_p = ps.get_domain(ostate, PS.DOMAIN_BEACON_ATTESTER, 1)
    assert dom_o == dom_p
    assert ospec.compute_signing_root(od, dom_o) == ps.compute_signing_root(pd, dom_p)


def test_pack_bits():
    from pos_evolution_b200.spec import pack_bits
    rows = [[True, False, True], [False] * 9 + [True], []]
    out = pack_bits(rows)
    assert out.shape == (3, 2) and out[0, 0] == 0b101 and out[1, 1] == 0b10 and out[2].sum() == 0


def test_store_to_arrays_topological():
    from pos_evolution_b200 import spec as PS

    class _NoGpu:
        pass
    ps = PS.Spec(PS.MINIMAL, engine=_NoGpu())
    parent, slot, roots, viable = scenarios.fork_tree(60, 2)
    rb = [bytes(r) for r in roots]
    just = PS.Checkpoint(1, rb[0])
    store = PS.Store(0, 0, just, just, just, PS.ZERO32, set())
    st = PS.BeaconState(9, PS.Fork(), bytes(32), [], [], [], [], [], [], current_justified_checkpoint=just, finalized_checkpoint=just)
    for b in range(60):
        store.blocks[rb[b]] = PS.BeaconBlock(int(slot[b]), rb[parent[b]] if b else bytes(32))
        store.block_states[rb[b]] = st
    order, index, p2, s2, r2, v2 = ps._store_arrays(store)
    assert sorted(order) == sorted(rb) and order[0] == rb[0]
    for i in range(1, 60):
        assert p2[i] < i and order[p2[i]] == store.blocks[order[i]].parent_root


def test_ffg_host_logic_matches_oracle():
    """The scalar half of process_justification_and_finalization in the product (2/3 tests, bit shifting, the four finalization rules;
    pos-evolution.md:817-852) against the oracle on 80 generated end-of-epoch states.  The balance sums come from a stand-in for
    b2_ffg_balances here (the kernel itself is checked on the GPU, tests/test_gpu_participation.py)."""
    import copy
    from pos_evolution_b200 import spec as PS
    ospec, ostate = scenarios.minimal_state(64, slot=9, pks=[bytes(48)] * 64)

    class _SumsOnly:                                    # Engine stand-in: the three sums computed on the host from what was "uploaded"
        def registry_load(self, pk, eff, flags):
            self.eff, self.flags = np.asarray(eff, dtype=np.uint64), np.asarray(flags, dtype=np.uint8)

        registry_update_balances = lambda self, eff, flags: self.registry_load(None, eff, flags)     # noqa: E731

        def participation_load(self, which, table):
            self.__dict__.setdefault("part", {})[which] = np.asarray(table, dtype=np.uint8)

        def ffg_balances(self, flag):
            act, sl, actp = (self.flags & 1) != 0, (self.flags & 2) != 0, (self.flags & 4) != 0
            s = lambda m: int(self.eff[m].astype(object).sum()) if m.any() else 0                   # noqa: E731
            cur, prev = (((self.part[w] >> flag) & 1) != 0 for w in (0, 1))
            return s(act), s(act & ~sl & cur), s(actp & ~sl & prev), s(act & ~sl)

    ps = PS.Spec(PS.MINIMAL, engine=_SumsOnly())
    for seed in range(80):
        so = scenarios.ffg_case(copy.deepcopy(ostate), seed)
        sp = PS.BeaconState(
            slot=so.slot, fork=PS.Fork(so.fork.previous_version, so.fork.current_version, so.fork.epoch),
            genesis_validators_root=so.genesis_validators_root,
            validators=[PS.Validator(v.pubkey, v.effective_balance, v.slashed, v.activation_epoch, v.exit_epoch) for v in so.validators],
            balances=list(so.balances), randao_mixes=list(so.randao_mixes), block_roots=list(so.block_roots),
            previous_epoch_participation=list(so.previous_epoch_participation), current_epoch_participation=list(so.current_epoch_participation),
            previous_justified_checkpoint=PS.Checkpoint(so.previous_justified_checkpoint.epoch, so.previous_justified_checkpoint.root),
            current_justified_checkpoint=PS.Checkpoint(so.current_justified_checkpoint.epoch, so.current_justified_checkpoint.root),
            finalized_checkpoint=PS.Checkpoint(so.finalized_checkpoint.epoch, so.finalized_checkpoint.root),
            justification_bits=list(so.justification_bits))
        ospec.process_justification_and_finalization(so)
        ps.process_justification_and_finalization(sp)
        assert scenarios.ffg_outcome(sp) == scenarios.ffg_outcome(so), seed


def test_attestation_wire_form_matches_oracle():
    """serialize_attestation / deserialize_attestation_data of the product against the oracle's SSZ (de)serialiser."""
    from oracle import ssz
    from pos_evolution_b200 import spec as PS
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 9, 63, 64, 512, 2048):
        bits = [bool(b) for b in rng.integers(0, 2, size=n)]
        data = PS.AttestationData(int(rng.integers(1 << 40)), int(rng.integers(64)), bytes(rng.integers(0, 256, 32, dtype=np.uint8)),
                                  PS.Checkpoint(int(rng.integers(1 << 30)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))),
                                  PS.Checkpoint(int(rng.integers(1 << 30)), bytes(rng.integers(0, 256, 32, dtype=np.uint8))))
        sig = bytes(rng.integers(0, 256, 96, dtype=np.uint8))
        enc = PS.serialize_attestation(PS.Attestation(bits, data, sig))
        d128 = PS.serialize_attestation_data(data)
        assert enc == ssz.serialize_attestation(bits, d128, sig)
        got_bits, got_data, got_sig = ssz.deserialize_attestation(enc, 2048)
        assert got_bits == bits and got_sig == sig and PS.deserialize_attestation_data(got_data) == data


# ------------------------------------------------------------------------------------------ registry cache / store mirror (host logic)
class _MirrorEngine:
    """Engine stand-in that records what the host layer uploads and answers get_head / get_weights with the numpy oracle over
    exactly the tables it was given -- so a stale or missing upload shows up as a wrong head."""

    def __init__(self):
        self.log = []

    def registry_load(self, pk, eff, flags):
        self.pk = np.array(pk, dtype=np.uint8).reshape(-1, 48).copy()
        self.n = self.pk.shape[0]
        self.eff, self.flags = np.array(eff, dtype=np.uint64), np.array(flags, dtype=np.uint8)
        self.m_epoch, self.m_block = np.zeros(self.n, np.uint64), np.zeros(self.n, np.uint32)
        self.m_has, self.equiv = np.zeros(self.n, np.uint8), np.zeros(self.n, np.uint8)
        self.log.append("registry_load")
        return np.ones(self.n, dtype=np.uint8)

    def registry_update_balances(self, eff, flags):
        self.eff, self.flags = np.array(eff, dtype=np.uint64), np.array(flags, dtype=np.uint8)
        self.log.append("balances")

    def tree_load(self, parent, slot, roots, viable):
        self.parent, self.roots, self.viable = np.array(parent), np.array(roots), np.array(viable)
        self.log.append("tree_load")

    def latest_messages_load(self, epoch, blk, has, eq):
        self.m_epoch, self.m_block = np.array(epoch, dtype=np.uint64), np.array(blk, dtype=np.uint32)
        self.m_has, self.equiv = np.array(has, dtype=np.uint8), np.array(eq, dtype=np.uint8)
        self.log.append("lmd_load")

    def latest_messages_update(self, members, off, bits, target_epoch, block_idx, accept=None):
        from oracle import fast
        for a in range(len(off) - 1):
            sel = [int(members[off[a] + j]) for j in range(off[a + 1] - off[a]) if (bits[a, j >> 3] >> (j & 7)) & 1]
            fast.lmd_update(self.m_epoch, self.m_block, self.m_has, self.equiv, sel, int(target_epoch[a]), int(block_idx[a]))
        self.log.append("lmd_update")

    def on_attester_slashing(self, i1, i2):
        for v in set(i1) & set(i2):
            self.equiv[v] = 1
        self.log.append("slashing")

    def fast_aggregate_verify(self, members, off, bits, msgs, sigs):
        self.log.append("verify")
        return np.ones(len(off) - 1, dtype=np.uint8)

    def get_weights(self, boost_idx, boost_score):
        from oracle import fast
        return fast.ghost_weights(self.parent, self.m_block, self.m_has, self.eff, self.flags & 1, self.equiv, boost_idx, boost_score)

    def get_head(self, justified_idx, boost_idx, boost_score):
        from oracle import fast
        w = self.get_weights(boost_idx, boost_score)
        return fast.ghost_head(self.parent, self.roots, fast.ghost_viable(self.parent, self.viable), w, justified_idx)


def _mirror_world(PS, n_val=64, n_blocks=40):
    parent, slot, roots, leaf_viable = scenarios.fork_tree(n_blocks, 6)
    rb = [bytes(r) for r in roots]
    rng = np.random.default_rng(3)
    pks = [hashlib.sha256(b"pk%d" % i).digest() + bytes(16) for i in range(n_val)]
    vals = [PS.Validator(pks[i], int(rng.integers(16, 33)) * 10**9) for i in range(n_val)]
    state = PS.BeaconState(slot=9, fork=PS.Fork(), genesis_validators_root=bytes(32), validators=vals, balances=[0] * n_val, randao_mixes=[bytes(32)] * 64,
                           block_roots=[bytes(32)] * 64, previous_epoch_participation=[0] * n_val, current_epoch_participation=[0] * n_val)
    just = PS.Checkpoint(1, rb[0])
    store = PS.Store(9 * 12, 0, just, just, just, PS.ZERO32, set())
    for b in range(n_blocks):
        store.blocks[rb[b]] = PS.BeaconBlock(int(slot[b]), rb[parent[b]] if b else bytes(32))
        bs = copy_state(PS, state)
        bs.current_justified_checkpoint, bs.finalized_checkpoint = just, just
        store.block_states[rb[b]] = bs
    store.checkpoint_states[just] = state
    for v in range(n_val):
        if rng.random() < 0.8:
            store.latest_messages[v] = PS.LatestMessage(1, rb[int(rng.integers(n_blocks))])
    return store, state, rb, slot


def copy_state(PS, s):
    import copy
    c = copy.copy(s)
    return c


def _scratch_head(PS, store):
    """the head a brand-new Spec (no caches) computes for the store"""
    return PS.Spec(PS.MINIMAL, engine=_MirrorEngine()).get_head(store)


def test_store_mirror_is_incremental_and_never_stale():
    from pos_evolution_b200 import spec as PS
    eng = _MirrorEngine()
    sp = PS.Spec(PS.MINIMAL, engine=eng)
    store, state, rb, slot = _mirror_world(PS)
    h0 = sp.get_head(store)
    assert eng.log == ["registry_load", "balances", "tree_load", "lmd_load"] and h0 == _scratch_head(PS, store)
    eng.log.clear()
    assert sp.get_head(store) == h0 and sp.get_weight(store, rb[0]) >= sp.get_weight(store, rb[1])
    assert eng.log == []                                       # nothing re-uploaded, nothing re-read: the fast path
    # votes through update_latest_messages reach the device table incrementally (K7), not by re-upload
    att = PS.Attestation([True] * 4, PS.AttestationData(8, 0, rb[7], PS.Checkpoint(0, rb[0]), PS.Checkpoint(2, rb[0])), bytes(96))
    sp.update_latest_messages(store, [1, 2, 3, 60], att)
    assert eng.log == ["lmd_update"]
    assert sp.get_head(store) == _scratch_head(PS, store)
    assert eng.log == ["lmd_update"]
    # an older vote changes nothing anywhere; a vote for a block the tree does not hold drops the mirror (full re-upload next time)
    old = PS.Attestation([True] * 4, PS.AttestationData(8, 0, rb[9], PS.Checkpoint(0, rb[0]), PS.Checkpoint(1, rb[0])), bytes(96))
    sp.update_latest_messages(store, [1, 2], old)
    assert store.latest_messages[1].root == rb[7] and sp.get_head(store) == _scratch_head(PS, store)
    eng.log.clear()
    ghost = PS.Attestation([True], PS.AttestationData(8, 0, b"\x99" * 32, PS.Checkpoint(0, rb[0]), PS.Checkpoint(3, rb[0])), bytes(96))
    sp.update_latest_messages(store, [5], ghost)
    assert sp.get_head(store) == _scratch_head(PS, store) and "lmd_load" in eng.log
    # the store changes shape behind the class's back: a new block, a message written straight into the dict, an equivocator
    eng.log.clear()
    store.blocks[b"\x42" * 32] = PS.BeaconBlock(int(slot[7]) + 1, rb[7])
    store.block_states[b"\x42" * 32] = store.block_states[rb[7]]
    assert sp.get_head(store) == _scratch_head(PS, store) and eng.log == ["tree_load", "lmd_load"]
    store.latest_messages[63] = PS.LatestMessage(4, b"\x42" * 32)
    assert sp.get_head(store) == _scratch_head(PS, store)
    store.equivocating_indices.add(1)
    assert sp.get_head(store) == _scratch_head(PS, store)
    # the proposer boost moves: no upload, new score
    eng.log.clear()
    store.proposer_boost_root = rb[3]
    assert sp.get_head(store) == _scratch_head(PS, store) and eng.log == []
    # a new justified checkpoint state (other balances): balances are re-read, pubkeys are not uploaded again
    st2 = copy_state(PS, state)
    st2.validators = [PS.Validator(v.pubkey, 17 * 10**9) for v in state.validators]
    store.checkpoint_states[store.justified_checkpoint] = st2
    eng.log.clear()
    assert sp.get_head(store) == _scratch_head(PS, store)
    assert "registry_load" not in eng.log and "balances" in eng.log


def test_registry_cache_is_keyed_on_content_not_on_addresses():
    """(id(validators), len) can alias after GC and misses an in-place pubkey replacement."""
    from pos_evolution_b200 import spec as PS
    eng = _MirrorEngine()
    sp = PS.Spec(PS.MINIMAL, engine=eng)
    store, state, rb, slot = _mirror_world(PS)
    ia = PS.IndexedAttestation([1, 2, 3], PS.AttestationData(8, 0, rb[1], PS.Checkpoint(0, rb[0]), PS.Checkpoint(1, rb[0])), bytes(96))
    sp.are_valid_indexed_attestations(state, [ia])
    assert eng.log.count("registry_load") == 1
    # another state object with an equal validator set (a target checkpoint state): no second upload
    twin = copy_state(PS, state)
    twin.validators = [PS.Validator(v.pubkey, v.effective_balance) for v in state.validators]
    sp.are_valid_indexed_attestations(twin, [ia])
    assert eng.log.count("registry_load") == 1
    # in-place replacement of one pubkey in the SAME list object: must re-upload (signatures would be checked against a stale key)
    state.validators[10] = PS.Validator(hashlib.sha256(b"other").digest() + bytes(16), state.validators[10].effective_balance)
    sp.are_valid_indexed_attestations(state, [ia])
    assert eng.log.count("registry_load") == 2 and bytes(eng.pk[10]) == bytes(state.validators[10].pubkey)
    # a same-length registry with different keys, whatever address it lands on
    for k in range(3):
        other = copy_state(PS, state)
        other.validators = [PS.Validator(hashlib.sha256(b"gen%d/%d" % (k, i)).digest() + bytes(16), 32 * 10**9) for i in range(len(state.validators))]
        sp.are_valid_indexed_attestations(other, [ia])
        assert bytes(eng.pk[0]) == bytes(other.validators[0].pubkey)
        del other
    assert eng.log.count("registry_load") == 5


def test_get_unslashed_participating_indices_matches_oracle():
    """the product's set form called Spec.has_flag, which did not exist."""
    from pos_evolution_b200 import spec as PS
    ospec, ostate = scenarios.minimal_state(64, slot=17, pks=[bytes(48)] * 64)
    rng = np.random.default_rng(1)
    ostate.current_epoch_participation = [int(x) for x in rng.integers(0, 8, size=64)]
    ostate.previous_epoch_participation = [int(x) for x in rng.integers(0, 8, size=64)]
    for i in (3, 9, 40):
        ostate.validators[i].slashed = True
    ostate.validators[5].exit_epoch = 1
    ps = PS.Spec(PS.MINIMAL, engine=_MirrorEngine())
    pstate = PS.BeaconState(slot=ostate.slot, fork=PS.Fork(), genesis_validators_root=bytes(32),
                            validators=[PS.Validator(v.pubkey, v.effective_balance, v.slashed, v.activation_epoch, v.exit_epoch) for v in ostate.validators],
                            balances=list(ostate.balances), randao_mixes=list(ostate.randao_mixes), block_roots=list(ostate.block_roots),
                            previous_epoch_participation=list(ostate.previous_epoch_participation),
                            current_epoch_participation=list(ostate.current_epoch_participation))
    for flag in range(3):
        for epoch in (ospec.get_previous_epoch(ostate), ospec.get_current_epoch(ostate)):
            assert ps.get_unslashed_participating_indices(pstate, flag, epoch) == ospec.get_unslashed_participating_indices(ostate, flag, epoch)
    assert PS.Spec.has_flag(0b101, 2) and not PS.Spec.has_flag(0b101, 1) and PS.Spec.add_flag(0b001, 2) == 0b101


def test_generated_constants_are_current():
    """csrc/consts.cuh (field constants, Frobenius coefficients, the exponent schedules with their run token) is exactly what
    tools/gen_consts.py renders: the generator self-checks every schedule against pow() when it builds them."""
    import subprocess
    import sys
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_consts.py"), "--check"])
    assert rc == 0, "pos_evolution_b200/csrc/consts.cuh is stale: run python tools/gen_consts.py"
