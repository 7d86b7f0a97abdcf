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
