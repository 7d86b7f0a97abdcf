"""Namespace in which the reference's own python blocks (tests/ref_blocks.py) can run: pyspec
integer types, preset constants and the helpers the reference calls but does not define (taken
from the oracle's restatement, oracle/spec.py)."""
from typing import Sequence, Set, Dict  # noqa: F401

import ref_blocks
from oracle import spec as S


class uint8(int):
    width = 1


class uint32(int):
    width = 4


class uint64(int):
    width = 8


def uint_to_bytes(v):
    return int(v).to_bytes(getattr(v, "width", 8), "little")


REF_FUNCS = ["get_committee_count_per_slot", "get_seed", "compute_committee", "compute_shuffled_index",
             "compute_proposer_index", "process_attestation", "get_head", "update_latest_messages",
             "process_justification_and_finalization", "weigh_justification_and_finalization"]


def namespace(spec: S.Spec):
    p = spec.p
    ns = dict(
        uint8=uint8, uint32=uint32, uint64=uint64, uint_to_bytes=uint_to_bytes, bytes_to_uint64=S.bytes_to_uint64,
        hash=S.hash32, Gwei=int, Epoch=int, max=max, min=min,
        SLOTS_PER_EPOCH=p.SLOTS_PER_EPOCH, MAX_COMMITTEES_PER_SLOT=p.MAX_COMMITTEES_PER_SLOT,
        TARGET_COMMITTEE_SIZE=p.TARGET_COMMITTEE_SIZE, SHUFFLE_ROUND_COUNT=p.SHUFFLE_ROUND_COUNT,
        EPOCHS_PER_HISTORICAL_VECTOR=p.EPOCHS_PER_HISTORICAL_VECTOR, MIN_SEED_LOOKAHEAD=p.MIN_SEED_LOOKAHEAD,
        MIN_ATTESTATION_INCLUSION_DELAY=p.MIN_ATTESTATION_INCLUSION_DELAY, MAX_EFFECTIVE_BALANCE=p.MAX_EFFECTIVE_BALANCE,
        PARTICIPATION_FLAG_WEIGHTS=S.PARTICIPATION_FLAG_WEIGHTS, WEIGHT_DENOMINATOR=S.WEIGHT_DENOMINATOR,
        PROPOSER_WEIGHT=S.PROPOSER_WEIGHT, LatestMessage=S.LatestMessage,
        # helpers the reference calls but never defines ("ext")
        get_active_validator_indices=spec.get_active_validator_indices, get_randao_mix=spec.get_randao_mix,
        get_previous_epoch=spec.get_previous_epoch, get_current_epoch=spec.get_current_epoch,
        compute_epoch_at_slot=spec.compute_epoch_at_slot,
        get_attestation_participation_flag_indices=spec.get_attestation_participation_flag_indices,
        get_base_reward=spec.get_base_reward, has_flag=spec.has_flag, add_flag=spec.add_flag,
        increase_balance=spec.increase_balance,
        get_filtered_block_tree=spec.get_filtered_block_tree,
        # FFG accounting (ref :793-852): constants and the helpers described only in prose (:805-811)
        GENESIS_EPOCH=S.GENESIS_EPOCH, TIMELY_TARGET_FLAG_INDEX=S.TIMELY_TARGET_FLAG_INDEX,
        JUSTIFICATION_BITS_LENGTH=S.JUSTIFICATION_BITS_LENGTH, Checkpoint=S.Checkpoint, get_block_root=spec.get_block_root,
        get_unslashed_participating_indices=spec.get_unslashed_participating_indices,
        get_total_active_balance=spec.get_total_active_balance, get_total_balance=spec.get_total_balance,
        get_latest_attesting_balance=spec.get_latest_attesting_balance,
    )
    ref_blocks.exec_functions(REF_FUNCS, ns)

    # ext helpers that must call the *reference* committee code, not the oracle's
    def get_beacon_committee(state, slot, index):
        epoch = spec.compute_epoch_at_slot(slot)
        cps = ns["get_committee_count_per_slot"](state, epoch)
        return ns["compute_committee"](indices=spec.get_active_validator_indices(state, epoch),
                                       seed=ns["get_seed"](state, epoch, S.DOMAIN_BEACON_ATTESTER),
                                       index=(slot % p.SLOTS_PER_EPOCH) * cps + index, count=cps * p.SLOTS_PER_EPOCH)

    def get_attesting_indices(state, data, bits):
        committee = get_beacon_committee(state, data.slot, data.index)
        return set(v for i, v in enumerate(committee) if bits[i])

    def get_indexed_attestation(state, attestation):
        return S.IndexedAttestation(sorted(get_attesting_indices(state, attestation.data, attestation.aggregation_bits)),
                                    attestation.data, attestation.signature)

    def get_beacon_proposer_index(state):
        epoch = spec.get_current_epoch(state)
        seed = S.hash32(ns["get_seed"](state, epoch, S.DOMAIN_BEACON_PROPOSER) + int(state.slot).to_bytes(8, "little"))
        return ns["compute_proposer_index"](state, spec.get_active_validator_indices(state, epoch), seed)

    ns.update(get_beacon_committee=get_beacon_committee, get_attesting_indices=get_attesting_indices,
              get_indexed_attestation=get_indexed_attestation, get_beacon_proposer_index=get_beacon_proposer_index,
              is_valid_indexed_attestation=spec.is_valid_indexed_attestation)
    return ns
